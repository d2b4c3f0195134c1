#!/bin/bash
# round 2, GPU session A2: EKLT -- column maximum by the last CTA of k_column_sums, Adam + step counter inside k_param_grad
cd "$(dirname "$0")/../.."
O=gpurun_out/r02a2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_zz_eklt.py -q --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt | cut -c1-200
EBOS_EKLT_FUSED_TAIL=0 timeout 900 python -m pytest tests/test_gpu_zz_eklt.py -q --timeout=600 -p no:cacheprovider -k "solve or drop_in or estimate_many" > $O/pytest_unfused.txt 2>&1; echo "rc=$?" >> $O/pytest_unfused.txt
tail -2 $O/pytest_unfused.txt | cut -c1-200
timeout 300 python bench.py --workload eklt --steps 12 --no-cpu > $O/eklt_fused.json 2> $O/eklt_fused.err
EBOS_EKLT_FUSED_TAIL=0 timeout 300 python bench.py --workload eklt --steps 12 --no-cpu > $O/eklt_unfused.json 2> $O/eklt_unfused.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02a2/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), round(d["ms_per_step"],3), d.get("eval_ms_per_level"), d.get("ms_per_window_single"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
