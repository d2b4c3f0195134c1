#!/bin/bash
# round 2, GPU session M: persistent executables (ops.ReplaySlot) in the solvers -- parity tests, concurrency sweeps
cd "$(dirname "$0")/../.."
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_zz_eklt.py tests/test_gpu_dropin.py -q --timeout=600 -p no:cacheprovider -k "estimate_many or drop_in or solver or solve or dropin" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -5 $O/pytest.txt
timeout 300 python profiles/tools/solve_host_profile.py > $O/host_profile.txt 2>&1; grep -n "^[0-2] {\|^{'windows" $O/host_profile.txt | cut -c1-400
for c in 2 4 8 12; do
  timeout 300 python bench.py --workload solve --solve-concurrency $c --no-cpu > $O/solve_c$c.json 2> $O/solve_c$c.err
done
for c in 1 2 4 8; do
  timeout 300 python bench.py --workload eklt --eklt-concurrency $c --steps 12 --no-cpu > $O/eklt_c$c.json 2> $O/eklt_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02m/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],3), d.get("host_ms_per_window"), d.get("ms_per_window_single"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-600:])
PY
