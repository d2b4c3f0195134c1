#!/bin/bash
# round 2, GPU session K: rolling estimate_many (cmax + EKLT) -- parity tests, concurrency sweeps, host time per window
cd "$(dirname "$0")/../.."
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_zz_eklt.py -q -s --timeout=600 -p no:cacheprovider -k "estimate_many or drop_in or preprocessing or solver_final" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt
for c in 3 4 6 8 12; do
  timeout 300 python bench.py --workload solve --solve-concurrency $c --no-cpu > $O/solve_c$c.json 2> $O/solve_c$c.err
done
for c in 1 2 4 8; do
  timeout 300 python bench.py --workload eklt --eklt-concurrency $c --steps 12 --no-cpu > $O/eklt_c$c.json 2> $O/eklt_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02k/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],3), d.get("host_ms_per_window"), d.get("ms_per_window_single"), d.get("eval_ms_per_level"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-600:])
PY
