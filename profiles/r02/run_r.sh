#!/bin/bash
# round 2, GPU session R: fused Adam + TV kernel, 2 rows per thread -- parity + A/B on the solve (1 and 8 windows in flight)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py -q --timeout=600 -p no:cacheprovider -k "folded or solver or estimate_many" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt | cut -c1-300
for c in 1 8; do
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_fold_c$c.json 2> $O/solve_fold_c$c.err
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c --no-fold-tv > $O/solve_nofold_c$c.json 2> $O/solve_nofold_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02r/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],4), d.get("launches_per_iteration"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
