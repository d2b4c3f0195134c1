#!/bin/bash
# round 2, GPU session H2: Adam grid with the same number of items per thread -- A/B on the solve
cd "$(dirname "$0")/../.."
O=gpurun_out/r02h2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused.py -q --timeout=600 -p no:cacheprovider -k "adam or solver" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -2 $O/pytest.txt | cut -c1-200
for c in 1 8; do
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_c$c.json 2> $O/solve_c$c.err
EBOS_ADAM_GRID_LEGACY=1 timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_legacy_c$c.json 2> $O/solve_legacy_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h2/*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"],2), round(d["ms_per_step"],4))
PY
