#!/bin/bash
# round 2, GPU session C2: launch granularity of the microbench (1 vs 5 / 10 evaluations per executable graph)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02c2; mkdir -p $O
timeout 600 python bench.py --no-e2e --no-cpu --steps 20 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c2/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["ms_per_step_five_evaluations_per_replay"], d["ms_per_step_eager"])
print(d["small_windows"])
print(d["eklt"]["gpu_launches"], d["eklt"]["value"])
PY
tail -3 $O/bench.err
