#!/bin/bash
# round 2, GPU session I: default bench after the affinity / TV-beside-exchange changes; solve concurrency A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02i; mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
for c in 4 6 12 16; do
  timeout 600 python bench.py --workload solve --no-cpu --solve-concurrency $c --steps 48 > $O/solve_c$c.json 2> $O/solve_c$c.err
done
tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i/bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["step_roofline"]["frac"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["host_numa_binding"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "solve", d["solve"]["value"], "giant", d["giant"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for c in (4,6,12,16):
    try:
        s=json.loads(open(f"gpurun_out/r02i/solve_c{c}.json").read().strip().splitlines()[-1]); print("solve conc",c,s["value"])
    except Exception as e: print(c,"ERR",e)
PY
