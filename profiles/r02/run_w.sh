#!/bin/bash
# round 2, GPU session W: what the driver runs at round end -- smoke(), the whole -m gpu suite, the default bench line, the reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02w; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.txt 2>&1; echo "rc=$?" >> $O/smoke.txt
tail -6 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt | cut -c1-200
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_roofline"]["frac"], d["roofline"]["kernel_ms"])
print("e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"))
print("small", [(x["events"], x["ms_per_step"]) for x in d["small_windows"]])
print("solve", d["solve"]["value"], d["solve"]["host_ms_per_window"], d["solve"]["roofline"]["frac"])
print("giant", d["giant"]["ms_per_step"])
print("eklt", d["eklt"]["value"], d["eklt"]["ms_per_window_single"], d["eklt"]["eval_ms_per_level"])
r=json.loads(open("gpurun_out/r02w/bench_reference.json").read().strip().splitlines()[-1])
print("reference", r.get("value"), r.get("ms_per_step"), r.get("config",{}).get("events_per_window"))
PY
