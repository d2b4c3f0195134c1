#!/bin/bash
# round 2, GPU session O: the whole -m gpu suite + the default bench line (all sub-records) + the reference arm, one B200
cd "$(dirname "$0")/../.."
O=gpurun_out/r02o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02o/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print("e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"))
print("small", [(x["events"], x["ms_per_step"]) for x in d["small_windows"]])
print("solve", d["solve"]["value"], d["solve"]["host_ms_per_window"], d["solve"]["roofline"]["frac"])
print("giant", d["giant"]["ms_per_step"])
print("eklt", d["eklt"]["value"], d["eklt"]["ms_per_window_single"], d["eklt"]["eval_ms_per_level"])
PY
