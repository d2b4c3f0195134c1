#!/bin/bash
# round 2, GPU session D: full GPU suite, smoke, default bench (with sub-records), reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s --timeout=900 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload eklt --steps 3 --warmup 1 --no-cpu > $O/bench_eklt.json 2> $O/bench_eklt.err
tail -4 $O/pytest_gpu.txt; tail -3 $O/smoke.txt; tail -3 $O/bench_default.err
python - <<'PY'
import json
for f in ("bench_default","bench_reference","bench_eklt"):
    try:
        d=json.loads(open(f"gpurun_out/r02d/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("step_roofline") or {}).get("frac"))
        for k in ("e2e","small_windows","solve","giant","launches_per_step","gpu_launches"):
            if k in d:
                v=d[k]
                if k in ("solve","giant"): v={kk:v[kk] for kk in ("value","ms_per_step","ms_per_window_per_gpu","parity_self_check","exchange","gpu_launches") if kk in v}
                print("  ",k, json.dumps(v)[:600])
    except Exception as e: print(f,"ERR",e)
PY
