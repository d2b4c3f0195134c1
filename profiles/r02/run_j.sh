#!/bin/bash
# round 2, GPU session J (2 GPUs): sharded parity test with all forced forms + start-up timing; giant window
cd "$(dirname "$0")/../.."
O=gpurun_out/r02j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharding.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest_sharding.txt 2>&1; echo "rc=$?" >> $O/pytest_sharding.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --workload giant --steps 30 > $O/giant_n2.json 2> $O/giant_n2.err
timeout 900 $TR --master-port 29532 bench.py --gpus 2 --no-cpu > $O/bench_default_n2.json 2> $O/bench_default_n2.err
tail -5 $O/pytest_sharding.txt | cut -c1-300
python - <<'PY'
import json
for f in ("giant_n2","bench_default_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r02j/{f}.json").read().strip().splitlines()[-1])
        g=d.get("giant", d)
        print(f, d.get("value"), d.get("ms_per_step"), "| giant:", g.get("ms_per_step"), g.get("exchange"), g.get("exchange_start_up_timing_ms"), (g.get("parity_self_check") or {}).get("ok"))
        if "solve" in d: print("  solve", d["solve"]["value"], "e2e", d["e2e"]["value"], d["host_numa_binding"])
    except Exception as e: print(f,"ERR",e)
PY
