#!/bin/bash
# round 2, 2-GPU session Y: the sharded evaluation captured as one executable graph -- parity test, eager vs replayed
N=${1:-2}
cd "$(dirname "$0")/../.."
O=gpurun_out/r02y$N; mkdir -p $O
[ "$N" -le 2 ] && timeout 900 python -m pytest tests/test_gpu_sharding.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest_sharding.txt 2>&1; echo "rc=$?" >> $O/pytest_sharding.txt
tail -4 $O/pytest_sharding.txt | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus $N --workload giant --steps 30 > $O/giant_replay.json 2> $O/giant_replay.err
timeout 600 $TR --master-port 29532 bench.py --gpus $N --workload giant --steps 30 --giant-eager > $O/giant_eager.json 2> $O/giant_eager.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("giant_replay","giant_eager"):
    try:
        d=json.loads(open(f"gpurun_out/r02y{N}/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("launch"), d.get("exchange")[:30], d.get("exchange_start_up_timing_ms"), d["parity_self_check"]["ok"])
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/r02y{N}/{f}.err").read()[-1500:])
PY
