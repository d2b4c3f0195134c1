#!/bin/bash
# round 2, multi-GPU session J2: in-switch (NVLS multimem) exchange of the event-sharded window -- parity test + timings
N=${1:-2}
cd "$(dirname "$0")/../.."
O=gpurun_out/r02j2_n$N; mkdir -p $O
if [ "$N" -le 2 ]; then
timeout 900 python -m pytest tests/test_gpu_sharding.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest_sharding.txt 2>&1; echo "rc=$?" >> $O/pytest_sharding.txt
grep -i "multimem\|passed\|failed\|rc=" $O/pytest_sharding.txt | cut -c1-200 | tail -8
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 bench.py --gpus $N --workload giant --steps 30 > $O/giant_auto.json 2> $O/giant_auto.err
EBOS_P2P_FORM=3 timeout 600 $TR --master-port 29562 bench.py --gpus $N --workload giant --steps 30 > $O/giant_multimem.json 2> $O/giant_multimem.err
[ "$N" -le 2 ] && EBOS_P2P_FORM=3 timeout 600 $TR --master-port 29563 bench.py --gpus $N --workload giant --steps 30 --giant-eager > $O/giant_multimem_eager.json 2> $O/giant_multimem_eager.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("giant_auto","giant_multimem","giant_multimem_eager"):
    try:
        d=json.loads(open(f"gpurun_out/r02j2_n{N}/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("launch")[:20], d.get("exchange")[:34], d.get("exchange_start_up_timing_ms"), d["parity_self_check"])
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/r02j2_n{N}/{f}.err").read()[-1200:])
PY
