#!/bin/bash
# round 2, multi-GPU session: bash profiles/r02/run_multi.sh N   (default bench at N GPUs + giant-window A/B of the exchange forms)
N=${1:-4}
cd "$(dirname "$0")/../.."
O=gpurun_out/r02n$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 600 $TR --master-port 29522 bench.py --gpus $N --workload giant --steps 30 > $O/giant_default.json 2> $O/giant_default.err
EBOS_P2P_FORM=1 timeout 600 $TR --master-port 29523 bench.py --gpus $N --workload giant --steps 30 > $O/giant_oneshot.json 2> $O/giant_oneshot.err
EBOS_NO_P2P=1 timeout 600 $TR --master-port 29524 bench.py --gpus $N --workload giant --steps 30 > $O/giant_nccl.json 2> $O/giant_nccl.err
tail -3 $O/bench_default.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("bench_default","giant_default","giant_oneshot","giant_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02n{N}/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("exchange"), json.dumps(d.get("parity_self_check"))[:260])
        for k in ("e2e","solve","giant"):
            if k in d:
                v=d[k]
                if k in ("solve","giant"): v={kk:v[kk] for kk in ("value","ms_per_step","ms_per_window_per_gpu","parity_self_check","exchange") if kk in v}
                if k=="e2e": v={kk:v[kk] for kk in ("value","ms_per_step","h2d_gbs_per_rank") if kk in v}
                print("  ",k, json.dumps(v)[:520])
        print("  numa", d.get("host_numa_binding"))
    except Exception as e: print(f,"ERR",e)
PY
