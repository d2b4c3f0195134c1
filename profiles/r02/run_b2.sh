#!/bin/bash
# round 2, multi-GPU session B2: the default bench line at N GPUs as the driver launches it
N=${1:-2}
cd "$(dirname "$0")/../.."
O=gpurun_out/r02b2_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"
timeout 600 $TR --master-port 29542 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02b2_n{N}/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"])
print("e2e", d["e2e"]["value"], d["e2e"]["h2d_gbs_per_rank"])
print("solve", d["solve"]["value"], d["solve"]["ms_per_window_per_gpu"], d["solve"]["parity_self_check"]["ok"])
print("giant", d["giant"]["ms_per_step"], d["giant"]["launch"], d["giant"]["exchange"][:40], d["giant"]["parity_self_check"]["ok"])
print("eklt", d["eklt"]["value"], d["eklt"]["gpu_launches"], d["eklt"]["launches_per_iteration_per_level"])
r=open(f"gpurun_out/r02b2_n{N}/bench_reference.json").read().strip().splitlines()
print("reference lines:", len(r), r[-1][:300] if r else None)
PY
