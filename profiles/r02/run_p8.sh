#!/bin/bash
# round 2, 8-GPU session P: default bench line at 8 GPUs (solve scaling with the rolling schedule / persistent executables)
N=8
cd "$(dirname "$0")/../.."
O=gpurun_out/r02p8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02p8/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
print("e2e", d["e2e"]["value"], d["e2e"]["h2d_gbs_per_rank"])
print("solve", d["solve"]["value"], d["solve"]["ms_per_window_per_gpu"], d["solve"]["host_ms_per_window"], d["solve"]["parity_self_check"]["ok"])
print("giant", d["giant"]["ms_per_step"], d["giant"]["exchange"][:40], d["giant"].get("exchange_start_up_timing_ms"), d["giant"]["parity_self_check"]["ok"])
print("eklt", d["eklt"]["value"], d["eklt"]["ms_per_window_single"], d["eklt"]["host_ms_per_window"])
PY
