#!/bin/bash
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r02i2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 profiles/tools/multicast_probe.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -12 | tee gpurun_out/r02i2/probe_n$N.txt
