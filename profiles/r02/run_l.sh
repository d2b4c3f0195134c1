#!/bin/bash
# round 2, GPU session L: host-side profile of the fused solve
cd "$(dirname "$0")/../.."
O=gpurun_out/r02l; mkdir -p $O
timeout 300 python profiles/tools/solve_host_profile.py > $O/host_profile.txt 2>&1
cat $O/host_profile.txt | cut -c1-250 | tail -60
