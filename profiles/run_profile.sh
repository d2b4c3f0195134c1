#!/bin/bash
# Profiling pass for one round (run under gpurun, one GPU):  bash profiles/run_profile.sh <tag>
# 1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
# 2. one full ncu capture of each kernel of the fused evaluation, condensed on the box (profiles/tools/ncu_summarize.py:
#    gpurun returns at most 64 MiB; the two event-kernel reports are ~33 MB each and are dropped after the summary)
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-subrecords"
# (k_adam only appears in the per-kernel timing section of bench.py; the solver uses it every iteration)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
for K in k_tile_splat_d k_win_bwd_g k_flow_tv_march k_gradmag_sep k_adam; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_${K} -f $BENCH > gpurun_out/${TAG}_${K}.log 2>&1
done
python profiles/tools/ncu_summarize.py gpurun_out/${TAG}_kernels_ncu_summary.txt gpurun_out/${TAG}_k_tile_splat_d.ncu-rep gpurun_out/${TAG}_k_win_bwd_g.ncu-rep gpurun_out/${TAG}_k_flow_tv_march.ncu-rep gpurun_out/${TAG}_k_gradmag_sep.ncu-rep gpurun_out/${TAG}_k_adam.ncu-rep
rm -f gpurun_out/${TAG}_k_tile_splat_d.ncu-rep gpurun_out/${TAG}_k_win_bwd_g.ncu-rep
ls -la gpurun_out | tail -12
