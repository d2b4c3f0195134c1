#!/bin/bash
# EKLT inner loop on one B200: parity tests, bench line (fp64, both kernel chains timed per level),
# ncu launch list, ncu --set full of the plane kernels.
#   gpurun --timeout 160 -- 'TAG=r02a EKLT_AB_TAIL=1 EKLT_AB_STORED=1 EKLT_AB_SEG=1 bash profiles/run_eklt_profile.sh'   (the r01i artefacts came from this script)
TAG=${TAG:-r02a}
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_zz_eklt.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/eklt_gpu_tests_${TAG}.log 2>&1
tail -3 gpurun_out/eklt_gpu_tests_${TAG}.log
# opt-in paths that have not run on hardware yet (EBOS_EKLT_TAIL=1): their own test first, then the A/B in the bench line
EBOS_TEST_EXPERIMENTAL=1 timeout 40 python -m pytest tests/test_gpu_zz_eklt.py -m gpu -q --tb=short -p no:cacheprovider -k experimental \
  > gpurun_out/eklt_gpu_tests_experimental.log 2>&1
tail -3 gpurun_out/eklt_gpu_tests_experimental.log
timeout 40 python bench.py --workload eklt --steps 3 --warmup 1 --no-cpu ${EKLT_AB_TAIL:+--eklt-ab-tail} ${EKLT_AB_STORED:+--eklt-ab-stored} ${EKLT_AB_SEG:+--eklt-ab-seg} ${EKLT_CACHE_GRAPHS:+--eklt-cache-graphs} > gpurun_out/eklt_bench_f64_${TAG}.json 2> gpurun_out/eklt_bench_f64_${TAG}.err
tail -c 900 gpurun_out/eklt_bench_f64_${TAG}.json; tail -3 gpurun_out/eklt_bench_f64_${TAG}.err
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${TAG}_eklt_launches.csv \
  python bench.py --workload eklt --steps 1 --warmup 1 --solve-iters 20 --eklt-no-graph --no-cpu > gpurun_out/eklt_ncu_run_${TAG}.log 2>&1
wc -l gpurun_out/${TAG}_eklt_launches.csv
timeout 40 ncu --set full --clock-control none --import-source on -k regex:'k_forward|k_backward|k_tv_roi|k_gather_cols' -c 4 \
  -o gpurun_out/${TAG}_eklt_planes -f python bench.py --workload eklt --steps 1 --warmup 1 --solve-iters 20 --eklt-no-graph --no-cpu \
  > gpurun_out/eklt_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
