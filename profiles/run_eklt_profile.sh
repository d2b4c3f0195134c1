#!/bin/bash
# EKLT inner loop on one B200: parity tests, bench lines (fp64 / fp32), ncu launch list.
#   gpurun --timeout 170 -- 'bash profiles/run_eklt_profile.sh'
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_zz_eklt.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/eklt_gpu_tests.log 2>&1
tail -3 gpurun_out/eklt_gpu_tests.log
timeout 60 python bench.py --workload eklt --steps 3 --warmup 1 > gpurun_out/eklt_bench_f64.json 2> gpurun_out/eklt_bench_f64.err
tail -c 1500 gpurun_out/eklt_bench_f64.json; tail -3 gpurun_out/eklt_bench_f64.err
timeout 40 python bench.py --workload eklt --steps 3 --warmup 1 --eklt-precision 32 --no-cpu > gpurun_out/eklt_bench_f32.json 2> gpurun_out/eklt_bench_f32.err
tail -c 600 gpurun_out/eklt_bench_f32.json
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r01h_eklt_launches.csv \
  python bench.py --workload eklt --steps 1 --warmup 1 --solve-iters 20 --eklt-no-graph --no-cpu > gpurun_out/eklt_ncu_run.log 2>&1
wc -l gpurun_out/r01h_eklt_launches.csv
