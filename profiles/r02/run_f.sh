#!/bin/bash
# round 2, GPU session F: blurred-objective tests, TV-after-splat step time, ncu launch list + full captures of the default step
cd "$(dirname "$0")/../.."
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_fused.py tests/test_gpu_dropin.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
timeout 600 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --no-cpu --no-e2e --no-subrecords --steps 3 --warmup 1 > $O/launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_tile_splat_d|k_win_bwd_g|k_gradmag_sep|k_flow_tv_march" -c 4 -s 20 -o $O/r02_default_kernels -f python bench.py --no-e2e --no-cpu --no-subrecords --steps 3 --warmup 1 > $O/ncu_full.log 2>&1
tail -4 $O/pytest.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f/bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["step_roofline"]["frac"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print(json.dumps(d["small_windows"])[:400]); print(d["solve"]["value"], d["giant"]["ms_per_step"], d["e2e"]["value"])
PY
