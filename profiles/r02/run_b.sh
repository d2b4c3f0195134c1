#!/bin/bash
# round 2, GPU session B: round-2 tile kernels v2 (table + rows in use), A/B + ncu of the two
cd "$(dirname "$0")/../.."
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_metrics.py tests/test_gpu_dropin.py -q -s -k "variants or benchmark_window or metrics or dropin" --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
B="python bench.py --no-e2e --no-cpu --steps 30"
timeout 300 $B > $O/bench_base.json 2> $O/bench_base.err
EBOS_SPLAT_V2=1 timeout 300 $B > $O/bench_v2.json 2> $O/bench_v2.err
EBOS_TILE=6 timeout 300 $B > $O/bench_v2merge.json 2> $O/bench_v2merge.err
EBOS_TILE_BWD=2 timeout 300 $B > $O/bench_tbwd.json 2> $O/bench_tbwd.err
EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 timeout 300 $B > $O/bench_both.json 2> $O/bench_both.err
EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_QOCC=5 EBOS_BOCC=5 timeout 300 $B > $O/bench_both_occ5.json 2> $O/bench_both_occ5.err
EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_QOCC=3 EBOS_BOCC=3 timeout 300 $B > $O/bench_both_occ3.json 2> $O/bench_both_occ3.err
# ncu: launch list + full capture of the two new kernels
EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_tile_splat_m|k_tile_bwd_m" -c 2 -s 10 -o $O/v2_kernels -f python bench.py --no-e2e --no-cpu --steps 3 --warmup 1 > $O/ncu.log 2>&1
tail -3 $O/pytest.txt; for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
