#!/bin/bash
# round 2, GPU session C: item size A/B (16 vs 32 events per thread) for the tile kernels; new metric / blur / cost tests
cd "$(dirname "$0")/../.."
O=gpurun_out/r02c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_metrics.py tests/test_gpu_dropin.py -q -s -k "variants or benchmark_window or metrics or dropin" --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
EBOS_ITEM_EVENTS=8176 timeout 900 python -m pytest tests/test_gpu_fused.py -q -s -k "variants or benchmark_window or dense_window" --timeout=600 -p no:cacheprovider > $O/pytest_8176.txt 2>&1; echo "rc=$?" >> $O/pytest_8176.txt
B="python bench.py --no-e2e --no-cpu --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run base A=1
run base_i8 EBOS_ITEM_EVENTS=8176
run v2 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2
run v2_i8 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=8176
run v2_i12 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=12272
run v2_i16 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=16368
run v2m_i8 EBOS_TILE=6 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=8176
run v2_i8_occ5 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=8176 EBOS_QOCC=5 EBOS_BOCC=5
run v2_i8_occ3 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2 EBOS_ITEM_EVENTS=8176 EBOS_QOCC=3 EBOS_BOCC=3
tail -3 $O/pytest.txt; tail -3 $O/pytest_8176.txt; for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernel_ms"]
    print(d["ms_per_step"], k['window_splat(+memset)'], k['window_backward'], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
