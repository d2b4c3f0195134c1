#!/bin/bash
# round 2, GPU session A: full GPU test-suite, A/B of the round-2 tile kernels, EKLT experimental switches
cd "$(dirname "$0")/../.."
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --timeout=900 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
B="python bench.py --no-e2e --no-cpu --steps 30"
timeout 300 $B > $O/bench_base.json 2> $O/bench_base.err
EBOS_SPLAT_MERGE=1 timeout 300 $B > $O/bench_merge.json 2> $O/bench_merge.err
EBOS_TILE_BWD=2 timeout 300 $B > $O/bench_tbwd.json 2> $O/bench_tbwd.err
EBOS_SPLAT_MERGE=1 EBOS_TILE_BWD=2 timeout 300 $B > $O/bench_both.json 2> $O/bench_both.err
EBOS_SPLAT_MERGE=1 EBOS_TILE_BWD=2 EBOS_QOCC=5 EBOS_BOCC=5 timeout 300 $B > $O/bench_both_occ5.json 2> $O/bench_both_occ5.err
EBOS_SPLAT_MERGE=1 EBOS_TILE_BWD=2 EBOS_QOCC=3 EBOS_BOCC=3 timeout 300 $B > $O/bench_both_occ3.json 2> $O/bench_both_occ3.err
EBOS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_eklt.py -q -k experimental -p no:cacheprovider > $O/eklt_experimental.txt 2>&1
timeout 600 python bench.py --workload eklt --steps 3 --warmup 1 --no-cpu --eklt-ab-tail --eklt-ab-stored --eklt-ab-seg > $O/eklt_ab.json 2> $O/eklt_ab.err
timeout 600 python bench.py --workload eklt --steps 3 --warmup 1 --no-cpu --eklt-cache-graphs > $O/eklt_cache.json 2> $O/eklt_cache.err
tail -3 $O/pytest_gpu.txt; for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
tail -5 $O/eklt_experimental.txt
