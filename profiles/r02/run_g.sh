#!/bin/bash
# round 2, GPU session G: blocked-striped storage of dense windows -- tests + A/B against linear storage
cd "$(dirname "$0")/../.."
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_metrics.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run blocked A=1
run linear EBOS_LINEAR=1
run blocked_v2 EBOS_SPLAT_V2=1 EBOS_TILE_BWD=2
run blocked_i16 EBOS_ITEM_EVENTS=16368
run blocked_i4 EBOS_ITEM_EVENTS=4592
run blocked_bocc5 EBOS_BOCC=5
run blocked_bocc3 EBOS_BOCC=3
tail -4 $O/pytest.txt; for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernel_ms"]
    print(d["ms_per_step"], k['window_splat(+memset)'], k['window_backward'], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
