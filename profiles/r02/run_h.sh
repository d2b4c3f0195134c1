#!/bin/bash
# round 2, GPU session H: blocked-striped storage with the 8-events-per-thread backward; A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py -q -s --timeout=600 -p no:cacheprovider -k "variants or benchmark or dense or full_size" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run blocked A=1
run linear EBOS_LINEAR=1
run blocked_g1 EBOS_GROUPS=1
run blocked_g4 EBOS_GROUPS=4
run blocked_v2splat EBOS_SPLAT_V2=1
run blocked2 A=1
tail -4 $O/pytest.txt; for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernel_ms"]
    print(d["ms_per_step"], k['window_splat(+memset)'], k['window_backward'], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
