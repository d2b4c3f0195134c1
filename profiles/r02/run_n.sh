#!/bin/bash
# round 2, GPU session N: block-aligned item cut of dense windows (full two-pass CTAs + remainder) -- parity + A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py -q --timeout=600 -p no:cacheprovider -k "variants or benchmark or dense or full_size or value_and_grad" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run cut16 A=1
run evencut EBOS_ITEM_CUT=0
run cut8 EBOS_ITEM_EVENTS=4080
run cut24 EBOS_ITEM_EVENTS=12272
run cut32 EBOS_ITEM_EVENTS=16368
run cut16b A=1
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernel_ms"]
    print(d["ms_per_step"], k['window_splat(+memset)'], k['window_backward'], d["step_roofline"]["frac"])
except Exception as e: print("ERR", e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
done
