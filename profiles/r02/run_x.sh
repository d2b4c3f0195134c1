#!/bin/bash
# round 2, GPU session X: flow gathers into registers of their own + selects (EBOS_GATHER_SELECT build) -- parity + A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02x; mkdir -p $O
ALT=$PWD/event_based_bos_b200/libebos_alt.so
EBOS_LIBRARY=$ALT timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py -q --timeout=600 -p no:cacheprovider > $O/pytest_alt.txt 2>&1; echo "rc=$?" >> $O/pytest_alt.txt
tail -4 $O/pytest_alt.txt | cut -c1-200
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
timeout 300 $B > $O/bench_base.json 2> $O/bench_base.err
EBOS_LIBRARY=$ALT timeout 300 $B > $O/bench_alt.json 2> $O/bench_alt.err
timeout 300 $B > $O/bench_base2.json 2> $O/bench_base2.err
EBOS_LIBRARY=$ALT timeout 300 $B > $O/bench_alt2.json 2> $O/bench_alt2.err
EBOS_LIBRARY=$ALT timeout 300 $B --events 500000 > $O/bench_alt_500k.json 2> $O/bench_alt_500k.err
timeout 300 $B --events 500000 > $O/bench_base_500k.json 2> $O/bench_base_500k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02x/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"],4), d["roofline"]["frac"], d.get("roofline",{}).get("kernel_ms"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
