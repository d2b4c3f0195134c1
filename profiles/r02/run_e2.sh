#!/bin/bash
# round 2, GPU session E2: aligned pair loads of dL/dIWE in the backward (shifted copy written by the cost kernel) -- parity + A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02e2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dropin.py tests/test_gpu_metrics.py -q --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -5 $O/pytest.txt | cut -c1-250
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run pairs A=1
run nopairs EBOS_NO_GPAIRS=1
run pairs2 A=1
run nopairs2 EBOS_NO_GPAIRS=1
timeout 300 python bench.py --workload solve --no-cpu > $O/solve_c8.json 2> $O/solve_c8.err
EBOS_NO_GPAIRS=1 timeout 300 python bench.py --workload solve --no-cpu > $O/solve_c8_nopairs.json 2> $O/solve_c8_nopairs.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e2/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d.get("roofline",{}).get("kernel_ms") or {}
        print(f.split('/')[-1], round(d["value"],2), round(d["ms_per_step"],4), d.get("roofline",{}).get("frac"), k.get('window_splat(+memset)'), k.get('window_backward'))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-500:])
PY
