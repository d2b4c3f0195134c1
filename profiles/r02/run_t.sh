#!/bin/bash
# round 2, GPU session T: Adam factors from the TV kernel (no pow / barrier in k_adam), k_gradmag_sep 7160 -> 2112 SASS
# instructions (peer loop not unrolled) -- parity + timings
cd "$(dirname "$0")/../.."
O=gpurun_out/r02t; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dropin.py tests/test_gpu_zz_eklt.py -q --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt | cut -c1-300
timeout 300 python bench.py --workload solve --no-cpu > $O/solve_c8.json 2> $O/solve_c8.err
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency 1 > $O/solve_c1.json 2> $O/solve_c1.err
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 > $O/bench_16mi.json 2> $O/bench_16mi.err
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 --events 500000 > $O/bench_500k.json 2> $O/bench_500k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],4), d.get("roofline",{}).get("kernel_ms"), d.get("launches_per_iteration"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
