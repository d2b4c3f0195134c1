#!/bin/bash
# round 2, GPU session V: gradient-magnitude frame as per-CTA strips (Sobel pairs once per position, in shared memory)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02v; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dropin.py tests/test_gpu_metrics.py -q --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -6 $O/pytest.txt | cut -c1-300
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 > $O/bench_16mi.json 2> $O/bench_16mi.err
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 --events 500000 > $O/bench_500k.json 2> $O/bench_500k.err
for c in 1 8; do
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_c$c.json 2> $O/solve_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02v/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],4), d.get("roofline",{}).get("kernel_ms"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
