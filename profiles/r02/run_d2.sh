#!/bin/bash
# round 2, GPU session D2: red.global.add.v2.f32 for the two column-adjacent taps of a cell (EBOS_RED_V2 build) -- parity + A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02d2; mkdir -p $O
ALT=$PWD/event_based_bos_b200/libebos_alt.so
EBOS_LIBRARY=$ALT timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py tests/test_gpu_dropin.py -q --timeout=600 -p no:cacheprovider > $O/pytest_alt.txt 2>&1; echo "rc=$?" >> $O/pytest_alt.txt
tail -3 $O/pytest_alt.txt | cut -c1-200
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
for tag in base alt base2 alt2; do
  L=""; case $tag in alt*) L=$ALT;; esac
  EBOS_LIBRARY=$L timeout 300 $B --events 500000 > $O/b500k_$tag.json 2> $O/b500k_$tag.err
  EBOS_LIBRARY=$L timeout 300 $B --events 1048576 > $O/b1m_$tag.json 2> $O/b1m_$tag.err
done
for tag in base alt; do
  L=""; case $tag in alt*) L=$ALT;; esac
  EBOS_LIBRARY=$L timeout 300 python bench.py --workload solve --no-cpu > $O/solve_c8_$tag.json 2> $O/solve_c8_$tag.err
  EBOS_LIBRARY=$L timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency 1 > $O/solve_c1_$tag.json 2> $O/solve_c1_$tag.err
  EBOS_LIBRARY=$L timeout 300 $B > $O/b16m_$tag.json 2> $O/b16m_$tag.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d2/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d.get("roofline",{}).get("kernel_ms") or {}
        print(f.split('/')[-1], round(d["value"],2), round(d["ms_per_step"],4), k.get('window_splat(+memset)'), k.get('window_backward'))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-500:])
PY
