#!/bin/bash
# round 2, GPU session Z: occupancy variants of the backward (EBOS_BOCC) and of the dense splat (EBOS_QOCC) after this round's changes
cd "$(dirname "$0")/../.."
O=gpurun_out/r02z; mkdir -p $O
B="python bench.py --no-e2e --no-cpu --no-subrecords --steps 30"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; }
run base A=1
run bocc3 EBOS_BOCC=3
run bocc5 EBOS_BOCC=5
run bocc6 EBOS_BOCC=6
run qocc5 EBOS_QOCC=5
run qocc6 EBOS_QOCC=6
run base2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02z/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernel_ms"]
        print(f.split('/')[-1], round(d["ms_per_step"],4), d["roofline"]["frac"], k['window_splat(+memset)'], k['window_backward'])
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
