#!/bin/bash
# round 2, GPU session U: A/B of the Adam-factor change; gradmag with the frame code unrolled again (2528 SASS instructions)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_zz_eklt.py -q --timeout=600 -p no:cacheprovider > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -3 $O/pytest.txt | cut -c1-300
for c in 1 8; do
timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_c$c.json 2> $O/solve_c$c.err
EBOS_ADAM_OWN_COEFS=1 timeout 300 python bench.py --workload solve --no-cpu --solve-concurrency $c > $O/solve_owncoefs_c$c.json 2> $O/solve_owncoefs_c$c.err
done
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 > $O/bench_16mi.json 2> $O/bench_16mi.err
timeout 300 python bench.py --no-e2e --no-cpu --no-subrecords --steps 30 --events 500000 > $O/bench_500k.json 2> $O/bench_500k.err
timeout 300 python bench.py --workload eklt --steps 12 --no-cpu > $O/eklt_pdl.json 2> $O/eklt_pdl.err
EBOS_NO_PDL=1 timeout 300 python bench.py --workload eklt --steps 12 --no-cpu > $O/eklt_nopdl.json 2> $O/eklt_nopdl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02u/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"],2), d["unit"], round(d["ms_per_step"],4), d.get("roofline",{}).get("kernel_ms"), d.get("eval_ms_per_level"), d.get("ms_per_window_single"))
    except Exception as e: print(f,"ERR",e, open(f.replace('.json','.err')).read()[-800:])
PY
