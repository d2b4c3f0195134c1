#!/bin/bash
# round 2, GPU session G2: ncu launch list of the SOLVER iteration (500 k events, one window, replayed graph nodes)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02g2; mkdir -p $O
timeout 900 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv --log-file $O/solve_launches.csv \
  python bench.py --workload solve --solve-concurrency 1 --solve-iters 40 --steps 2 --warmup 1 --no-cpu > $O/solve_launches.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02g2/solve_launches.csv") if l.startswith('"'))]
h=rows[0]; ix={n:i for i,n in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault((r[ix['Kernel Name']][:44], r[ix['Grid Size']]), []).append(float(r[ix['Metric Value']]))
    except Exception: pass
for k,v in agg.items(): print(k, len(v), 'avg us', round(sum(v)/len(v)/1000,2))
PY
tail -2 $O/solve_launches.log | cut -c1-200
