#!/bin/bash
# round 2, GPU session E (2 GPUs): event-sharded parity test on hardware, default bench at N=2, giant-window A/B of the exchanges
cd "$(dirname "$0")/../.."
O=gpurun_out/r02e; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharding.py -q -s --timeout=600 -p no:cacheprovider > $O/pytest_sharding.txt 2>&1; echo "rc=$?" >> $O/pytest_sharding.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 > $O/bench_default_n2.json 2> $O/bench_default_n2.err; echo "rc=$?" >> $O/bench_default_n2.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --workload giant --steps 30 > $O/giant_n2_oneshot.json 2> $O/giant_n2_oneshot.err
EBOS_P2P_FORM=2 timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload giant --steps 30 > $O/giant_n2_twoshot.json 2> $O/giant_n2_twoshot.err
EBOS_NO_P2P=1 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload giant --steps 30 > $O/giant_n2_nccl.json 2> $O/giant_n2_nccl.err
tail -12 $O/pytest_sharding.txt; tail -3 $O/bench_default_n2.err
python - <<'PY'
import json
for f in ("bench_default_n2","giant_n2_oneshot","giant_n2_twoshot","giant_n2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02e/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("exchange"), json.dumps(d.get("parity_self_check"))[:300])
        for k in ("e2e","solve","giant"):
            if k in d:
                v=d[k]
                if k in ("solve","giant"): v={kk:v[kk] for kk in ("value","ms_per_step","ms_per_window_per_gpu","parity_self_check","exchange") if kk in v}
                if k=="e2e": v={kk:v[kk] for kk in ("value","ms_per_step","h2d_gbs_per_rank") if kk in v}
                print("  ",k, json.dumps(v)[:500])
        print("  numa", d.get("host_numa_binding"))
    except Exception as e: print(f,"ERR",e)
PY
