#!/usr/bin/env python
"""Benchmark of the contrast-maximisation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fused|solve|giant|eklt]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload "fused" (default, BASELINE config 2 at its largest single-GPU size): one step = one fused warp -> IWE -> cost ->
backward evaluation (gradient-magnitude objective + 0.5*TV) over one window of 16 Mi synthetic events on a 1280x720
grid, fp32.  `value` = events/s with the prepared window resident in HBM (inputs > L2); `e2e` = the same through the
public API from pinned HOST buffers (H2D of the raw sensor stream, ingestion, window preparation, evaluation, D2H of
loss + gradient inside the timed region).  The SAME JSON line carries the other BASELINE configurations as
sub-records, each preceded by an in-process parity self-check whose result is printed in the record:
  "small_windows"  config 2's small end (1 Mi and 0.5 Mi events),
  "solve"          configs 3/4: full per-window flow solves, windows/s, window-sharded over the ranks (no collective),
  "giant"          config 5: ONE 128 Mi-event window sharded by events over the ranks (strong scaling; exchange named).
`--workload solve|giant|eklt` print one of them as the main line instead (eklt = BASELINE config 1, hot_plate1).
N > 1 (torchrun): the default workload runs one 16 Mi window per GPU (weak scaling, no data-path collective).

`--impl reference` times the reference's CPU torch path (the oracle port: same torch ops as src/warp.py +
src/event_image_converter.py + autograd) on the host cores, on the SAME configuration (16 Mi events per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 720, 1280
P_BYTES = H * W * 4
COST = "gradient_magnitude"
TV_WEIGHT = 0.5


def measured_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def soak(self, fn, min_samples: int = 4, max_s: float = 3.0):
        """Keep the GPU under the same load until nvidia-smi has produced a few samples (the timed
        region itself can be shorter than one sampling period)."""
        t0 = time.perf_counter()
        while self.proc is not None and len(self.rows) < min_samples and time.perf_counter() - t0 < max_s:
            for _ in range(20):
                fn()
            torch.cuda.synchronize()

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cuda_time_ms(fn, reps: int):
    """Average ms per call of `fn` over `reps` calls, CUDA events on the current stream."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def graph_time_ms(fn, reps: int):
    """Average device ms per call of `fn`, with `reps` calls captured in ONE CUDA graph so that host launch
    overhead (ctypes + driver, ~10 us per call) does not hide short kernels.  CUDA events around the replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def dist_setup(n_gpus: int):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def bind_to_gpu_numa(local: int):
    """Pin this process to the CPUs closest to its GPU BEFORE any pinned host buffer is allocated (cudaHostAlloc places
    pages by the calling thread's policy: first touch on the local NUMA node): with all ranks on node 0 the round-1
    end-to-end numbers stopped scaling at ~110-180 GB/s aggregate H2D.  NVML's ideal-CPU set for the device (what
    `nvidia-smi topo -m` prints as CPU affinity), sysfs as fallback.  Returns a short description for the JSON line."""
    before = len(os.sched_getaffinity(0))
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        try:
            pynvml.nvmlDeviceSetCpuAffinity(h)
            after = len(os.sched_getaffinity(0))
            if 0 < after <= before:
                return {"source": "nvml ideal cpu set", "cpus": after, "cpus_before": before}
        except Exception:
            pass
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"source": "sysfs numa_node", "numa_node": node, "cpus": len(allowed), "cpus_before": before}
    except Exception:
        return None
    return None


class near_gpu:
    """`with near_gpu(local) as info:` -- allocate pinned host buffers inside; the process's CPU affinity is restored on
    exit (the CPU baseline must see all host cores again)."""

    def __init__(self, local: int):
        self.local, self.saved = local, None

    def __enter__(self):
        self.saved = os.sched_getaffinity(0)
        return bind_to_gpu_numa(self.local)

    def __exit__(self, *a):
        try:
            os.sched_setaffinity(0, self.saved)
        except Exception:
            pass


def all_ranks_equal(t: torch.Tensor, world: int) -> bool:
    """True when `t` holds bit-identical values on every rank (MAX and MIN all-reduce agree element for element)."""
    if world == 1:
        return True
    import torch.distributed as dist

    hi, lo = t.clone(), t.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's torch path (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_reference_fused(n_events: int, steps: int, warmup: int, seed: int = 0):
    from oracle import spec

    torch.set_num_threads(os.cpu_count() or 1)
    ev = torch.from_numpy(spec.synthetic_events(n_events, (H, W), seed=seed))
    flow = torch.from_numpy(spec.synthetic_flow((H, W), seed=seed))
    for _ in range(warmup):
        spec.cmax_value_and_grad(ev, flow, (H, W), cost=COST, tv_weight=TV_WEIGHT)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        spec.cmax_value_and_grad(ev, flow, (H, W), cost=COST, tv_weight=TV_WEIGHT)
        times.append(time.perf_counter() - t0)
    return n_events / float(np.mean(times)), float(np.mean(times)) * 1e3, torch.get_num_threads()


def cpu_reference_solve(n_events: int, iters: int, seed: int = 0):
    from oracle import spec

    torch.set_num_threads(os.cpu_count() or 1)
    ev = torch.from_numpy(spec.synthetic_events(n_events, (H, W), seed=seed))
    t0 = time.perf_counter()
    spec.solve_dense_flow(ev, (H, W), iters, COST, TV_WEIGHT)
    dt = time.perf_counter() - t0
    return dt / iters, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_events or args.events      # the GPU arm's configuration unless a smaller sample is asked for
    if args.workload == "solve":
        s_per_it, cores = cpu_reference_solve(args.solve_events, max(2, min(args.steps, 5)))
        value = 1.0 / (s_per_it * args.solve_iters)
        line = {"metric": "windows/s full per-window flow solve", "value": value, "unit": "windows/s",
                "ms_per_step": s_per_it * args.solve_iters * 1e3,
                "config": {"workload": f"full per-window dense flow solve, {args.solve_events} events, 1280x720, "
                                       f"{COST}+{TV_WEIGHT}*TV, Adam {args.solve_iters} iterations (extrapolated from "
                                       f"timed iterations)"},
                "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "port",
                                 "sample": f"{max(2, min(args.steps, 5))} Adam iterations timed, x{args.solve_iters}"}}
    elif args.workload == "eklt":
        iters = sum(args.solve_iters // (4 + 1 - s + 1) for s in range(1, 5))
        s_eval, cores = cpu_reference_eklt(args.solve_events, evals=max(2, min(args.steps, 5)))
        value = 1.0 / (s_eval * iters)
        line = {"metric": "windows/s hot_plate1 EKLT solve (PatchEkltPyramid2)", "value": value, "unit": "windows/s",
                "ms_per_step": s_eval * iters * 1e3,
                "config": {"workload": f"configs/hot_plate1.yaml pipeline: PatchEkltPyramid2 objective, {args.solve_events} "
                                       f"synthetic events + synthetic frame, 1280x720, {iters} iterations (extrapolated "
                                       f"from timed finest-level evaluations)"},
                "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "port",
                                 "sample": f"{max(2, min(args.steps, 5))} iterations (objective + autograd backward, the "
                                           f"reference's torch ops in float64, finest level) timed at {s_eval:.3f} s "
                                           f"each, x{iters}"}}
    else:
        # a CPU step at 16 Mi events takes ~1 s: the whole --steps/--warmup run stays within a few minutes
        steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
        value, ms, cores = cpu_reference_fused(n_sample, steps, warmup)
        line = {"metric": "events/s fwd+bwd warp->IWE->cost", "value": value, "unit": "events/s", "ms_per_step": ms,
                "config": fused_config(n_sample, None),
                "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_sample} events per step (the GPU arm's window), {steps} steps after {warmup} "
                                           f"warm-up; oracle torch-CPU restatement of the reference ops, fp32, all host threads"}}
        line["config"]["cpu_steps_timed"] = steps
    line.update({"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                 "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                 "dtype": "f64" if args.workload == "eklt" else "f32", "data": "synthetic",
                 "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)


def fused_config(n: int, packed):
    """`config` of the default workload -- identical for both arms (the driver compares them key by key)."""
    cfg = {"workload": f"single-window fused warp->IWE->{COST}+{TV_WEIGHT}*TV forward+backward microbench, "
                       f"{n} synthetic events (uniform, flow U(-3,3)), 1280x720, fp32, atomic mode; one window per GPU",
           "events_per_window": n,
           "l2_policy": f"inputs larger than L2 ({8 * n / 1e6:.0f}-{12 * n / 1e6:.0f} MB sorted SoA per pass)"
           if 8 * n > 126e6 else "inputs fit L2 (warm-L2 number)"}
    if packed is not None:
        cfg["window_layout"] = "packed (row,col,dt) 8 B/event" if packed else "generic (x,y,dt) 12 B/event"
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def small_window_points(args, dev, sizes=(1 << 20, 500_000)):
    """BASELINE config 2's small end: the same fused evaluation on 1 Mi and 0.5 Mi-event windows (CUDA-graph replay;
    the working set fits L2, so these are warm-L2 numbers -- stated in the record)."""
    from event_based_bos_b200 import ops
    from event_based_bos_b200.utils import synthetic_events, synthetic_flow

    peak, _ = measured_peak_gbs()
    flow = torch.from_numpy(synthetic_flow((H, W), seed=0)).to(dev)
    out = []
    for n in sizes:
        ev = torch.from_numpy(synthetic_events(n, (H, W), seed=0)).to(dev)
        win = ops.PreparedWindow(ev, (H, W), "first", True)
        cap = ops.CmaxGraph(win, flow, COST, 1.0, TV_WEIGHT)
        for _ in range(5):
            cap.replay()
        ms_single = cuda_time_ms(cap.replay, max(args.steps, 50))
        # ten evaluations per executable graph, as the solver issues its iterations (ops.ReplaySlot): at this size one
        # graph launch per evaluation costs as much as one of its kernels
        ws = ops.CmaxWorkspace(H, W, (0, 0), dev)
        slot, side = ops.ReplaySlot(), torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.cmax_value_and_grad(win, flow, COST, 1.0, TV_WEIGHT, workspace=ws)
            slot.capture(lambda: [ops.cmax_value_and_grad(win, flow, COST, 1.0, TV_WEIGHT, workspace=ws) for _ in range(10)])
            for _ in range(3):
                slot.launch()
            ms = cuda_time_ms(slot.launch, max(args.steps, 50) // 5) / 10.0
        torch.cuda.current_stream().wait_stream(side)
        alg = 32 * n + 11 * P_BYTES
        out.append({"events": n, "ms_per_step": round(ms, 5), "value": n / (ms * 1e-3), "unit": "events/s",
                    "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "l2_policy": "inputs fit L2 (warm-L2 number)",
                    "launch": "ten evaluations per executable-graph replay", "ms_per_step_one_evaluation_per_replay": round(ms_single, 5)})
        del cap, slot, ws, win, ev
    return out


def run_fused(args, rank, world, local):
    from event_based_bos_b200 import _capi, ops
    from event_based_bos_b200.utils import synthetic_events, synthetic_flow

    n = args.events
    dev = torch.device("cuda", local)
    events_np, flow_np = synthetic_events(n, (H, W), seed=rank), synthetic_flow((H, W), seed=rank)
    with near_gpu(local) as numa:           # pinned pages land on the GPU's NUMA node (first touch)
        ev_host = torch.from_numpy(events_np).pin_memory()
        flow_host = torch.from_numpy(flow_np).pin_memory()
    del events_np, flow_np
    ev = ev_host.to(dev, non_blocking=True)
    flow = flow_host.to(dev, non_blocking=True)
    window = ops.PreparedWindow(ev, (H, W), "first", True, allow_packed=not args.no_packed)
    ws = ops.CmaxWorkspace(H, W, (0, 0), dev)      # clean-workspace protocol: zero between evaluations
    ws_k = ops.CmaxWorkspace(H, W, (0, 0), dev)    # scratch for the per-kernel timings (written directly)
    ws_k.clean = False

    def step():
        ops.cmax_value_and_grad(window, flow, COST, 1.0, TV_WEIGHT, None, False, (0, 0), ws)

    # the timed step: the same evaluation captured once as a CUDA graph (public API ops.CmaxGraph) and replayed
    captured = ops.CmaxGraph(window, flow, COST, 1.0, TV_WEIGHT, None, False, (0, 0), ws)
    launches, other_nodes = ops.count_launches(step)     # counted from a stream capture of the same call
    for _ in range(max(args.warmup, 3)):
        captured.replay()
    barrier(world)
    with ClockSampler(local) as clocks:
        ms = cuda_time_ms(captured.replay, args.steps)
        barrier(world)
        ms = max_over_ranks(ms, world)
        eager_ms = max_over_ranks(cuda_time_ms(step, args.steps), world)
        # the same evaluations issued five per executable graph (how the solvers issue their iterations, ops.ReplaySlot)
        slot5, side5 = ops.ReplaySlot(), torch.cuda.Stream(device=dev)
        side5.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side5):
            slot5.capture(lambda: [step() for _ in range(5)])
            slot5.launch()
            ms5 = cuda_time_ms(slot5.launch, max(1, args.steps // 5)) / 5.0
        torch.cuda.current_stream().wait_stream(side5)
        ms5 = max_over_ranks(ms5, world)
        del slot5
        # per-kernel timing of the kernels of the step (same stream, CUDA events around a graph of `steps` calls)
        lib = _capi.load()
        p = _capi.ptr

        def cur():
            return torch.cuda.current_stream().cuda_stream  # evaluated at call time (graph capture stream)

        def splat_only():
            lib.ebos_window_splat(p(window.buffer), window.n, window.flags, p(flow), H, W, 0, 0, 0, p(ws_k.iwe), cur())

        def bwd_only():
            lib.ebos_window_backward(p(window.buffer), window.n, window.flags, p(flow), H, W, 0, 0, 0, p(ws_k.grad_iwe),
                                     _capi.COST_GRADMAG, p(ws_k.iwe), p(ws_k.acc), 0, 1.0, p(ws_k.dflow), cur())

        def cost_only():
            lib.ebos_iwe_cost(_capi.COST_GRADMAG, p(ws_k.iwe), H, W, 0, 1.0, 0, p(ws_k.acc), p(ws_k.grad_iwe), cur())

        def tv_only():
            lib.ebos_flow_tv(p(flow), 0, H, W, TV_WEIGHT, 0, p(ws_k.acc), p(ws_k.dflow), cur())

        adam_m, adam_v = torch.zeros_like(flow), torch.zeros_like(flow)
        adam_p = flow.clone()

        def adam_only():
            lib.ebos_adam_step(p(adam_p), p(ws_k.dflow), p(adam_m), p(adam_v), adam_p.numel(), 0.05, 0.9, 0.999, 1e-8, 3, 0, cur())

        def var_only():
            lib.ebos_iwe_cost(_capi.COST_VARIANCE, p(ws_k.iwe), H, W, 0, 1.0, 0, p(ws_k.acc), 0, cur())

        k_ms = {name: graph_time_ms(fn, args.steps) for name, fn in
                (("window_splat(+memset)", splat_only), ("window_backward", bwd_only), ("iwe_cost_gradmag", cost_only),
                 ("flow_tv", tv_only), ("iwe_cost_variance", var_only), ("adam_step", adam_only))}
        clocks.soak(captured.replay)
    clk = clocks.summary()

    # end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = e2e_fused(args, rank, world, local, dev, ev_host, flow_host, ws, lib)
    small = None if args.no_subrecords else small_window_points(args, dev)
    packed = bool(window.packed)
    del captured, window, ev
    torch.cuda.empty_cache()
    solve_rec = giant_rec = eklt_rec = None
    if not args.no_subrecords:
        solve_rec = solve_record(args, rank, world, local, quick=True)
        torch.cuda.empty_cache()
        giant_rec = giant_record(args, rank, world, local)
        torch.cuda.empty_cache()
        eklt_rec = eklt_record(args, rank, world, local, quick=True)
        torch.cuda.empty_cache()
    if rank != 0:
        return
    peak, peak_kind = measured_peak_gbs()
    alg_bytes_step = 32 * n + 11 * P_BYTES  # 32N + 8P (fwd+cost+bwd) + 3P (TV: flow read 2P + weights P)
    splat_bytes = 16 * n + 3 * P_BYTES      # events + flow read (2P) + IWE write (P)
    bwd_bytes = 16 * n + 3 * P_BYTES        # events + dIWE read (P) + dflow write (2P)
    dom = max(("window_splat(+memset)", "window_backward"), key=lambda k: k_ms[k])
    dom_bytes = splat_bytes if dom.startswith("window_splat") else bwd_bytes
    achieved = dom_bytes / (k_ms[dom] * 1e-3) / 1e9
    line = {
        "metric": "events/s fwd+bwd warp->IWE->cost", "value": world * n / (ms * 1e-3), "unit": "events/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": fused_config(n, None),
        "window_layout": "packed (row,col,dt) 8 B/event" if packed else "generic (x,y,dt) 12 B/event",
        "clocks": clk,
        "step_roofline": {"algorithmic_bytes": alg_bytes_step, "achieved_gbs": alg_bytes_step / (ms * 1e-3) / 1e9,
                          "frac": alg_bytes_step / (ms * 1e-3) / 1e9 / peak},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": measured_traffic(dom) if n == (1 << 24) and packed else None,
                     "peak_source": peak_kind, "algorithmic_bytes": dom_bytes,
                     "kernel_ms": {k: round(v, 4) for k, v in k_ms.items()}},
        # kernels of this library enqueued per step, COUNTED from a stream capture of the step (ops.count_launches),
        # times the timed steps; memset nodes listed separately
        "gpu_launches": launches * args.steps,
        "launches_per_step": {"kernels": launches, "memset_nodes": other_nodes},
        "ms_per_step_eager": round(eager_ms, 4), "ms_per_step_five_evaluations_per_replay": round(ms5, 4),
        "launch": "one CUDA-graph replay per step (ops.CmaxGraph); ms_per_step_eager = the same launches issued eagerly",
        "host_numa_binding": numa,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if small is not None:
        line["small_windows"] = small
    if solve_rec is not None:
        line["solve"] = solve_rec
    if giant_rec is not None:
        line["giant"] = giant_rec
    if eklt_rec is not None:
        line["eklt"] = eklt_rec
    if not args.no_cpu and world == 1:   # the CPU baseline is timed on rank 0 at N=1 only
        n_cpu = args.cpu_events or n
        v, cms, cores = cpu_reference_fused(n_cpu, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} events (the same window size), mean of 3 evaluations after 1 warm-up "
                                          f"(oracle torch-CPU restatement of the reference ops, fp32, all host threads)"}
    print(json.dumps(line), flush=True)


def e2e_fused(args, rank, world, local, dev, ev_host, flow_host, ws, lib):
    """The same evaluation end to end from pinned host buffers, software-pipelined over two streams (the H2D copy of step
    i+1 overlaps ingestion + preparation + evaluation + D2H of step i; every step's copies are inside the timed region).
    Primary: the RAW sensor stream (x,y int16; t int32 us; p bool = 9 B/event, what the camera delivers and
    `ebos_ingest_raw` consumes).  Secondary: the reference loader's [N,4] fp32 rows (16 B/event)."""
    from event_based_bos_b200 import _capi as capi
    from event_based_bos_b200 import ops

    n = ev_host.shape[0]
    p = capi.ptr
    evn = ev_host.numpy()
    raw_np = [np.ascontiguousarray(a) for a in (evn[:, 1].astype(np.int16), evn[:, 0].astype(np.int16),
                                                np.round(evn[:, 2].astype(np.float64) * 1e6).astype(np.int32),
                                                evn[:, 3].astype(np.uint8))]
    with near_gpu(local):
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        grad_host = torch.zeros((2, H, W), dtype=torch.float32).pin_memory()
        raw_host = [torch.from_numpy(a).pin_memory() for a in raw_np]
    del raw_np
    copy_stream = torch.cuda.Stream(device=dev)
    fl_bufs = [torch.empty_like(flow_host, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    rows_dev = torch.empty((n, 4), dtype=torch.float32, device=dev)
    cnt_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    t0_us = int(raw_host[2][0])

    def cur():
        return torch.cuda.current_stream().cuda_stream

    def evaluate(rows, fl):
        win = ops.PreparedWindow(rows, (H, W), "first", True, validate=False, allow_packed=not args.no_packed)
        loss, grad = ops.cmax_value_and_grad(win, fl, COST, 1.0, TV_WEIGHT, None, False, (0, 0), ws)
        loss_host.copy_(loss, non_blocking=True)
        grad_host.copy_(grad, non_blocking=True)

    def pipelined(reps, bufs, hosts, consume):
        main = torch.cuda.current_stream()
        for b in range(2):
            free[b].record(main)

        def upload(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])           # the step that used this pair of buffers is done
                for dst, src in zip(bufs[b], hosts):
                    dst.copy_(src, non_blocking=True)
                fl_bufs[b].copy_(flow_host, non_blocking=True)
                ready[b].record(copy_stream)

        upload(0)
        for i in range(reps):
            b = i % 2
            if i + 1 < reps:
                upload(i + 1)
            main.wait_event(ready[b])
            consume(bufs[b], fl_bufs[b])
            free[b].record(main)

    def consume_raw(bufs, fl):
        x, y, t, pp = bufs
        capi.check(lib.ebos_ingest_raw(p(x), p(y), p(t), p(pp), n, 0, 0, 0, 0, 0, t0_us, 1, 0, p(rows_dev), p(cnt_dev),
                                       0, 0, cur()), "ebos_ingest_raw")
        evaluate(rows_dev, fl)

    reps = max(6, min(args.steps, 12))
    out = {}
    raw_bufs = [[torch.empty_like(a, device=dev) for a in raw_host] for _ in range(2)]
    pipelined(4, raw_bufs, raw_host, consume_raw)
    barrier(world)
    raw_ms = max_over_ranks(cuda_time_ms(lambda: pipelined(reps, raw_bufs, raw_host, consume_raw), 1) / reps, world)
    serial_ms = max_over_ranks(cuda_time_ms(lambda: pipelined(1, raw_bufs, raw_host, consume_raw), max(3, min(args.steps, 6))), world)
    del raw_bufs
    row_bufs = [[torch.empty_like(ev_host, device=dev)] for _ in range(2)]
    pipelined(4, row_bufs, [ev_host], lambda bufs, fl: evaluate(bufs[0], fl))
    barrier(world)
    rows_ms = max_over_ranks(cuda_time_ms(lambda: pipelined(reps, row_bufs, [ev_host], lambda bufs, fl: evaluate(bufs[0], fl)), 1) / reps, world)
    h2d_raw, h2d_rows, d2h = 9 * n + 2 * P_BYTES, 16 * n + 2 * P_BYTES, 2 * P_BYTES + 4
    out = {"value": world * n / (raw_ms * 1e-3), "unit": "events/s", "ms_per_step": raw_ms,
           "h2d_bytes_per_step": h2d_raw, "d2h_bytes_per_step": d2h,
           "h2d_gbs_per_rank": h2d_raw / (raw_ms * 1e-3) / 1e9,
           "includes": "H2D of the raw sensor stream (int16 x,y; int32 t; bool p = 9 B/event) + flow from pinned host memory, "
                       "ingestion (ebos_ingest_raw), window preparation (sort), fused evaluation, D2H loss + gradient; "
                       "software-pipelined over two streams (the H2D of step i+1 overlaps the compute of step i)",
           "ms_per_step_serial": serial_ms,
           "rows_fp32": {"value": world * n / (rows_ms * 1e-3), "unit": "events/s", "ms_per_step": rows_ms,
                         "h2d_bytes_per_step": h2d_rows, "h2d_gbs_per_rank": h2d_rows / (rows_ms * 1e-3) / 1e9,
                         "includes": "same, fed with the reference loader's [N,4] fp32 event rows (16 B/event)"}}
    return out


def solve_record(args, rank, world, local, quick=False):
    """BASELINE configs 3/4: full per-window dense flow solves (host events in -> host flow out through the public
    solver API), independent windows sharded over the ranks with NO data-path collective (weak scaling: every rank
    solves its own windows).  Returns the record (rank 0) or None."""
    from event_based_bos_b200 import ops, solver
    from event_based_bos_b200.utils import smooth_flow, synthetic_bos_events, synthetic_events

    n, iters = args.solve_events, args.solve_iters
    dev = torch.device("cuda", local)
    cfg = {"outer_padding": 0, "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": iters},
           "cmax": {"cost_with_weight": {COST: 1.0, "image_gradient": TV_WEIGHT}, "lr": 0.05,
                    "fold_tv": args.fold_tv}}
    slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, cfg, None)
    gt = smooth_flow((H, W), seed=rank)
    conc = max(1, args.solve_concurrency)
    # parity self-check (window sharding has no numerical coupling between ranks: the check is that a rank's solve of
    # a window equals rank 0's solve of the same window): every rank solves the SAME probe window from the same
    # tie-free start; RMS distance of the final flows to rank 0's, bar 1e-3 px (north_star)
    probe_cfg = {"outer_padding": 0, "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": 40},
                 "cmax": {"cost_with_weight": {COST: 1.0, "image_gradient": TV_WEIGHT}, "lr": 0.05}}
    probe_slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, probe_cfg, None)
    probe_ev = synthetic_bos_events(200_000, (H, W), smooth_flow((H, W), seed=99), seed=4242).astype(np.float64)
    flow0 = np.random.default_rng(7).uniform(-0.5, 0.5, (2, H, W))
    mine = torch.from_numpy(probe_slv.estimate(probe_ev, flow0=flow0)).to(dev)
    ref0 = mine.clone()
    if world > 1:
        import torch.distributed as dist

        dist.broadcast(ref0, src=0)
    rms = torch.sqrt(torch.mean((mine - ref0) ** 2)).reshape(1)
    if world > 1:
        dist.all_reduce(rms, op=dist.ReduceOp.MAX)
    parity = {"check": "every rank solves the same 200 k-event probe window (40 Adam iterations, tie-free start); "
                       "max over ranks of the RMS distance to rank 0's flow", "rms_px": float(rms), "bar_px": 1e-3,
              "ok": bool(float(rms) <= 1e-3)}
    # launches of ONE solver iteration, counted from a stream capture of the public one-call iteration
    tiny = ops.PreparedWindow(torch.from_numpy(synthetic_events(4096, (H, W), seed=1)).to(dev), (H, W), "first", True)
    f_t = torch.zeros((2, H, W), device=dev)
    m_t, v_t, st_t = torch.zeros_like(f_t), torch.zeros_like(f_t), torch.zeros(1, dtype=torch.int32, device=dev)
    ws_t = ops.CmaxWorkspace(H, W, (0, 0), dev)
    if slv.fold_tv and ops.fused_tv_supported(tiny) and iters % 2 == 0:     # the iteration the solver issues (TV inside Adam)
        f_o = torch.zeros_like(f_t)
        ws_t.zero_dflow()
        it_launches, it_other = ops.count_launches(
            lambda: ops.cmax_adam_iteration_fused_tv(tiny, f_t, f_o, m_t, v_t, st_t, ws_t, COST, 1.0, TV_WEIGHT))
    else:
        it_launches, it_other = ops.count_launches(lambda: ops.cmax_adam_iteration(tiny, f_t, m_t, v_t, st_t, ws_t, COST, 1.0, TV_WEIGHT))
    del tiny, f_t, m_t, v_t, ws_t
    # pinned host buffers (the e2e contract: inputs come from pinned host memory), handed over as numpy views
    host_events = [synthetic_bos_events(n, (H, W), gt, seed=1000 * rank + i).astype(np.float64) for i in range(max(2, conc))]
    with near_gpu(local):    # pinned inputs and the solver's pinned staging buffers land on the GPU's NUMA node
        pinned = [torch.from_numpy(a).pin_memory() for a in host_events]
        windows = [t.numpy() for t in pinned]
        for _ in range(1 if quick else max(1, min(args.warmup, 2))):
            slv.estimate(windows[0])
        if conc > 1:
            slv.estimate_many(windows[:conc] * 2, concurrency=conc)   # warm-up of every stream slot (staging buffers)
    del host_events
    # (sub-record of the default line: 6 rounds of the slots, ~1 s -- the start-up of the rolling schedule, ~4 ms of host
    #  planning before the first graph replay, must not weigh on a steady-state rate)
    n_solves = 6 * conc if quick else max(args.steps, 6 * conc)
    batch = [windows[i % len(windows)] for i in range(n_solves)]
    barrier(world)
    with ClockSampler(local) as clocks:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if conc > 1:
            slv.estimate_many(batch, concurrency=conc)  # independent windows, `conc` solves in flight (public API)
        else:
            for w in batch:
                slv.estimate(w)                         # host events in, host flow out: this IS the public API
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n_solves
        barrier(world)
        ms = max_over_ranks(ms, world)
        clocks.soak(lambda: None, max_s=0.5)
    if rank != 0:
        return None
    peak, peak_kind = measured_peak_gbs()
    alg = iters * (32 * n + 25 * P_BYTES)
    rec = {"metric": "windows/s full per-window flow solve", "value": world / (ms * 1e-3), "unit": "windows/s",
           "n_gpus": world, "ms_per_window_per_gpu": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"full per-window dense flow solve, {n} BOS-like events, 1280x720, {COST}+{TV_WEIGHT}*TV, "
                                  f"Adam lr 0.05, K = {iters} iterations, zero init; host events in -> host flow out; "
                                  f"{conc} independent windows in flight per GPU; windows sharded over the ranks, no collective",
                      "iterations": iters, "windows_timed_per_gpu": n_solves, "concurrency": conc},
           "host_ms_per_window": {k: round(v, 3) for k, v in slv.last_many_stats.items()},
           "parity_self_check": parity,
           "clocks": clocks.summary(),
           "roofline": {"bound": "hbm", "kernel": "whole solve (all kernels)", "achieved": alg / (ms * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                        "peak_source": peak_kind, "algorithmic_bytes_per_window": alg},
           "e2e": {"value": world / (ms * 1e-3), "unit": "windows/s", "h2d_bytes_per_step": 32 * n,
                   "d2h_bytes_per_step": 4 * P_BYTES},
           "gpu_launches": n_solves * iters * it_launches,
           "launches_per_iteration": {"kernels": it_launches, "memset_nodes": it_other}}
    if not args.no_cpu and world == 1:
        s_per_it, cores = cpu_reference_solve(n, 3)
        rec["cpu_baseline"] = {"value": 1.0 / (s_per_it * iters), "unit": "windows/s", "cores": cores, "kind": "port",
                               "sample": f"3 Adam iterations of the oracle loop timed ({s_per_it:.3f} s/it), x{iters}"}
    return rec


def run_solve(args, rank, world, local):
    rec = solve_record(args, rank, world, local)
    if rank != 0:
        return
    rec.update({"steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_window_per_gpu"], "vs_baseline": None})
    print(json.dumps(rec), flush=True)


# ------------------------------------------------------------------------------------------------
# hot_plate1 pipeline (BASELINE config 1): PatchEkltPyramid2, the EKLT inner loop
HOT_PLATE1_SOLVER = {
    "filter": {"filters": None, "parameters": {"xmin": 0, "xmax": 720, "ymin": 320, "ymax": 960}},
    "method": "patch_eklt_pyramid2", "outer_padding": 0,
    "cost_with_weight": {"diff_norm": 1.0, "image_gradient": 0.5, "flow_norm_pxy": 0.1},
    "optimizer": {"method": "Adam", "n_iter": 600, "parameters": {}},
    "generative_ml": {"weight_loss_by_event_hist": False, "weight_sigma": 5, "weight_loss_by_inverse_event_hist": True,
                      "optimize_warp": True, "iwe_sigma": 2, "no_polarity": False, "model_image": "current",
                      "use_log_intensity": False, "poisson_model": True},
    "patch_eklt": {"patch_size": 4, "sliding_window": 2, "do_event_thresholding": False, "event_thres": 8},
}


def eklt_inputs(n_events: int, seed: int = 0):
    """events over the full sensor, synthetic uint8 frame = smooth texture + noise."""
    rng = np.random.default_rng(seed)
    ev = np.stack([rng.integers(0, H, n_events), rng.integers(0, W, n_events),
                   np.sort(rng.uniform(0, 1.0 / 120.0, n_events)), rng.integers(0, 2, n_events)], axis=1).astype(np.float64)
    yy, xx = np.mgrid[0:H, 0:W]
    frame = np.clip(120 + 60 * np.sin(xx / 11.0) * np.cos(yy / 7.0) + rng.normal(0, 4, (H, W)), 0, 255).astype(np.uint8)
    return ev, frame


def cpu_reference_eklt(n_events: int, evals: int = 2):
    """What the reference executes per iteration: the level objective with the reference's torch ops (float64) +
    autograd backward (oracle/spec_eklt_torch.py, pinned to the reference's outputs), at the finest level, on all host
    threads.  Returns (seconds per iteration, threads)."""
    from oracle import spec_eklt as E
    from oracle import spec_eklt_torch as TT

    torch.set_num_threads(os.cpu_count() or 1)
    ev, frame = eklt_inputs(n_events)
    gx, gy = E.frame_gradients(frame)
    roi = (0, 720, 320, 960)
    meas, winv, _ = E.measurement_and_weights(ev, (H, W), roi)
    patch, ph, pw = E.pyramid_levels((H, W))[-1]
    th = torch.from_numpy(np.concatenate([np.random.default_rng(1).uniform(-1, 1, (1, ph, pw)), np.zeros((2, ph, pw))]))
    planes = [torch.from_numpy(np.ascontiguousarray(a)).double() for a in (gx, gy, meas, winv)]
    TT.value_and_grad(th, *planes, roi, patch)
    t0 = time.perf_counter()
    for _ in range(evals):
        TT.value_and_grad(th, *planes, roi, patch)
    return (time.perf_counter() - t0) / evals, torch.get_num_threads()


def run_eklt(args, rank, world, local):
    line = eklt_record(args, rank, world, local)
    if rank == 0:
        print(json.dumps(line), flush=True)


def eklt_record(args, rank, world, local, quick=False):
    """One step = one complete PatchEkltPyramid2.estimate (host events + frame in, host flow out): 4 pyramid levels,
    120 + 150 + 200 + 300 = 770 objective/gradient/Adam iterations at n_iter = 600, float64 like the reference.
    `quick` (sub-record of the default line): no A/B timings of the alternative kernel chains, no CPU baseline."""
    from event_based_bos_b200 import eklt, solver

    cfg = json.loads(json.dumps(HOT_PLATE1_SOLVER))
    cfg["optimizer"]["n_iter"] = args.solve_iters
    cfg["eklt"] = {"precision": args.eklt_precision, "cuda_graph": not args.eklt_no_graph}
    slv = solver.collections["patch_eklt_pyramid2"]((H, W), (720, 640), {}, cfg, None)
    ev, frame = eklt_inputs(args.solve_events, seed=rank)
    conc = max(1, args.eklt_concurrency)
    for _ in range(max(1, min(args.warmup, 2))):
        slv.estimate(ev, frame=frame)
    if conc > 1:
        slv.estimate_many([ev] * conc, frames=[frame] * conc, concurrency=conc)      # warm-up of every slot
    barrier(world)
    # latency of ONE estimate() (nothing else on the GPU), then the throughput with `conc` windows in flight
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(max(2, args.steps // 4)):
        slv.estimate(ev, frame=frame)
    b.record()
    torch.cuda.synchronize()
    ms_single = a.elapsed_time(b) / max(2, args.steps // 4)
    n_windows = max(args.steps, 4 * conc) if conc > 1 else args.steps
    barrier(world)
    with ClockSampler(local) as clocks:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if conc > 1:
            slv.estimate_many([ev] * n_windows, frames=[frame] * n_windows, concurrency=conc)
        else:
            for _ in range(n_windows):
                slv.estimate(ev, frame=frame)
        b.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(a.elapsed_time(b) / n_windows, world)
        barrier(world)
        clocks.soak(lambda: None, max_s=0.5)
    # device time of one evaluation per level (value + gradient, no Adam), 20 evaluations per CUDA graph
    dt = torch.float32 if args.eklt_precision == "32" else torch.float64
    prob = eklt.EkltProblem(slv._gradient_x_torch, slv._gradient_y_torch, slv.cache_measured, slv.weight_inverse,
                            (0, 720, 320, 960), (1.0, 0.5, 0.1))
    per_level, per_level_legacy, per_level_stored, per_level_seg = {}, {}, {}, {}
    launches_per_level = {}
    from event_based_bos_b200 import ops as _ops
    for patch, ph, pw in slv.levels:
        lvl = prob.level(patch)
        th = slv.best_params_per_scale[slv.levels.index((patch, ph, pw)) + 1].to(dt).contiguous()
        per_level[patch] = graph_time_ms(lambda: lvl.value_and_grad(th), 20)
        # kernels / memset nodes of ONE solver iteration at this level, counted from a stream capture (nothing runs)
        th_c, m_c, v_c = th.clone(), torch.zeros_like(th), torch.zeros_like(th)
        st_c = torch.zeros(1, dtype=torch.int32, device=th.device)
        launches_per_level[patch] = _ops.count_launches(lambda: lvl.adam_iteration(th_c, m_c, v_c, st_c, 0.05))
        if quick:
            continue
        os.environ["EBOS_EKLT_LEGACY"] = "1"          # the first chain (whole-image TV kernel, per-cell 2-D gather)
        try:
            per_level_legacy[patch] = graph_time_ms(lambda: lvl.value_and_grad(th), 20)
        finally:
            os.environ.pop("EBOS_EKLT_LEGACY", None)
        for sw, store in (("EBOS_EKLT_GATHER_SEG", per_level_seg), ("EBOS_EKLT_STORED", per_level_stored)):
            os.environ[sw] = "0"                      # A/B: the pre-round-2 kernels (warp-per-cell gather, re-evaluating backward)
            try:
                store[patch] = graph_time_ms(lambda: lvl.value_and_grad(th), 20)
            finally:
                os.environ.pop(sw, None)
    if rank != 0:
        return None
    iters = [args.solve_iters // (len(slv.levels) + 1 - s + 1) for s in range(1, len(slv.levels) + 1)]
    elem = 4 if args.eklt_precision == "32" else 8
    plane = H * W * elem
    alg_eval = 25 * plane                                    # csrc/ebos_eklt.cu header: 25 plane passes per evaluation
    alg = sum(iters) * (alg_eval + 14 * 0)                   # Adam on [3,ph,pw] is negligible
    peak, peak_kind = measured_peak_gbs()
    worst = max(per_level, key=per_level.get)
    line = {"metric": "windows/s hot_plate1 EKLT solve (PatchEkltPyramid2)", "value": world / (ms * 1e-3),
            "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if elem == 4 else "f64", "data": "synthetic",
            "config": {"workload": f"configs/hot_plate1.yaml pipeline: PatchEkltPyramid2.estimate, {args.solve_events} "
                                   f"synthetic events + synthetic frame, 1280x720, ROI [0:720,320:960], n_iter "
                                   f"{args.solve_iters} -> {sum(iters)} iterations over 4 levels; host events + frame in "
                                   f"-> host flow out; {conc} independent windows in flight (estimate_many)",
                       "iterations_per_level": iters, "windows_timed_per_gpu": n_windows, "concurrency": conc,
                       "l2_policy": "working set 25 planes x 7.4 MB (fp64) > L2"},
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": f"one objective evaluation at patch {worst} (all kernels of the chain)",
                         "achieved": alg_eval / (per_level[worst] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_eval / (per_level[worst] * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_kind},
            "ms_per_window_single": ms_single, "host_ms_per_window": {k: round(v, 3) for k, v in slv.last_many_stats.items()},
            "eval_ms_per_level": per_level, "eval_ms_per_level_legacy_chain": per_level_legacy or None,
            "eval_ms_per_level_without_stored_planes": per_level_stored or None,
            "eval_ms_per_level_without_segment_gather": per_level_seg or None,
            "e2e": {"value": world / (ms * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": int(ev.nbytes + frame.nbytes), "d2h_bytes_per_step": 2 * H * W * 8},
            "gpu_launches": n_windows * sum(it * launches_per_level[patch][0] for it, (patch, _, _) in zip(iters, slv.levels)),
            "launches_per_iteration_per_level": {str(k): {"kernels": v[0], "memset_nodes": v[1]} for k, v in launches_per_level.items()}}
    if not args.no_cpu and world == 1 and not quick:
        s_eval, cores = cpu_reference_eklt(args.solve_events)
        line["cpu_baseline"] = {"value": 1.0 / (s_eval * sum(iters)), "unit": "windows/s", "cores": cores,
                                "kind": "port", "sample": f"2 iterations (objective + autograd backward with the "
                                                          f"reference's torch ops, float64, finest level) at "
                                                          f"{s_eval:.3f} s each, x{sum(iters)}"}
    return line


def giant_parity_check(rank, world, local):
    """In-process self-check of the event-sharded path on a 2 Mi-event probe window: the sharded loss / gradient
    against the single-GPU fused evaluation of the same window (bar 1e-5, north_star atomic-mode tolerance), and
    bit-identity of the gradient across ranks (every rank must apply the same Adam step)."""
    from event_based_bos_b200 import ops, sharding
    from event_based_bos_b200.utils import synthetic_events, synthetic_flow

    dev = torch.device("cuda", local)
    n = 1 << 21
    ev = torch.from_numpy(synthetic_events(n, (H, W), seed=77)).to(dev)
    flow = torch.from_numpy(synthetic_flow((H, W), seed=77)).to(dev)
    s0, s1 = sharding.shard_events(n, rank, world)
    obj = sharding.cuda_event_sharded_objective(ev[s0:s1].contiguous(), (H, W), cost=COST, tv_weight=TV_WEIGHT)
    loss, grad = obj.value_and_grad(flow)
    loss, grad = loss.clone(), grad.clone()
    same = all_ranks_equal(grad, world) and all_ranks_equal(loss, world)
    out = {"check": "2 Mi-event probe window: event-sharded loss / gradient vs the single-GPU fused evaluation; gradient "
                    "bit-identical on every rank", "ranks_bit_identical": bool(same), "bar": 1e-5}
    if rank == 0:
        win = ops.PreparedWindow(ev, (H, W), "first", True)
        l1, g1 = ops.cmax_value_and_grad(win, flow, COST, 1.0, TV_WEIGHT)
        out["loss_rel_err"] = abs(float(loss) - float(l1)) / abs(float(l1))
        out["grad_rel_err"] = rel_max(grad, g1)
        out["ok"] = bool(same and out["loss_rel_err"] <= 1e-5 and out["grad_rel_err"] <= 1e-5)
    if hasattr(obj, "close"):
        obj.close()
    del obj
    import gc
    gc.collect()
    barrier(world)
    return out


def giant_record(args, rank, world, local):
    """BASELINE config 5: ONE window of `--giant-events` events sharded by events over the ranks (strong scaling):
    partial IWE -> exchange 1 -> cost (redundant) -> partial dflow -> exchange 2 (sharding.py).  At N = 1 the same
    window runs on one GPU without exchange: the reference point of the strong-scaling curve."""
    from event_based_bos_b200 import sharding
    from event_based_bos_b200.utils import synthetic_flow

    parity = giant_parity_check(rank, world, local)
    n_total = args.giant_events
    s0, s1 = sharding.shard_events(n_total, rank, world)
    n = s1 - s0
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    ev = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ev[:, 0] = torch.randint(0, H, (n,), device=dev, generator=gen).float()
    ev[:, 1] = torch.randint(0, W, (n,), device=dev, generator=gen).float()
    ev[:, 2] = torch.sort(torch.rand(n, device=dev, generator=gen) * (1.0 / 120.0))[0]
    ev[:, 3] = torch.randint(0, 2, (n,), device=dev, generator=gen).float()
    flow = torch.from_numpy(synthetic_flow((H, W), seed=0)).to(dev)
    obj = sharding.cuda_event_sharded_objective(ev, (H, W), cost=COST, tv_weight=TV_WEIGHT)
    del ev

    def step():
        obj.value_and_grad(flow)

    steps = max(args.steps, 10)
    steps += steps % 2
    for _ in range(max(args.warmup, 3)):
        step()
    # the evaluation as ONE executable graph (two evaluations per replay: the two-shot form alternates its gradient planes)
    replay = None
    if world > 1 and not args.giant_eager and hasattr(obj, "replayable"):
        replay = obj.replayable(flow, 2)
    launch_mode = "eager launches from Python"
    if replay is not None:
        launch_mode = "two evaluations per executable-graph replay (ops.ReplaySlot)"
        for _ in range(2):
            replay()
    barrier(world)
    with ClockSampler(local) as clocks:
        if replay is not None:
            ms = max_over_ranks(cuda_time_ms(replay, steps // 2) / 2.0, world)
        else:
            ms = max_over_ranks(cuda_time_ms(step, steps), world)
        barrier(world)
        # clock samples under load: a FIXED number of extra steps (the step holds collectives, so a time-based
        # soak would let the ranks run different counts and dead-lock)
        for _ in range(50):
            if replay is not None:
                replay()
            else:
                step()
                step()
        torch.cuda.synchronize()
        barrier(world)
    exchange = getattr(obj, "exchange", "none")
    autotune = getattr(obj, "exchange_autotune", None)
    launches = getattr(obj, "launches_per_evaluation", None)
    # release the symmetric-memory planes deterministically, on every rank, before anything else captures a stream
    del replay
    if hasattr(obj, "close"):
        obj.close()
    del obj
    import gc
    gc.collect()
    torch.cuda.synchronize()
    barrier(world)
    if rank != 0:
        return None
    peak, peak_kind = measured_peak_gbs()
    alg = 32 * n_total + world * 11 * P_BYTES   # the plane work is replicated on every rank
    return {"metric": "events/s fwd+bwd warp->IWE->cost (one event-sharded window)", "value": n_total / (ms * 1e-3),
            "unit": "events/s", "n_gpus": world, "ms_per_step": ms, "steps": steps, "higher_is_better": True,
            "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"single giant window of {n_total} synthetic events sharded by events over {world} GPU(s), "
                                   f"1280x720, fp32, {COST}+{TV_WEIGHT}*TV forward+backward",
                       "events_total": n_total, "events_per_gpu": n, "l2_policy": "inputs larger than L2"},
            "exchange": exchange, "launch": launch_mode,
            "exchange_start_up_timing_ms": autotune,
            "exchange_bytes_per_evaluation_per_rank": 0 if world == 1 else 3 * P_BYTES,   # partial IWE (P) + partial dflow (2P)
            "parity_self_check": parity,
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": "whole evaluation (all kernels + exchanges)",
                         "achieved": alg / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / world / peak, "traffic": None, "peak_source": peak_kind},
            "gpu_launches": (launches or 7) * steps}


def run_giant(args, rank, world, local):
    rec = giant_record(args, rank, world, local)
    if rank != 0:
        return
    rec.update({"warmup": max(args.warmup, 3), "vs_baseline": None})
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fused", choices=["fused", "solve", "giant", "eklt"])
    ap.add_argument("--giant-events", type=int, default=1 << 27)
    ap.add_argument("--events", type=int, default=1 << 24)
    ap.add_argument("--cpu-events", type=int, default=0, help="events per CPU step (0 = the GPU arm's window size)")
    ap.add_argument("--no-subrecords", action="store_true", help="default workload: skip small_windows / solve / giant")
    ap.add_argument("--solve-events", type=int, default=500000)
    ap.add_argument("--solve-iters", type=int, default=600)
    ap.add_argument("--solve-concurrency", type=int, default=8, help="independent windows in flight per GPU (solve workload)")
    ap.add_argument("--eklt-precision", default="64", choices=["32", "64"], help="dtype of the eklt workload (reference: 64)")
    ap.add_argument("--fold-tv", action="store_true", help="solve workload: the fused Adam + TV kernel instead of TV kernel + Adam kernel (A/B)")
    ap.add_argument("--eklt-concurrency", type=int, default=4, help="independent windows in flight per GPU (eklt workload)")
    ap.add_argument("--eklt-no-graph", action="store_true", help="eager launches in the eklt workload (for ncu launch lists)")
    ap.add_argument("--giant-eager", action="store_true", help="giant workload: eager launches instead of the captured evaluation (A/B)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="force the generic 12 B/event window layout")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = dist_setup(args.gpus)
    try:
        {"solve": run_solve, "giant": run_giant, "eklt": run_eklt}.get(args.workload, run_fused)(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
