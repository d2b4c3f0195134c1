"""Host-side time of one fused solve, phase by phase (run on a GPU box): where do the ~14 ms per window go that
estimate_many spends outside planning / waiting / copying?"""
import time, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from event_based_bos_b200 import solver, ops
from event_based_bos_b200.utils import smooth_flow, synthetic_bos_events

H, W = 720, 1280
cfg = {"outer_padding": 0, "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": 600},
       "cmax": {"cost_with_weight": {"gradient_magnitude": 1.0, "image_gradient": 0.5}, "lr": 0.05}}
slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, cfg, None)
ev = synthetic_bos_events(500000, (H, W), smooth_flow((H, W), seed=0), seed=1).astype(np.float64)
pin = torch.from_numpy(ev).pin_memory().numpy()
slv.estimate(pin); slv.estimate(pin)
T = lambda: time.perf_counter()
for rep in range(3):
    torch.cuda.synchronize(); t = [T()]
    x0 = slv._upload_flow0(None); t.append(T())
    e = slv._upload_events(pin); t.append(T())
    advance, n = slv._plan_fused(e, x0); t.append(T())
    for _ in range(n): advance()
    t.append(T())
    out = slv._finish(x0); stage = slv._stage(out, 0); stage.copy_(out, non_blocking=True); t.append(T())
    torch.cuda.synchronize(); t.append(T())
    r = stage.numpy().copy(); t.append(T())
    del advance; t.append(T())
    del e, x0, out; t.append(T())
    names = ["flow0", "upload_events", "plan(prepare+capture)", "replays", "finish+d2h queue", "gpu wait", "host copy", "del graph", "del tensors"]
    print(rep, {k: round((b - a) * 1e3, 3) for k, a, b in zip(names, t, t[1:])}, flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
slv.estimate_many([pin] * 16, concurrency=8)
pr.disable()
print(slv.last_many_stats)
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
