"""How far from its bar does the solve inside __graft_entry__.smoke() land, run to run (the atomic order differs)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from event_based_bos_b200 import solver
from oracle import spec

H, W = 96, 128
ev_s = torch.from_numpy(spec.synthetic_events(50000, (H, W), seed=1))
flow0 = torch.from_numpy(spec.synthetic_flow((H, W), seed=2, max_val=0.5))
for iters in (10, 20):
    ref = spec.solve_dense_flow(ev_s, (H, W), iters, tv_weight=0.5, flow0=flow0).numpy().astype(np.float64)
    cfg = {"outer_padding": 0, "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": iters},
           "cmax": {"cost_with_weight": {"gradient_magnitude": 1.0, "image_gradient": 0.5}, "lr": 0.05}}
    slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, cfg, None)
    r = []
    for _ in range(12):
        got = slv.estimate(ev_s.numpy().astype(np.float64), flow0=flow0.numpy())
        r.append(float(np.sqrt(np.mean((got - ref) ** 2))))
    print(iters, "iterations: rms px min %.2e max %.2e" % (min(r), max(r)), ["%.1e" % v for v in r], flush=True)
