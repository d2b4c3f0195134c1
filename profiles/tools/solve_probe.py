"""Host-side timing probe of the solver phases (not a test)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from event_based_bos_b200 import solver
from event_based_bos_b200.utils import smooth_flow, synthetic_bos_events
H, W = 720, 1280
cfg = {"outer_padding": 0, "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": 600},
       "cmax": {"cost_with_weight": {"gradient_magnitude": 1.0, "image_gradient": 0.5}, "lr": 0.05}}
slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, cfg, None)
gt = smooth_flow((H, W), seed=0)
w = synthetic_bos_events(500000, (H, W), gt, seed=1).astype(np.float64)
slv.estimate(w)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    ev = slv._upload_events(w); x0 = slv._upload_flow0(None)
    t1 = time.perf_counter()
    adv, n = slv._plan_fused(ev, x0)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    for _ in range(n): adv()
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    out = slv._download(slv._finish(x0))
    t6 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.2f} ms | plan(host) {1e3*(t2-t1):.2f} | plan drain {1e3*(t3-t2):.2f} | issue {n} replays {1e3*(t4-t3):.2f} | gpu drain {1e3*(t5-t4):.2f} | finish+D2H {1e3*(t6-t5):.2f}")
