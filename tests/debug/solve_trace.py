"""Debug aid (not a test): per-iteration comparison of the fp64 fused CUDA solve with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import spec
from event_based_bos_b200 import ops

z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "reference_path_v1.npz"))
H, W, iters, lr, tvw = z["solve_f64/cfg"]; H, W, iters = int(H), int(W), int(iters)
ev = torch.from_numpy(z["solve_f64/events"])
dt = torch.float64
x = torch.zeros(2, H, W, dtype=dt); m = torch.zeros_like(x); v = torch.zeros_like(x)
win = ops.PreparedWindow(ev.cuda(), (H, W), "first", True, dtype=dt)
ws = ops.CmaxWorkspace(H, W, (0, 0), "cuda", dt)
xc = torch.zeros(2, H, W, dtype=dt, device="cuda"); mc = torch.zeros_like(xc); vc = torch.zeros_like(xc)
for it in range(1, 8):
    loss, grad = spec.cmax_value_and_grad(ev, x, (H, W), cost="gradient_magnitude", tv_weight=float(tvw))
    lc, gc = ops.cmax_value_and_grad(win, xc, "gradient_magnitude", 1.0, float(tvw), None, False, (0, 0), ws)
    gd = (gc.cpu() - grad).abs()
    xd = (xc.cpu() - x).abs()
    print(f"it {it}: loss {float(loss):.15e} vs {float(lc):.15e}; grad maxabs diff {gd.max():.3e} (|g|max {grad.abs().max():.3e}); x diff before step {xd.max():.3e}")
    if gd.max() > 1e-14:
        idx = torch.nonzero(gd > 1e-14)
        print("   n bad", len(idx), "first:", idx[:6].tolist())
        for c, r, col in idx[:6].tolist():
            print(f"   [{c},{r},{col}] cuda {gc[c,r,col].item():.6e} ref {grad[c,r,col].item():.6e} x_cuda {xc[c,r,col].item():.6e} x_ref {x[c,r,col].item():.6e}")
    spec.adam_update(x, grad, m, v, it, lr=float(lr))
    ops.adam_step(xc, gc, mc, vc, it, float(lr))
