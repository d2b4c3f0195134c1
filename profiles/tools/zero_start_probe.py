"""Measurement probes used during round 1 (run on the GPU box from the repo root: python profiles/tools/<name>.py)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tests')
from test_gpu_fused import _run_solver, _rms
g = np.load('/root/repo/tests/golden/reference_path_v1.npz')
H, W, iters, lr, tvw = g["solve_f64/cfg"]; H, W = int(H), int(W)
print("ref gap", _rms(g["solve_f32/flow"], g["solve_f64/flow"]))
for precision in ("64", "32"):
    for fused, graph in ((True, True), (True, False), (False, False)):
        flow = _run_solver(g[f"solve_f{precision}/events"], H, W, iters, lr, tvw, precision, fused, graph)
        print(precision, fused, graph, _rms(flow, g["solve_f64/flow"]))
