"""Run-to-run spread of the fp32 solve against the reference goldens (not a test)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_fused import _run_solver, _rms
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "reference_path_v1.npz"))
H, W, iters, lr, tvw = g["solve_init_f32/cfg"]
a, b = [], []
for i in range(12):
    flow = _run_solver(g["solve_init_f32/events"], int(H), int(W), iters, lr, tvw, "32", True, i % 2 == 0, flow0=g["solve_init_f32/flow0"])
    a.append(_rms(flow, g["solve_init_f32/flow"])); b.append(_rms(flow, g["solve_init_f64/flow"]))
print("vs ref f32: min %.3e max %.3e" % (min(a), max(a)))
print("vs ref f64: min %.3e max %.3e" % (min(b), max(b)))
H, W, iters, lr, tvw = g["solve_f64/cfg"]
z = [_rms(_run_solver(g["solve_f32/events"], int(H), int(W), iters, lr, tvw, "32", True, True), g["solve_f64/flow"]) for _ in range(8)]
print("zero start fp32 vs ref f64: min %.3e max %.3e (ref gap %.3e)" % (min(z), max(z), _rms(g["solve_f32/flow"], g["solve_f64/flow"])))
