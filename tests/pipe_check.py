"""Helper for test_gpu_fused.py::test_all_streaming_kernel_variants_match_oracle: run in a fresh process with
EBOS_TILE / EBOS_PIPE set to force one kernel family, print max relative errors of the fused path against the
CPU oracle as JSON."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_based_bos_b200 import ops  # noqa: E402
from oracle import spec  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


out = {}
H, W = 48, 80
for n in (1, 7, 2047, 2048, 2049, 70001):        # tails, multi-chunk CTAs, tiles with several work items (8192 events)
    for packed in (True, False):
        for weighted in (False, True):
            ev = torch.from_numpy(spec.synthetic_events(n, (H, W), seed=n))
            if not packed:
                ev[:, 0] = torch.clamp(ev[:, 0] + 0.25, max=H - 0.5)
            flow = torch.from_numpy(spec.synthetic_flow((H, W), seed=n, max_val=6.0))   # > halo of the tile kernels
            wts = torch.from_numpy(np.random.default_rng(n).uniform(0.5, 1.5, n).astype(np.float32)) if weighted else None
            f = flow.clone().requires_grad_()
            if n > 1:
                warped = spec.warp_dense_flow(ev, f, (H, W))
                iwe = spec.bilinear_vote(warped, (H, W), (1, 1), weight=1.0 if wts is None else wts)
                loss = spec.gradient_magnitude(iwe) + 0.5 * spec.total_variation(f, 1.0)
                loss.backward()
            win = ops.PreparedWindow(ev.cuda(), (H, W), "first", True, weight=None if wts is None else wts.cuda())
            assert win.packed == packed, (win.packed, packed)
            got_iwe = ops.window_splat(win, flow.cuda(), (1, 1)).cpu()
            l, g = ops.cmax_value_and_grad(win, flow.cuda(), "gradient_magnitude", 1.0, 0.5, None, False, (1, 1))
            lv, gv = ops.cmax_value_and_grad(win, flow.cuda(), "image_variance", 1.0, 0.0, None, True, (1, 1))
            torch.cuda.synchronize()
            key = f"n{n}_p{int(packed)}_w{int(weighted)}"
            if n > 1:
                out[key] = [rel(got_iwe.numpy(), iwe.detach().numpy()), rel(g.cpu().numpy(), f.grad.numpy()),
                            abs(float(l) - float(loss)) / abs(float(loss))]
            else:
                out[key] = [0.0 if torch.isnan(got_iwe).any() else 1.0, 0.0, 0.0]  # single timestamp: NaN lands on pixel 0
print(json.dumps(out))
