"""Measurement probes used during round 1 (run on the GPU box from the repo root: python profiles/tools/<name>.py)."""
import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from event_based_bos_b200 import ops
from event_based_bos_b200.utils import synthetic_events, synthetic_flow
H, W = 720, 1280
n = 1 << 24
ev = torch.from_numpy(synthetic_events(n, (H, W), seed=0)).cuda()
flow = torch.from_numpy(synthetic_flow((H, W), seed=0)).cuda()
win = ops.PreparedWindow(ev, (H, W), "first", True)
iwe = ops.window_splat(win, flow).clone()
torch.cuda.synchronize()
torch.save(iwe.cpu(), sys.argv[1])
print("sum", float(iwe.double().sum()), "max", float(iwe.max()))
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    d = (iwe.cpu().double() - ref.double()).abs()
    print("max abs diff", float(d.max()), "rel-to-max", float(d.max() / ref.abs().max()), "rms", float(d.pow(2).mean().sqrt()))
