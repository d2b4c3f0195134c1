"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container-only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

The reference (tub-rip/event_based_bos at /root/reference) has no tests or fixtures for this path,
so the oracle is pinned to outputs of the reference itself: this script imports the reference's
`Warp`, `EventImageConverter`, `costs.ImageGradient`, `utils.SobelTorch` and torch's Adam, feeds them
seeded inputs and stores inputs + outputs.  `inds` / `inds_mask` / `vals` of the bilinear vote are
captured by intercepting `Tensor.scatter_add_` while the reference runs (nothing is re-implemented
here).  The fixtures travel to the GPU box; the reference tree does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import ref_import

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class ScatterTap:
    """Records (index, src) of every Tensor.scatter_add_ call made while active."""

    def __init__(self):
        self.calls = []

    def __enter__(self):
        self._orig = torch.Tensor.scatter_add_
        tap = self

        def patched(self_t, dim, index, src):
            tap.calls.append((index.detach().clone(), src.detach().clone()))
            return tap._orig(self_t, dim, index, src)

        torch.Tensor.scatter_add_ = patched
        return self

    def __exit__(self, *a):
        torch.Tensor.scatter_add_ = self._orig


def make_events(rng, n, H, W, float_coords=False, t_max=1.0 / 120.0, dtype=np.float32):
    if float_coords:
        x = rng.uniform(0, H - 1e-3, n)
        y = rng.uniform(0, W - 1e-3, n)
    else:
        x = rng.integers(0, H, n).astype(np.float64)
        y = rng.integers(0, W, n).astype(np.float64)
    t = np.sort(rng.uniform(0.0, t_max, n))
    p = rng.integers(0, 2, n).astype(np.float64)
    return np.stack([x, y, t, p], axis=1).astype(dtype)


def main():
    ref = ref_import.load()
    Warp, Imager = ref.warp.Warp, ref.event_image_converter.EventImageConverter
    os.makedirs(OUT_DIR, exist_ok=True)
    g = {}
    torch.set_num_threads(1)

    # ---- 1. warp + vote, several configurations ------------------------------------------------
    cases = [
        # name, H, W, N, dtype, flow_max, float_coords, direction, padding
        ("c0_f32_first", 48, 64, 4096, np.float32, 3.0, False, "first", 0),
        ("c1_f32_middle_float", 48, 64, 4096, np.float32, 3.0, True, "middle", 0),
        ("c2_f32_last_oob", 40, 56, 4096, np.float32, 20.0, False, "last", 0),
        ("c3_f32_frac_pad", 40, 56, 3000, np.float32, 6.0, True, 0.3, 3),
        ("c4_f64_first", 48, 64, 4096, np.float64, 3.0, False, "first", 0),
        ("c5_f64_before_pad", 32, 40, 2000, np.float64, 10.0, True, "before", 2),
        ("c6_f32_after", 32, 40, 2000, np.float32, 2.0, False, "after", 0),
        ("c7_f32_dense", 16, 24, 6000, np.float32, 1.5, False, "first", 0),
    ]
    names = []
    for ci, (name, H, W, N, dt, fmax, fc, direction, pad) in enumerate(cases):
        rng = np.random.default_rng(100 + ci)
        ev = make_events(rng, N, H, W, float_coords=fc, dtype=dt)
        flow = rng.uniform(-fmax, fmax, (2, H, W)).astype(dt)
        warper = Warp((H, W), normalize_t=True)
        imager = Imager((H, W), outer_padding=pad)
        tev, tflow = torch.from_numpy(ev), torch.from_numpy(flow)
        warped, feat = warper.warp_event(tev, tflow, "dense-flow", direction=direction)
        with ScatterTap() as tap:
            iwe = imager.create_iwe(warped, method="bilinear_vote", sigma=0)
        inds, vals = tap.calls[0]
        g[f"{name}/events"] = ev
        g[f"{name}/flow"] = flow
        g[f"{name}/warped"] = warped.numpy()
        g[f"{name}/iwe"] = iwe.numpy()
        g[f"{name}/inds"] = inds.numpy().reshape(-1)
        g[f"{name}/vals"] = vals.numpy().reshape(-1)
        g[f"{name}/mask"] = imager.create_eventmask(warped).numpy()
        g[f"{name}/meta"] = np.array([H, W, pad, -1 if isinstance(direction, str) else direction], dtype=np.float64)
        g[f"{name}/direction"] = np.array(str(direction))
        names.append(name)
        assert set(feat.keys()) == {"determinant", "trace", "divergence", "straint", "absement"}
    g["warp_cases"] = np.array(names)

    # not normalised (velocity form) + numpy branch + weights + sigma + batched + 2dof
    rng = np.random.default_rng(7)
    H, W, N = 32, 48, 2500
    ev = make_events(rng, N, H, W, float_coords=True, dtype=np.float32)
    flow = rng.uniform(-300, 300, (2, H, W)).astype(np.float32)  # px / s
    warped, _ = Warp((H, W), normalize_t=False).warp_event(torch.from_numpy(ev), torch.from_numpy(flow), "dense-flow", "first")
    g["nonorm/events"], g["nonorm/flow"], g["nonorm/warped"] = ev, flow, warped.numpy()

    ev64 = make_events(rng, N, H, W, float_coords=True, dtype=np.float64)
    flow64 = rng.uniform(-3, 3, (2, H, W))
    warped_np, _ = Warp((H, W), normalize_t=True).warp_event(ev64, flow64, "dense-flow", "first")
    g["numpy/events"], g["numpy/flow"], g["numpy/warped"] = ev64, flow64, warped_np
    g["numpy/iwe_sigma0"] = Imager((H, W)).create_iwe(warped_np, "bilinear_vote", sigma=0)
    g["numpy/iwe_sigma1"] = Imager((H, W)).create_iwe(warped_np, "bilinear_vote", sigma=1)
    g["numpy/iwe_polarity"] = Imager((H, W)).create_iwe(warped_np, "polarity", sigma=0)
    g["numpy/eventmask"] = Imager((H, W)).create_eventmask(warped_np)

    wts = rng.uniform(0.0, 2.0, N).astype(np.float32)
    wev = torch.from_numpy(g["c1_f32_middle_float/warped"][:N])
    g["weighted/events"] = wev.numpy()
    g["weighted/weight"] = wts
    g["weighted/iwe"] = Imager((48, 64)).create_image_from_events_tensor(wev, "bilinear_vote", weight=torch.from_numpy(wts), sigma=0).numpy()
    g["sigma/iwe_sigma1"] = Imager((48, 64)).create_iwe(wev, "bilinear_vote", sigma=1).numpy()
    g["sigma/iwe_sigma3"] = Imager((48, 64)).create_iwe(wev, "bilinear_vote", sigma=3).numpy()

    evb = np.stack([make_events(rng, 1500, H, W, dtype=np.float32), make_events(rng, 1500, H, W, True, 0.02, np.float32)])
    flowb = rng.uniform(-3, 3, (2, 2, H, W)).astype(np.float32)
    wb, _ = Warp((H, W), normalize_t=True).warp_event(torch.from_numpy(evb), torch.from_numpy(flowb), "dense-flow", "middle")
    g["batched/events"], g["batched/flow"], g["batched/warped"] = evb, flowb, wb.numpy()
    g["batched/iwe"] = Imager((H, W)).create_iwe(wb, "bilinear_vote", sigma=0).numpy()

    theta = np.array([1.7, -2.3], dtype=np.float32)
    w2, _ = Warp((H, W), normalize_t=True).warp_event(torch.from_numpy(ev), torch.from_numpy(theta), "2d-translation", "first")
    g["twodof/theta"], g["twodof/warped"] = theta, w2.numpy()

    # ---- 2. TV regulariser (reference ImageGradient) -----------------------------------------------
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        rng = np.random.default_rng(21)
        fl = torch.from_numpy(rng.uniform(-3, 3, (2, 24, 36))).to(dt).requires_grad_()
        wt = torch.from_numpy(rng.uniform(0.1, 1.5, (24, 36))).to(dt)
        cost = ref.costs.functions["image_gradient"](direction="minimize")
        loss = cost.calculate({"flow": fl, "omit_boundary": False, "weights": wt})
        loss.backward()
        g[f"tv_{tag}/flow"], g[f"tv_{tag}/weights"] = fl.detach().numpy(), wt.numpy()
        g[f"tv_{tag}/loss"], g[f"tv_{tag}/grad"] = loss.detach().numpy(), fl.grad.numpy()
        fl2 = fl.detach().clone().requires_grad_()
        loss1 = cost.calculate({"flow": fl2, "omit_boundary": True, "weights": 1.0})
        loss1.backward()
        g[f"tv_{tag}/loss_w1"], g[f"tv_{tag}/grad_w1"] = loss1.detach().numpy(), fl2.grad.numpy()

    # ---- 3. composed path: loss and flow gradient through the reference operators --------------------
    def data_cost(kind, iwe, omit, prec):
        if kind == "image_variance":
            img = iwe[..., 1:-1, 1:-1] if omit else iwe
            return -torch.var(img)
        sob = ref.utils.SobelTorch(ksize=3, in_channels=1, precision=prec)(iwe[None, None]) / 8.0
        gx, gy = sob[:, 0], sob[:, 1]
        if omit:
            gx, gy = gx[..., 1:-1, 1:-1], gy[..., 1:-1, 1:-1]
        return -torch.mean(gx * gx + gy * gy)

    def composed_loss(ev, fl, H, W, kind, omit, tvw, pad, prec, direction="first"):
        warped, _ = Warp((H, W), normalize_t=True).warp_event(ev, fl, "dense-flow", direction=direction)
        iwe = Imager((H, W), outer_padding=pad).create_iwe(warped, method="bilinear_vote", sigma=0)
        loss = data_cost(kind, iwe, omit, prec)
        if tvw:
            loss = loss + tvw * ref.costs.functions["image_gradient"]().calculate({"flow": fl, "omit_boundary": omit, "weights": 1.0})
        return loss, iwe

    comp = []
    for ci, (H, W, N, fmax, kind, omit, tvw, pad, dt) in enumerate([
        (32, 48, 6000, 3.0, "image_variance", False, 0.5, 0, np.float64),
        (32, 48, 6000, 3.0, "gradient_magnitude", False, 0.5, 0, np.float64),
        (32, 48, 6000, 8.0, "gradient_magnitude", True, 0.0, 0, np.float64),
        (32, 48, 6000, 8.0, "image_variance", True, 0.25, 2, np.float64),
        (32, 48, 6000, 3.0, "image_variance", False, 0.5, 0, np.float32),
        (32, 48, 6000, 3.0, "gradient_magnitude", False, 0.5, 0, np.float32),
        (24, 40, 3000, 12.0, "gradient_magnitude", False, 0.1, 3, np.float32),
    ]):
        rng = np.random.default_rng(300 + ci)
        ev = make_events(rng, N, H, W, float_coords=(ci % 2 == 1), dtype=dt)
        fl = torch.from_numpy(rng.uniform(-fmax, fmax, (2, H, W)).astype(dt)).requires_grad_()
        loss, iwe = composed_loss(torch.from_numpy(ev), fl, H, W, kind, omit, tvw, pad, "64" if dt == np.float64 else "32")
        loss.backward()
        name = f"comp{ci}"
        g[f"{name}/events"], g[f"{name}/flow"] = ev, fl.detach().numpy()
        g[f"{name}/loss"], g[f"{name}/grad"], g[f"{name}/iwe"] = loss.detach().numpy(), fl.grad.numpy(), iwe.detach().numpy()
        g[f"{name}/cfg"] = np.array([H, W, omit, tvw, pad], dtype=np.float64)
        g[f"{name}/kind"] = np.array(kind)
        comp.append(name)
    g["comp_cases"] = np.array(comp)

    # ---- 4. Adam solve with the reference loop idiom (patch_eklt_pyramid2.py:259-288) -----------------
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        H, W, N, iters = 24, 32, 4000, 40
        rng = np.random.default_rng(55)
        ev = torch.from_numpy(make_events(rng, N, H, W, dtype=dt))
        x0 = torch.zeros(2, H, W, dtype=ev.dtype).requires_grad_()
        optimizer = torch.optim.Adam([x0], lr=0.05)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, iters, 0.1)
        hist = []
        for it in range(iters):
            optimizer.zero_grad()
            loss, _ = composed_loss(ev, x0, H, W, "gradient_magnitude", False, 0.5, 0, "64" if dt == np.float64 else "32")
            hist.append(float(loss))
            loss.backward()
            optimizer.step()
            scheduler.step()
        g[f"solve_{tag}/events"], g[f"solve_{tag}/flow"] = ev.numpy(), x0.detach().numpy()
        g[f"solve_{tag}/history"] = np.array(hist)
        g[f"solve_{tag}/cfg"] = np.array([H, W, iters, 0.05, 0.5])

    # Same loop from a tie-free initial iterate.  From an all-zero start the objective sits on exact ties
    # (sign(0) in the TV term, exactly-zero data gradients): a 1e-17 rounding residual is turned by Adam into a
    # 1e-10 step that breaks the tie and changes the gradient by O(1e-4), so two correct implementations (or two
    # torch versions) end up +-lr apart at those pixels.  A random initial flow removes the ties.
    for tag, dt, iters in (("f64", np.float64, 40), ("f32", np.float32, 40), ("f64_long", np.float64, 150)):
        H, W, N = 24, 32, 4000
        rng = np.random.default_rng(55)
        ev = torch.from_numpy(make_events(rng, N, H, W, dtype=dt))
        init = np.random.default_rng(56).uniform(-0.5, 0.5, (2, H, W)).astype(dt)
        x0 = torch.from_numpy(init.copy()).requires_grad_()
        optimizer = torch.optim.Adam([x0], lr=0.05)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, iters, 0.1)
        hist = []
        for it in range(iters):
            optimizer.zero_grad()
            loss, _ = composed_loss(ev, x0, H, W, "gradient_magnitude", False, 0.5, 0, "64" if dt == np.float64 else "32")
            hist.append(float(loss.detach()))
            loss.backward()
            optimizer.step()
            scheduler.step()
        g[f"solve_init_{tag}/events"], g[f"solve_init_{tag}/flow0"] = ev.numpy(), init
        g[f"solve_init_{tag}/flow"], g[f"solve_init_{tag}/history"] = x0.detach().numpy(), np.array(hist)
        g[f"solve_init_{tag}/cfg"] = np.array([H, W, iters, 0.05, 0.5])

    path = os.path.join(OUT_DIR, "reference_path_v1.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays, {os.path.getsize(path) / 1e6:.2f} MB; torch {torch.__version__}")


if __name__ == "__main__":
    sys.exit(main())
