"""Generate tests/golden/reference_metrics_v1.npz by running the UNMODIFIED reference (container-only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_metrics

Rows f-3 / f-4 of SURVEY.md section 8: the flow-error metrics (`utils.calculate_flow_error_numpy` /
`calculate_flow_error_tensor`, src/utils/flow_utils.py:705-821, with the event mask of
`EventImageConverter.create_eventmask`, src/solver/base.py:289-317) and the Gaussian-blurred IWE
(`create_image_from_events_tensor(..., sigma > 0)`, src/event_image_converter.py:399-404, forward and autograd
backward).  Nothing is re-implemented here: inputs are seeded, outputs are what the reference returns.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_metrics_v1.npz")
KEYS = ("EPE", "1PE", "2PE", "3PE", "5PE", "10PE", "20PE", "AE")


def main() -> None:
    ref = ref_import.load()
    out = {}
    rng = np.random.default_rng(42)
    cases = []
    # (name, batch, H, W, dtype, with mask, special ground-truth values)
    for name, B, H, W, dt, use_mask, special in (("plain64", 1, 24, 36, np.float64, False, False),
                                                 ("mask64", 2, 33, 47, np.float64, True, False),
                                                 ("holes64", 2, 20, 28, np.float64, True, True),
                                                 ("plain32", 1, 40, 56, np.float32, False, False),
                                                 ("mask32", 3, 17, 23, np.float32, True, True)):
        gt = (rng.uniform(-1, 1, (B, 2, H, W)) * rng.choice([0.5, 3.0, 15.0, 40.0], (B, 1, H, W))).astype(dt)
        pred = (gt + rng.normal(0, 1.0, gt.shape) * rng.choice([0.1, 2.0, 12.0], (B, 1, H, W))).astype(dt)
        pred[:, :, : H // 4] = gt[:, :, : H // 4]          # exact matches: cosine may exceed 1 by rounding
        if special:
            gt[0, 0, 3, 4] = 0.0                           # invalid: zero in one channel
            gt[0, :, 5, 6] = 0.0
            gt[-1, 1, 2, 2] = np.nan                       # NaN ground truth: |nan| > 0 is False -> masked out
            pred[0, 0, 7, 7] = 1e6
        mask = (rng.uniform(size=(B, 1, H, W)) > 0.35) if use_mask else None
        res = ref.utils.calculate_flow_error_numpy(gt, pred, event_mask=mask)
        out[f"{name}/gt"], out[f"{name}/pred"] = gt, pred
        if mask is not None:
            out[f"{name}/mask"] = mask
        out[f"{name}/numpy"] = np.array([res[k] for k in KEYS], dtype=np.float64)
        ts = rng.uniform(0.2, 2.0, (B, 1)).astype(dt)
        rt = ref.utils.calculate_flow_error_tensor(torch.from_numpy(gt), torch.from_numpy(pred),
                                                   event_mask=None if mask is None else torch.from_numpy(mask),
                                                   time_scale=torch.from_numpy(ts))
        out[f"{name}/time_scale"] = ts
        out[f"{name}/tensor"] = np.array([float(rt[k]) for k in KEYS], dtype=np.float64)
        cases.append(name)
    # infinite ground truth: upstream multiplies by the mask, inf * 0 = nan poisons EPE and AE (kept, not "fixed")
    gt = rng.uniform(-3, 3, (1, 2, 8, 9))
    pred = rng.uniform(-3, 3, (1, 2, 8, 9))
    gt[0, 0, 1, 1] = np.inf
    with np.errstate(all="ignore"):
        res = ref.utils.calculate_flow_error_numpy(gt, pred)
    out["inf64/gt"], out["inf64/pred"] = gt, pred
    out["inf64/numpy"] = np.array([res[k] for k in KEYS], dtype=np.float64)
    out["cases"] = np.array(cases)
    # event mask through the reference's own converter + SolverBase-style slicing
    H, W = 30, 44
    ev = np.stack([rng.integers(0, H, 400), rng.integers(0, W, 400), np.sort(rng.uniform(0, 0.01, 400)),
                   rng.integers(0, 2, 400)], 1).astype(np.float64)
    imager = ref.event_image_converter.EventImageConverter((H, W))
    out["evmask/events"] = ev
    out["evmask/mask"] = imager.create_eventmask(ev)

    # Gaussian-blurred IWE: forward + autograd backward through torchvision's gaussian_blur
    blur_cases = []
    for name, H, W, n, sigma, dt in (("b32", 24, 36, 3000, 1, torch.float32), ("b64", 31, 45, 5000, 3, torch.float64),
                                     ("b64s", 2, 5, 40, 2, torch.float64), ("b32f", 20, 28, 2000, 0.7, torch.float32)):
        ev = torch.from_numpy(np.stack([rng.uniform(0, H - 1, n), rng.uniform(0, W - 1, n), np.sort(rng.uniform(0, 1, n)),
                                        rng.integers(0, 2, n)], 1)).to(dt)
        ev.requires_grad_()
        imager = ref.event_image_converter.EventImageConverter((H, W))
        img = imager.create_image_from_events_tensor(ev, "bilinear_vote", sigma=sigma)
        probe = torch.from_numpy(rng.normal(size=(H, W))).to(dt)
        (img * probe).sum().backward()
        plain = imager.create_image_from_events_tensor(ev.detach(), "bilinear_vote", sigma=0)
        out[f"{name}/events"] = ev.detach().numpy()
        out[f"{name}/sigma"] = np.float64(sigma)
        out[f"{name}/iwe"] = plain.numpy()
        out[f"{name}/blurred"] = img.detach().numpy()
        out[f"{name}/probe"] = probe.numpy()
        out[f"{name}/grad_events"] = ev.grad.numpy()
        # the adjoint of the blur alone: d(sum(blur(I) * probe)) / dI
        I = plain.clone().requires_grad_()
        from torchvision.transforms.functional import gaussian_blur
        (gaussian_blur(I[None, None], kernel_size=3, sigma=sigma)[0, 0] * probe).sum().backward()
        out[f"{name}/adjoint"] = I.grad.numpy()
        blur_cases.append(name)
    out["blur_cases"] = np.array(blur_cases)

    # the reference's cost registry on the argument dict PatchEkltPyramid2 builds (src/solver/patch_eklt_pyramid2.py:
    # 380-392), through HybridCost with the weights of configs/hot_plate1.yaml; value, per-term history, autograd gradients
    H, W = 26, 38
    pred = torch.from_numpy(rng.normal(size=(H, W))).requires_grad_()
    meas = torch.from_numpy(rng.normal(size=(H, W)))
    flow = torch.from_numpy(rng.uniform(-2, 2, (2, H, W))).requires_grad_()
    pxy = torch.from_numpy(rng.uniform(-2, 2, (2, H, W))).requires_grad_()
    wts = torch.from_numpy(rng.uniform(0.1, 1.0, (H, W)))
    cww = {"diff_norm": 1.0, "image_gradient": 0.5, "flow_norm_pxy": 0.1, "flow_norm": 0.25}
    hybrid = ref.costs.HybridCost("minimize", cww, store_history=True)
    loss = hybrid.calculate({"prediction": pred, "measurement": meas, "weights": wts, "flow": flow, "omit_boundary": False,
                             "pxy": pxy})
    loss.backward()
    hist = hybrid.get_history()
    out["costs/prediction"], out["costs/measurement"] = pred.detach().numpy(), meas.numpy()
    out["costs/flow"], out["costs/pxy"], out["costs/weights"] = flow.detach().numpy(), pxy.detach().numpy(), wts.numpy()
    out["costs/names"] = np.array(list(cww))
    out["costs/weight_values"] = np.array(list(cww.values()))
    out["costs/loss"] = np.float64(loss.item())
    out["costs/terms"] = np.array([hist[k][0] for k in cww])
    out["costs/grad_prediction"], out["costs/grad_flow"], out["costs/grad_pxy"] = pred.grad.numpy(), flow.grad.numpy(), pxy.grad.numpy()
    out["costs/numpy_terms"] = np.array([
        ref.costs.functions["diff_norm"]("minimize").calculate({"prediction": pred.detach().numpy(), "measurement": meas.numpy(),
                                                                 "weights": None}),
        ref.costs.functions["flow_norm_pxy"]("minimize").calculate({"pxy": pxy.detach().numpy()}),
        ref.costs.functions["flow_norm"]("maximize").calculate({"flow": flow.detach().numpy()})])
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1e3:.1f} kB")


if __name__ == "__main__":
    main()
