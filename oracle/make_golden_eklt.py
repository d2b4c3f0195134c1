"""Generate tests/golden/reference_eklt_v1.npz by running the UNMODIFIED reference (container-only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_eklt            # reference_eklt_v1.npz
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_eklt variants   # reference_eklt_variants_v1.npz

SURVEY 8f-1: the EKLT inner loop of `PatchEkltPyramid2` (what configs/hot_plate1.yaml runs).  The script builds
the reference solver from the shipped config (only the image size, ROI and iteration count are changed), lets it
compute its own frame gradients (`_set_frame`) and event histogram (`calculate_iwe_cache`,
`_make_measured_increment`), and records

  * per pyramid level: `_objective_scipy(theta)` and its torch-autograd gradient at theta = 0-translation start
    and at random theta (the two regimes of the grid_sample cell choice),
  * the dense fields `_extrapolate_dense_flow_from_estimates` / `..._translation_...` at the random theta,
  * a complete `estimate()` (coarse to fine, Adam) with the per-level results and the returned dense flow.

Nothing is re-implemented here.  The fixture travels to the GPU box; the reference tree does not.
"""
from __future__ import annotations

import copy
import os
import tempfile
from unittest import mock

import numpy as np
import torch
import yaml

from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_eklt_v1.npz")

IMAGE = (112, 176)
ROI = (8, 104, 40, 136)          # xmin, xmax (rows), ymin, ymax (cols)
N_EVENTS = 6000
N_ITER = 24


def build_solver(ref, n_iter, gml_overrides=None, drop_costs=()):
    cfg = yaml.safe_load(open(os.path.join(ref_import.REFERENCE_ROOT, "configs", "hot_plate1.yaml")))
    slv = copy.deepcopy(cfg["solver"])
    slv["optimizer"]["n_iter"] = n_iter
    slv["generative_ml"].update(gml_overrides or {})
    for k in drop_costs:
        slv["cost_with_weight"].pop(k)
    slv["filter"]["parameters"].update(dict(xmin=ROI[0], xmax=ROI[1], ymin=ROI[2], ymax=ROI[3]))
    vis = mock.MagicMock()
    vis.save_dir = tempfile.mkdtemp()
    cls = ref.solver.collections["patch_eklt_pyramid2"]
    s = cls(IMAGE, (ROI[1] - ROI[0], ROI[3] - ROI[2]), {}, slv, vis)
    s._video_maker = mock.MagicMock()
    return s, slv


def synthetic_inputs(seed=0):
    import cv2

    rng = np.random.default_rng(seed)
    H, W = IMAGE
    # events cluster on a blob texture so that the histogram has structure
    x = rng.integers(0, H, N_EVENTS)
    y = rng.integers(0, W, N_EVENTS)
    t = np.sort(rng.uniform(0, 1.0 / 120.0, N_EVENTS))
    p = (np.sin(x / 7.0) + np.cos(y / 9.0) + rng.normal(0, 0.5, N_EVENTS) > 0).astype(np.float64)
    events = np.stack([x, y, t, p], axis=1).astype(np.float64)
    frame = cv2.GaussianBlur(rng.uniform(0, 255, (H, W)), None, 3).astype(np.uint8)
    return events, frame


def main():
    ref = ref_import.load()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())          # @utils.profile writes optimize.prof into the cwd
    try:
        out = {}
        events, frame = synthetic_inputs()
        out["events"] = events
        out["frame"] = frame
        out["image"] = np.array(IMAGE)
        out["roi"] = np.array(ROI)
        out["n_iter"] = np.array(N_ITER)

        s, slv = build_solver(ref, N_ITER)
        out["cost_weights"] = np.array([slv["cost_with_weight"][k] for k in ("diff_norm", "image_gradient",
                                                                             "flow_norm_pxy")], dtype=np.float64)
        s._set_frame(frame)
        s.calculate_iwe_cache(events)
        out["grad_x"] = s._gradient_x.copy()
        out["grad_y"] = s._gradient_y.copy()
        out["histogram_blurred"] = s.cache_histogram.copy()
        out["weight_inverse"] = s.weight_inverse.copy()
        roi = {"xmin": ROI[0], "xmax": ROI[1], "ymin": ROI[2], "ymax": ROI[3]}

        rng = np.random.default_rng(1)
        levels = []
        for scale in range(s.coarest_scale, s.finest_scale):
            s.overload_patch_configuration(scale)
            s.estimate_mask_patch = torch.ones(s.patch_image_size).double()
            s.n_parameter_dim = len(s._initialize_velocity())
            meas_np, w_np = s._make_measured_increment(events, roi)
            assert w_np is None
            meas = torch.from_numpy(meas_np).double() * s.estimate_mask_dense()
            if scale == s.coarest_scale:
                out["measured"] = meas.numpy().copy()
            ph, pw = s.patch_image_size
            levels.append((s.patch_size[0], ph, pw))
            thetas = {
                # the reference's start: random intensity, ZERO translation (samples sit beside the pixel centres)
                "start": np.concatenate([rng.uniform(-1, 1, (1, ph, pw)), np.zeros((2, ph, pw))]),
                "random": np.concatenate([rng.uniform(-1, 1, (1, ph, pw)), rng.uniform(-1.5, 1.5, (2, ph, pw))]),
                # translations that push samples out of the image (zeros padding of grid_sample)
                "far": np.concatenate([rng.uniform(-1, 1, (1, ph, pw)), rng.uniform(-40, 40, (2, ph, pw))]),
            }
            for name, th in thetas.items():
                x = torch.from_numpy(th).double().requires_grad_()
                loss = s._objective_scipy(x, meas, roi, None)
                loss.backward()
                key = f"L{scale}_{name}"
                out[key + "_theta"] = th
                out[key + "_loss"] = np.array(loss.item())
                out[key + "_grad"] = x.grad.numpy().copy()
                if name == "random":
                    out[key + "_flow"] = s._extrapolate_dense_flow_from_estimates(x).detach().numpy().copy()
                    out[key + "_trans"] = s._extrapolate_dense_translation_from_estimates(x).detach().numpy().copy()
                    out[key + "_pred"] = s._make_prediction_torch(x, roi, None).detach().numpy().copy()
            s.cost_func.clear_history()
        out["levels"] = np.array(levels)

        # ---- a complete coarse-to-fine estimate ------------------------------------------------------------
        s, _ = build_solver(ref, N_ITER)
        np.random.seed(7)                   # `_initialize_velocity` draws the intensity start from np.random
        per_scale = {}
        orig = s.run_estimation_per_scale

        def recording(ev, params):
            r = orig(ev, params)
            per_scale[s.current_scale] = r.copy()
            return r

        s.run_estimation_per_scale = recording
        starts = []
        orig_init = s._initialize_velocity

        def init_recording():
            v = orig_init()
            starts.append(v.copy())
            return v

        s._initialize_velocity = init_recording
        flow = s.estimate(events, frame=frame)
        out["solve_flow"] = flow
        # x0 of the coarsest level: one draw per patch AFTER the first call that only measures the dimension
        ph, pw = levels[0][1], levels[0][2]
        x0 = np.concatenate(starts[1:1 + ph * pw]).reshape((3, ph, pw))
        out["solve_x0"] = x0
        for k, v in per_scale.items():
            out[f"solve_L{k}"] = v
        np.savez_compressed(OUT, **out)
        print("wrote", OUT, os.path.getsize(OUT), "bytes;", "levels", levels)
    finally:
        os.chdir(cwd)


OUT_VARIANTS = OUT.replace("reference_eklt_v1", "reference_eklt_variants_v1")

# the other switch combinations of generative_ml.* (same inputs, level of 16-px patches, random theta)
VARIANTS = {
    "flow_warp": ({"poisson_model": False}, ()),                                   # theta = (v_x, v_y, p_x, p_y)
    "poisson_nowarp": ({"optimize_warp": False}, ("flow_norm_pxy",)),              # theta = (intensity)
    "flow_nowarp": ({"poisson_model": False, "optimize_warp": False}, ("flow_norm_pxy",)),
    "no_polarity": ({"no_polarity": True}, ()),
    "hist_weights": ({"weight_loss_by_event_hist": True}, ()),
    "all": ({"poisson_model": False, "no_polarity": True, "weight_loss_by_event_hist": True}, ()),
}


def make_variants(ref):
    out = {}
    events, frame = synthetic_inputs()
    roi = {"xmin": ROI[0], "xmax": ROI[1], "ymin": ROI[2], "ymax": ROI[3]}
    rng = np.random.default_rng(11)
    for name, (gml, drop) in VARIANTS.items():
        s, slv = build_solver(ref, N_ITER, gml, drop)
        s._set_frame(frame)
        s.calculate_iwe_cache(events)
        s.overload_patch_configuration(3)
        s.estimate_mask_patch = torch.ones(s.patch_image_size).double()
        s.n_parameter_dim = len(s._initialize_velocity())
        meas_np, w_np = s._make_measured_increment(events, roi)
        meas = torch.from_numpy(meas_np).double() * s.estimate_mask_dense()
        weights = None if w_np is None else torch.from_numpy(w_np).double() * s.estimate_mask_dense()
        ph, pw = s.patch_image_size
        nd = s.n_parameter_dim
        th = rng.uniform(-1, 1, (nd, ph, pw))
        if gml.get("optimize_warp", True):
            th[-2:] *= 1.5
        x = torch.from_numpy(th).double().requires_grad_()
        loss = s._objective_scipy(x, meas, roi, weights)
        loss.backward()
        out[name + "_theta"] = th
        out[name + "_loss"] = np.array(loss.item())
        out[name + "_grad"] = x.grad.numpy().copy()
        out[name + "_measured"] = meas.numpy().copy()
        if weights is not None:
            out[name + "_weights"] = weights.numpy().copy()
        if gml.get("no_polarity", False):        # weight_inverse comes from |pos + neg| instead of |pos - neg|
            out[name + "_weight_inverse"] = s.weight_inverse.copy()
        out[name + "_cost_weights"] = np.array([slv["cost_with_weight"].get(k, 0.0) for k in
                                                ("diff_norm", "image_gradient", "flow_norm_pxy")], dtype=np.float64)
        out[name + "_flags"] = np.array([int(gml.get("poisson_model", True)), int(gml.get("optimize_warp", True)),
                                         int(gml.get("no_polarity", False))])
    out["patch"] = np.array(16)
    np.savez_compressed(OUT_VARIANTS, **out)
    print("wrote", OUT_VARIANTS, os.path.getsize(OUT_VARIANTS), "bytes")


if __name__ == "__main__":
    import sys

    if "variants" in sys.argv[1:]:
        cwd = os.getcwd()
        os.chdir(tempfile.mkdtemp())
        try:
            make_variants(ref_import.load())
        finally:
            os.chdir(cwd)
    else:
        main()
