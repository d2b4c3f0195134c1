"""SURVEY 8f-1, CPU side.

1. oracle/spec_eklt.py against tests/golden/reference_eklt_v1.npz (outputs of the unmodified reference
   `PatchEkltPyramid2`: objective value, autograd gradient, dense fields, a complete coarse-to-fine estimate).
2. the per-pixel / per-cell functions the CUDA kernels are built from (csrc/ebos_eklt_math.cuh), compiled by g++
   into a serial checker (tests/eklt_host_check.cpp) and walked in kernel order, against the oracle and the goldens.
   This is a check of the ARITHMETIC in the GPU-less build container; the GPU parity tests are in test_gpu_eklt.py.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import spec_eklt as E

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_eklt_v1.npz")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLDEN)
    d = {k: g[k] for k in g.files}
    d["roi_t"] = tuple(int(v) for v in d["roi"])
    d["levels_t"] = [tuple(int(v) for v in l) for l in d["levels"]]
    return d


def _objective(gold, theta, patch, **kw):
    wd, wtv, wp = gold["cost_weights"]
    return E.objective(theta, gold["grad_x"], gold["grad_y"], gold["measured"], gold["weight_inverse"], gold["roi_t"],
                       patch, wd, wtv, wp, **kw)


# ---- 1. oracle vs reference -------------------------------------------------------------------------------------
def test_pyramid_geometry_matches_reference(gold):
    H, W = (int(v) for v in gold["image"])
    assert E.pyramid_levels((H, W)) == gold["levels_t"]
    # hot_plate1 size: 12 x 20 patches of 64 px, 24 rows off the lattice
    assert E.pyramid_levels((720, 1280)) == [(64, 12, 20), (32, 23, 40), (16, 45, 80), (8, 90, 160)]
    g = E.upsample_geometry((720, 1280), 12, 20, 64)
    assert (g["h1"], g["w1"], g["dense_h"]) == (88, 64, 896)
    assert [E.level_iterations(600, 4, l) for l in range(4)] == [120, 150, 200, 300]


@pytest.mark.parametrize("name", ["start", "random", "far"])
def test_oracle_objective_and_gradient_match_reference_autograd(gold, name):
    for scale, (patch, ph, pw) in enumerate(gold["levels_t"], 1):
        key = f"L{scale}_{name}"
        r = _objective(gold, gold[key + "_theta"], patch)
        assert abs(r["loss"] - float(gold[key + "_loss"])) <= 1e-13
        ref = gold[key + "_grad"]
        assert np.abs(r["grad"] - ref).max() <= 1e-12 * np.abs(ref).max()
        if name == "random":
            assert np.abs(r["flow"] - gold[key + "_flow"]).max() <= 1e-14
            assert np.abs(r["trans"] - gold[key + "_trans"]).max() <= 1e-14
            assert np.abs(r["pred"] - gold[key + "_pred"]).max() <= 1e-15


def test_oracle_gradient_matches_finite_differences(gold):
    """Independent of the reference: central differences of the restated forward (smooth directions only)."""
    patch, ph, pw = gold["levels_t"][2]
    th = gold["L3_random_theta"].copy()
    r = _objective(gold, th, patch)
    rng = np.random.default_rng(0)
    for _ in range(3):
        d = rng.normal(size=th.shape)
        d[0] = 0                                      # TV of the intensity flow is only piecewise smooth; keep to pxy
        eps = 1e-7
        lp = _objective(gold, th + eps * d, patch, want_grad=False)["loss"]
        lm = _objective(gold, th - eps * d, patch, want_grad=False)["loss"]
        fd = (lp - lm) / (2 * eps)
        an = float(np.sum(r["grad"] * d))
        assert abs(fd - an) <= 2e-5 * max(1.0, abs(an))


def test_oracle_solve_matches_reference_estimate(gold):
    """Coarse-to-fine Adam from the reference's own start.  The first level is reproduced to rounding; later levels
    drift by ~1e-6 because sign() of rounding-level TV differences (replicate-padded rows) is chaotic -- in the
    reference itself as much as here."""
    H, W = (int(v) for v in gold["image"])
    wd, wtv, wp = gold["cost_weights"]
    th = gold["solve_x0"].copy()
    n_iter = int(gold["n_iter"])
    levels = gold["levels_t"]
    tol = [1e-12, 1e-4, 1e-4, 1e-4]
    for li, (patch, ph, pw) in enumerate(levels):
        if li > 0:
            th = E.resize_params(th, (ph, pw))
        th, losses = E.solve_level(th, E.level_iterations(n_iter, len(levels), li), gold["grad_x"], gold["grad_y"],
                                   gold["measured"], gold["weight_inverse"], gold["roi_t"], patch, w_data=wd, w_tv=wtv,
                                   w_pxy=wp)
        assert np.abs(th - gold[f"solve_L{li + 1}"]).max() <= tol[li]
    patch, ph, pw = levels[-1]
    dense = E.upsample_patch(E.sobel_over_8(th[0]), patch, (H, W)) * E.roi_mask((H, W), gold["roi_t"])
    assert np.sqrt(np.mean((dense - gold["solve_flow"]) ** 2)) <= 1e-5


# ---- 2. kernel arithmetic (serial g++ build of the device functions) vs oracle -----------------------------------
@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("eklt") / "libeklt_host_check.so"
    cmd = [gxx, "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "eklt_host_check.cpp"),
           "-o", str(out)]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_value_and_grad(lib, theta, gx, gy, meas, winv, roi, patch, weights, dtype=np.float64, poisson=True, warp=True,
                        no_polarity=False, hist_weights=None, stored=False):
    H, W = gx.shape
    nt, ph, pw = theta.shape
    wd, wtv, wp = (float(v) for v in weights)
    flags = (1 if poisson else 0) | (2 if warp else 0) | (4 if no_polarity else 0)
    dims = (ctypes.c_int * 9)(H, W, ph, pw, patch, *roi)
    f64 = int(dtype == np.float64)
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    theta, gx, gy, meas, winv = c(theta), c(gx), c(gy), c(meas), c(winv)
    hw = None if hist_weights is None else c(hist_weights)
    hw_p = None if hw is None else _p(hw)
    pf = np.zeros((2, ph, pw), dtype)
    q = np.zeros((H, W), dtype)
    F = np.zeros((2, H, W), dtype)
    tr = np.zeros((2, H, W), dtype)
    sums = np.zeros(2)
    lib.eklt_host_forward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 9
    lib.eklt_host_forward(dims, f64, flags, _p(theta), _p(gx), _p(gy), hw_p, _p(pf), _p(q), _p(F), _p(tr), _p(sums))
    colsum = np.zeros(W)
    scal = np.zeros(4)
    lib.eklt_host_columns.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    lib.eklt_host_columns(dims, f64, _p(q), _p(meas), float(sums[0]), wd, _p(colsum), _p(scal))
    # TV of the masked flow: the gather-form functions of k_tv_roi over the ROI box (dF outside the box stays NaN:
    # the backward must not read it)
    dF = np.full((2, H, W), np.nan, dtype)
    tv_sum = ctypes.c_double(0.0)
    lib.eklt_host_tv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                 ctypes.c_void_p, ctypes.c_void_p]
    lib.eklt_host_tv(dims, f64, _p(F), _p(winv), wtv / (2.0 * H * W), _p(dF), ctypes.byref(tv_sum))
    tv = tv_sum.value / (2.0 * H * W)
    dU = np.zeros((4, H, W), dtype)
    dPad = np.zeros((4, ph + 2, pw + 2), dtype)
    dP = np.zeros((4, ph, pw), dtype)
    grad = np.zeros((nt, ph, pw), dtype)
    lib.eklt_host_backward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 9 + \
                                      [ctypes.c_double] + [ctypes.c_void_p] * 4
    if stored:
        assert f64
        lib.eklt_host_backward_stored.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 9 + \
                                                 [ctypes.c_double] + [ctypes.c_void_p] * 4
        lib.eklt_host_backward_stored(dims, flags, _p(theta), _p(pf), _p(gx), _p(gy), hw_p, _p(meas), _p(dF), _p(colsum),
                                      _p(scal), wp, _p(dU), _p(dPad), _p(dP), _p(grad))
    else:
        lib.eklt_host_backward(dims, f64, flags, _p(theta), _p(pf), _p(gx), _p(gy), hw_p, _p(meas), _p(dF), _p(colsum),
                               _p(scal), wp, _p(dU), _p(dPad), _p(dP), _p(grad))
    loss = wd * scal[1] + wtv * tv + (wp * sums[1] / (H * W) if warp else 0.0)
    return {"loss": loss, "grad": grad, "q": q, "F": F, "trans": tr, "pf": pf, "colsum": colsum, "n": scal[0], "dU": dU,
            "dF": dF, "tv": tv}


@pytest.mark.parametrize("name", ["start", "random", "far"])
def test_kernel_arithmetic_fp64_matches_oracle_and_reference(gold, host_lib, name):
    for scale, (patch, ph, pw) in enumerate(gold["levels_t"], 1):
        key = f"L{scale}_{name}"
        th = gold[key + "_theta"]
        h = host_value_and_grad(host_lib, th, gold["grad_x"], gold["grad_y"], gold["measured"], gold["weight_inverse"],
                                gold["roi_t"], patch, gold["cost_weights"])
        r = _objective(gold, th, patch)
        # same operation order as the oracle: the fields agree bit for bit (this is what keeps the sign() of the
        # rounding-level TV differences identical)
        assert np.array_equal(h["pf"], E.sobel_over_8(th[0]))
        assert np.array_equal(h["F"], r["flow"] * E.roi_mask(r["q"].shape, gold["roi_t"])[None])
        assert np.array_equal(h["trans"], r["trans"])
        assert np.abs(h["q"] - r["q"]).max() <= 1e-12 * np.abs(r["q"]).max()
        assert np.abs(h["colsum"] - r["colsum"]).max() <= 1e-13
        assert abs(h["loss"] - r["loss"]) <= 1e-13
        assert np.abs(h["grad"] - r["grad"]).max() <= 1e-11 * np.abs(r["grad"]).max()
        ref = gold[key + "_grad"]
        assert abs(h["loss"] - float(gold[key + "_loss"])) <= 1e-13
        assert np.abs(h["grad"] - ref).max() <= 1e-11 * np.abs(ref).max()


def test_kernel_arithmetic_fp32_close_to_fp64(gold, host_lib):
    """The fp32 instantiation (speed path) on fp32-rounded inputs stays within fp32 accuracy of the fp64 reference --
    also at the zero-translation start, because the sample positions are evaluated in double in both instantiations
    (with float32 positions the start gradient was 3-35 % off: another bilinear cell, another one-sided difference)."""
    for scale, (patch, ph, pw) in enumerate(gold["levels_t"], 1):
        for name in ("random", "start"):
            key = f"L{scale}_{name}"
            h = host_value_and_grad(host_lib, gold[key + "_theta"], gold["grad_x"], gold["grad_y"], gold["measured"],
                                    gold["weight_inverse"], gold["roi_t"], patch, gold["cost_weights"], dtype=np.float32)
            ref = gold[key + "_grad"]
            assert abs(h["loss"] - float(gold[key + "_loss"])) <= 2e-6 * abs(float(gold[key + "_loss"])), key
            assert np.abs(h["grad"] - ref).max() <= 5e-6 * np.abs(ref).max(), key


def test_kernel_arithmetic_general_roi_and_odd_sizes(host_lib):
    """Sizes that are not multiples of anything, ROI touching the border, patch grid with a single row."""
    rng = np.random.default_rng(5)
    # the last three: patch sizes that are not powers of two (true division in the tap arithmetic, odd supports)
    for (H, W, patch, roi) in [(37, 53, 8, (0, 37, 0, 53)), (50, 70, 64, (3, 47, 10, 70)), (33, 130, 16, (5, 6, 7, 9)),
                               (37, 53, 5, (2, 30, 0, 53)), (40, 60, 12, (0, 40, 7, 41)), (30, 31, 7, (1, 29, 1, 30))]:
        ph, pw = E.patch_grid((H, W), patch)
        th = np.concatenate([rng.uniform(-1, 1, (1, ph, pw)), rng.uniform(-2, 2, (2, ph, pw))])
        gx, gy = rng.normal(size=(2, H, W)) * 50
        M = E.roi_mask((H, W), roi)
        meas = rng.normal(size=(H, W)) * M
        meas /= np.linalg.norm(meas)
        winv = rng.uniform(0.05, 1.0, (H, W))
        w = (1.0, 0.5, 0.1)
        h = host_value_and_grad(host_lib, th, gx, gy, meas, winv, roi, patch, w)
        r = E.objective(th, gx, gy, meas, winv, roi, patch, *w)
        assert abs(h["loss"] - r["loss"]) <= 1e-12
        assert np.abs(h["grad"] - r["grad"]).max() <= 1e-10 * np.abs(r["grad"]).max()


def test_preprocessing_oracle_matches_reference(gold):
    """`_set_frame` / `calculate_iwe_cache` / `_make_measured_increment` restated (oracle) vs the reference's own
    cv2.Sobel, cv2.GaussianBlur and scipy gaussian_filter outputs stored in the fixture."""
    H, W = (int(v) for v in gold["image"])
    gx, gy = E.frame_gradients(gold["frame"])
    assert np.array_equal(gx, gold["grad_x"]) and np.array_equal(gy, gold["grad_y"])
    meas, winv, blurred = E.measurement_and_weights(gold["events"], (H, W), gold["roi_t"])
    assert np.abs(blurred - gold["histogram_blurred"]).max() <= 1e-14
    assert np.abs(meas - gold["measured"]).max() <= 1e-15
    assert np.abs(winv - gold["weight_inverse"]).max() <= 1e-13


def test_separable_correlation_arithmetic_matches_oracle(gold, host_lib):
    """The border indexing and tap loop of ebos_sepconv2d (serial build) vs the oracle, both border modes."""
    rng = np.random.default_rng(2)
    for (H, W) in [(19, 45), (1, 7), (6, 1), (112, 176)]:
        img = rng.normal(size=(H, W))
        for taps_r, taps_c in [(np.array([-1.0, 0.0, 1.0]), np.array([1.0, 2.0, 1.0])),
                               (E.cv2_gaussian_taps(2.0), E.cv2_gaussian_taps(2.0)),
                               (E.scipy_gaussian_taps(10.0), E.scipy_gaussian_taps(3.0))]:
            for border in (0, 1):
                tmp, out = np.zeros((H, W)), np.zeros((H, W))
                tr, tc = np.ascontiguousarray(taps_r), np.ascontiguousarray(taps_c)
                host_lib.eklt_host_sepconv(_p(img), H, W, _p(tr), len(tr), _p(tc), len(tc), border, 1, _p(tmp), _p(out))
                ref = E.correlate_separable(img, tr, tc, border)
                assert np.abs(out - ref).max() <= 1e-13, (H, W, len(tr), border)


HOT_PLATE1_SOLVER = {
    "filter": {"filters": None, "parameters": {"xmin": 0, "xmax": 720, "ymin": 320, "ymax": 960}},
    "method": "patch_eklt_pyramid2", "cost_with_weight": {"diff_norm": 1.0, "image_gradient": 0.5, "flow_norm_pxy": 0.1},
    "optimizer": {"method": "Adam", "n_iter": 600, "parameters": {}},
    "generative_ml": {"weight_loss_by_event_hist": False, "weight_loss_by_inverse_event_hist": True, "optimize_warp": True,
                      "iwe_sigma": 2, "no_polarity": False, "model_image": "current", "use_log_intensity": False,
                      "poisson_model": True},
    "patch_eklt": {"patch_size": 4, "sliding_window": 2, "do_event_thresholding": False, "event_thres": 8},
}


def test_solver_registry_and_config_contract():
    """Host logic of the drop-in solver (no GPU): registry name, level schedule, start-value draws, and loud failure
    for objectives other than hot_plate1's."""
    import copy

    from event_based_bos_b200 import eklt, solver

    cls = solver.collections["patch_eklt_pyramid2"]
    s = cls((720, 1280), (720, 640), {}, copy.deepcopy(HOT_PLATE1_SOLVER), None)
    assert s.levels == [(64, 12, 20), (32, 23, 40), (16, 45, 80), (8, 90, 160)] == E.pyramid_levels((720, 1280))
    assert (s.coarest_scale, s.finest_scale) == (1, 5)
    assert [600 // (s.finest_scale - sc + 1) for sc in range(1, 5)] == [E.level_iterations(600, 4, l) for l in range(4)]
    assert s.estimate_mask_dense_numpy.sum() == 720 * 640
    np.random.seed(3)
    v = s._initialize_velocity()
    np.random.seed(3)
    assert v.shape == (3,) and v[0] == np.random.random() * 2.0 - 1 and v[1] == v[2] == 0.0
    assert eklt.patch_grid((720, 1280), 64) == E.patch_grid((720, 1280), 64)
    assert np.allclose(eklt.gaussian_taps_cv2(2.0), E.cv2_gaussian_taps(2.0), rtol=0, atol=0)
    assert np.allclose(eklt.gaussian_taps_scipy(10.0), E.scipy_gaussian_taps(10.0), rtol=0, atol=0)
    for path, value in [(("optimizer", "method"), "Newton-CG"), (("generative_ml", "angle_model"), True),
                        (("generative_ml", "sobel_ksize"), 5), (("cost_with_weight", "total_variation"), 3.0)]:
        cfg = copy.deepcopy(HOT_PLATE1_SOLVER)
        cfg[path[0]][path[1]] = value
        with pytest.raises(NotImplementedError):
            cls((720, 1280), (720, 640), {}, cfg, None)
    cfg = copy.deepcopy(HOT_PLATE1_SOLVER)
    cfg["generative_ml"]["optimize_warp"] = False               # flow_norm_pxy without a translation: KeyError upstream too
    with pytest.raises(KeyError):
        cls((720, 1280), (720, 640), {}, cfg, None)
    # start values of the other parameterisations (src/solver/generative_max_likelihood.py:425-450)
    for poisson, warp, n_dim in [(True, False, 1), (False, True, 4), (False, False, 2)]:
        cfg = copy.deepcopy(HOT_PLATE1_SOLVER)
        cfg["generative_ml"].update({"poisson_model": poisson, "optimize_warp": warp, "weight_sigma": 5})
        if not warp:
            cfg["cost_with_weight"].pop("flow_norm_pxy")
        v = cls((720, 1280), (720, 640), {}, cfg, None)._initialize_velocity()
        assert v.shape == (n_dim,) and (poisson or not v.any())
    import torch

    if not torch.cuda.is_available():          # no CPU fallback: the estimate needs the device
        with pytest.raises(RuntimeError):
            s.estimate(np.zeros((4, 4)), frame=np.zeros((720, 1280), np.uint8))


# ---- 3. the other switch combinations of generative_ml.* ---------------------------------------------------------
VARIANTS = os.path.join(HERE, "golden", "reference_eklt_variants_v1.npz")


def _variant_cases(gold):
    v = np.load(VARIANTS)
    for name in sorted({k[:-len("_theta")] for k in v.files if k.endswith("_theta")}):
        poisson, warp, no_pol = (bool(x) for x in v[name + "_flags"])
        yield {
            "name": name, "theta": v[name + "_theta"], "loss": float(v[name + "_loss"]), "grad": v[name + "_grad"],
            "measured": v[name + "_measured"], "cost_weights": tuple(float(x) for x in v[name + "_cost_weights"]),
            "weights": v[name + "_weights"] if name + "_weights" in v.files else None,
            "winv": v[name + "_weight_inverse"] if name + "_weight_inverse" in v.files else gold["weight_inverse"],
            "poisson": poisson, "warp": warp, "no_polarity": no_pol, "patch": int(v["patch"]),
        }


def test_oracle_switch_variants_match_reference_autograd(gold):
    """poisson_model / optimize_warp / no_polarity / weight_loss_by_event_hist in the combinations of
    oracle/make_golden_eklt.py (run on the unmodified reference with the shipped config otherwise)."""
    names = []
    for c in _variant_cases(gold):
        wd, wtv, wp = c["cost_weights"]
        r = E.objective(c["theta"], gold["grad_x"], gold["grad_y"], c["measured"], c["winv"], gold["roi_t"], c["patch"],
                        wd, wtv, wp, poisson=c["poisson"], warp=c["warp"], no_polarity=c["no_polarity"],
                        weights=c["weights"])
        assert abs(r["loss"] - c["loss"]) <= 1e-13, c["name"]
        assert np.abs(r["grad"] - c["grad"]).max() <= 1e-12 * np.abs(c["grad"]).max(), c["name"]
        names.append(c["name"])
    assert names == ["all", "flow_nowarp", "flow_warp", "hist_weights", "no_polarity", "poisson_nowarp"]


def test_kernel_arithmetic_switch_variants_match_reference(gold, host_lib):
    """The same combinations through the serial build of the device functions (flags / weights of the C-ABI)."""
    for c in _variant_cases(gold):
        h = host_value_and_grad(host_lib, c["theta"], gold["grad_x"], gold["grad_y"], c["measured"], c["winv"],
                                gold["roi_t"], c["patch"], c["cost_weights"], poisson=c["poisson"], warp=c["warp"],
                                no_polarity=c["no_polarity"], hist_weights=c["weights"])
        assert abs(h["loss"] - c["loss"]) <= 1e-13, c["name"]
        assert np.abs(h["grad"] - c["grad"]).max() <= 1e-11 * np.abs(c["grad"]).max(), c["name"]


def test_preprocessing_oracle_variants_match_reference(gold):
    """no_polarity (|pos + neg| histogram) and weight_loss_by_event_hist (Gaussian-blurred |hist| weights)."""
    v = np.load(VARIANTS)
    H, W = (int(x) for x in gold["image"])
    m, wi, _, w = E.measurement_and_weights(gold["events"], (H, W), gold["roi_t"], weight_sigma=5.0)
    assert np.abs(m - v["hist_weights_measured"]).max() <= 1e-15
    assert np.abs(w - v["hist_weights_weights"]).max() <= 1e-14
    m, wi, _ = E.measurement_and_weights(gold["events"], (H, W), gold["roi_t"], no_polarity=True)
    assert np.abs(m - v["no_polarity_measured"]).max() <= 1e-15
    assert np.abs(wi - v["no_polarity_weight_inverse"]).max() <= 1e-13


def test_tv_gather_form_matches_oracle_adjoint(gold, host_lib):
    """k_tv_roi's arithmetic: value and gradient of the TV term restricted to the ROI box equal the oracle's full-image
    torch.gradient adjoint wherever the backward reads it (inside the ROI), for ROIs on and off the image border."""
    rng = np.random.default_rng(9)
    for (H, W, roi) in [(20, 31, (0, 20, 0, 31)), (20, 31, (1, 19, 2, 30)), (9, 12, (3, 4, 5, 6)), (2, 2, (0, 2, 0, 2)),
                        (16, 16, (0, 1, 15, 16))]:
        M = E.roi_mask((H, W), roi)
        F = rng.normal(size=(2, H, W)) * M[None]
        F[:, roi[0]:roi[0] + 1] = np.round(F[:, roi[0]:roi[0] + 1])          # exact ties -> sign(0) paths
        winv = rng.uniform(0.05, 1.0, (H, W))
        dims = (ctypes.c_int * 9)(H, W, 1, 1, max(H, W), *roi)
        dF = np.full((2, H, W), np.nan)
        val = ctypes.c_double(0.0)
        host_lib.eklt_host_tv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_void_p]
        host_lib.eklt_host_tv(dims, 1, _p(F), _p(winv), 1.0 / F.size, _p(dF), ctypes.byref(val))
        gx, gy = E._tv_parts(F, winv)
        ref = E._tv_adjoint(np.sign(gx) * winv / F.size, np.sign(gy) * winv / F.size)
        inside = M[None].astype(bool).repeat(2, 0)
        assert np.abs(dF[inside] - ref[inside]).max() <= 1e-15, (H, W, roi)
        assert abs(val.value - (np.abs(gx) + np.abs(gy)).sum()) <= 1e-12, (H, W, roi)


def test_torch_op_restatement_matches_reference_and_analytic_oracle(gold):
    """oracle/spec_eklt_torch.py (the reference's torch ops + autograd; what the CPU arm of bench.py times) against the
    reference goldens, all levels, regimes and switch variants; and against the analytic numpy oracle on fresh inputs."""
    import torch

    from oracle import spec_eklt_torch as TT

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()
    wd, wtv, wp = (float(x) for x in gold["cost_weights"])
    planes = [t(gold[k]) for k in ("grad_x", "grad_y", "measured", "weight_inverse")]
    for scale, (patch, ph, pw) in enumerate(gold["levels_t"], 1):
        for name in ("start", "random", "far"):
            key = f"L{scale}_{name}"
            loss, grad = TT.value_and_grad(t(gold[key + "_theta"]), *planes, gold["roi_t"], patch, wd, wtv, wp)
            assert abs(loss - float(gold[key + "_loss"])) <= 1e-13, key
            assert np.abs(grad.numpy() - gold[key + "_grad"]).max() <= 1e-12 * np.abs(gold[key + "_grad"]).max(), key
    for c in _variant_cases(gold):
        loss, grad = TT.value_and_grad(t(c["theta"]), planes[0], planes[1], t(c["measured"]), t(c["winv"]), gold["roi_t"],
                                       c["patch"], *c["cost_weights"], poisson=c["poisson"], warp=c["warp"],
                                       no_polarity=c["no_polarity"],
                                       weights=None if c["weights"] is None else t(c["weights"]))
        assert abs(loss - c["loss"]) <= 1e-13, c["name"]
        assert np.abs(grad.numpy() - c["grad"]).max() <= 1e-12 * np.abs(c["grad"]).max(), c["name"]
    rng = np.random.default_rng(21)
    H, W, patch, roi = 45, 77, 16, (4, 40, 9, 70)
    ph, pw = E.patch_grid((H, W), patch)
    th = np.concatenate([rng.uniform(-1, 1, (1, ph, pw)), rng.uniform(-2, 2, (2, ph, pw))])
    gx, gy = rng.normal(size=(2, H, W)) * 40
    meas = rng.normal(size=(H, W)) * E.roi_mask((H, W), roi)
    meas /= np.linalg.norm(meas)
    winv = rng.uniform(0.05, 1, (H, W))
    r = E.objective(th, gx, gy, meas, winv, roi, patch)
    loss, grad = TT.value_and_grad(t(th), t(gx), t(gy), t(meas), t(winv), roi, patch)
    assert abs(loss - r["loss"]) <= 1e-13
    assert np.abs(grad.numpy() - r["grad"]).max() <= 1e-12 * np.abs(r["grad"]).max()


def test_stored_planes_backward_arithmetic_matches_reference(gold, host_lib):
    """The stored-planes backward (default since round 2): the backward rebuilt from the six planes the forward stores gives the
    reference's gradient at every level and regime, and with event-histogram weights."""
    for scale, (patch, ph, pw) in enumerate(gold["levels_t"], 1):
        for name in ("start", "random", "far"):
            key = f"L{scale}_{name}"
            h = host_value_and_grad(host_lib, gold[key + "_theta"], gold["grad_x"], gold["grad_y"], gold["measured"],
                                    gold["weight_inverse"], gold["roi_t"], patch, gold["cost_weights"], stored=True)
            ref = gold[key + "_grad"]
            assert np.abs(h["grad"] - ref).max() <= 1e-11 * np.abs(ref).max(), key
    for c in _variant_cases(gold):
        if c["no_polarity"] or not c["warp"]:
            continue                                    # the C entry keeps the re-evaluating backward for these
        h = host_value_and_grad(host_lib, c["theta"], gold["grad_x"], gold["grad_y"], c["measured"], c["winv"],
                                gold["roi_t"], c["patch"], c["cost_weights"], poisson=c["poisson"], warp=True,
                                hist_weights=c["weights"], stored=True)
        assert np.abs(h["grad"] - c["grad"]).max() <= 1e-11 * np.abs(c["grad"]).max(), c["name"]


def test_c_abi_argument_validation_without_a_device():
    """The EKLT entry points validate before they touch CUDA: bad arguments come back as EBOS_ERR_* with a message, on a
    machine without a GPU as well (no compute is launched here)."""
    from event_based_bos_b200 import _capi

    lib = _capi.load()
    ws64 = lib.ebos_eklt_workspace_bytes(720, 1280, 12, 20, 64, _capi.EBOS_F64)
    ws32 = lib.ebos_eklt_workspace_bytes(720, 1280, 12, 20, 64, _capi.EBOS_F32)
    plane = 720 * 1280 * 8
    assert 15 * plane < ws64 < 16 * plane and ws32 < ws64          # q, F(2), dF(2), dU(4), stored(6) + small arrays
    assert lib.ebos_eklt_workspace_bytes(0, 1280, 12, 20, 64, _capi.EBOS_F64) == 0
    one = 1 << 12                                   # any non-null "pointer": validation fails before it is dereferenced
    common = lambda **kw: dict(dict(theta=one, flags=3, gx=one, gy=one, meas=one, winv=one, weights=0, H=720, W=1280, ph=12,
                                    pw=20, patch=64, x0=0, x1=720, y0=320, y1=960, dtype=_capi.EBOS_F64, ws=one, ws_bytes=ws64,
                                    loss=one, grad=one), **kw)

    def call(**kw):
        a = common(**kw)
        return lib.ebos_eklt_value_and_grad(a["theta"], a["flags"], a["gx"], a["gy"], a["meas"], a["winv"], a["weights"],
                                            a["H"], a["W"], a["ph"], a["pw"], a["patch"], a["x0"], a["x1"], a["y0"], a["y1"],
                                            1.0, 0.5, 0.1, a["dtype"], a["ws"], a["ws_bytes"], a["loss"], a["grad"], 0)

    assert call(theta=0) == -1 and "null" in _capi.last_error()                       # EBOS_ERR_BAD_ARG
    assert call(flags=8) == -1 and "flag" in _capi.last_error()
    assert call(dtype=7) == -4                                                       # EBOS_ERR_UNSUPPORTED
    assert call(ph=11) == -1 and "patch grid" in _capi.last_error()                  # not ceil(720/64)
    assert call(x1=721) == -1 and "ROI" in _capi.last_error()
    assert call(ws_bytes=ws64 - 1) == -3 and "workspace" in _capi.last_error()       # EBOS_ERR_WORKSPACE
    taps = np.array([1.0, 2.0, 1.0])
    tp = taps.ctypes.data_as(ctypes.c_void_p)
    assert lib.ebos_sepconv2d(one, 8, 8, tp, 3, tp, 2, 0, _capi.EBOS_F64, one, one, 0) == -1       # even tap count
    assert lib.ebos_sepconv2d(one, 8, 8, tp, 3, tp, 3, 2, _capi.EBOS_F64, one, one, 0) == -1       # unknown border
    assert lib.ebos_eklt_upsample(one, 2, 720, 1280, 12, 19, 64, _capi.EBOS_F64, one, 0) == -1     # wrong patch grid


def test_fp32_arithmetic_solve_within_1e3_px_of_reference(gold, host_lib):
    """The float32 instantiation of the kernel arithmetic (serial build) driven through the complete coarse-to-fine Adam
    schedule of the golden `estimate()`: the final flow stays within the north-star bar (1e-3 px RMS) of the float64
    reference.  Evidence for `solver.eklt.precision: "32"`; the GPU tests gate the float64 default."""
    H, W = (int(v) for v in gold["image"])
    n_iter, levels = int(gold["n_iter"]), gold["levels_t"]
    th = gold["solve_x0"].astype(np.float32)
    for li, (patch, ph, pw) in enumerate(levels):
        if li > 0:
            th = E.resize_params(th.astype(np.float64), (ph, pw)).astype(np.float32)
        m, v = np.zeros_like(th), np.zeros_like(th)
        for it in range(E.level_iterations(n_iter, len(levels), li)):
            h = host_value_and_grad(host_lib, th, gold["grad_x"], gold["grad_y"], gold["measured"], gold["weight_inverse"],
                                    gold["roi_t"], patch, gold["cost_weights"], dtype=np.float32)
            g = h["grad"].astype(np.float32)
            m = np.float32(0.9) * m + np.float32(0.1) * g
            v = np.float32(0.999) * v + np.float32(0.001) * g * g
            step = np.float32(0.05 / (1 - 0.9 ** (it + 1)))
            th = th - step * (m / (np.sqrt(v) * np.float32(1 / np.sqrt(1 - 0.999 ** (it + 1))) + np.float32(1e-8)))
    patch = levels[-1][0]
    dense = E.upsample_patch(E.sobel_over_8(th[0].astype(np.float64)), patch, (H, W)) * E.roi_mask((H, W), gold["roi_t"])
    rms = np.sqrt(np.mean((dense - gold["solve_flow"]) ** 2))
    assert rms <= 1e-3, rms


def test_segment_gather_arithmetic_matches_transposed_upsampling(host_lib):
    """The segment-form column gather (default since round 2): the warp/lane-group walk of k_gather_cols_seg + k_gather_rows_thread gives the
    transposed bilinear up-sampling R^T dU C for every even patch size dividing 32, with images that are not multiples of
    anything; the serial build also checks the alignment claim (one floor cell per lane group)."""
    rng = np.random.default_rng(12)
    for (H, W, patch) in [(37, 53, 8), (40, 100, 16), (21, 70, 4), (50, 70, 32), (19, 33, 2), (720 // 8, 1280 // 8, 8)]:
        ph, pw = E.patch_grid((H, W), patch)
        geo = E.upsample_geometry((H, W), ph, pw, patch)
        dims = (ctypes.c_int * 9)(H, W, ph, pw, patch, 0, H, 0, W)
        dU = rng.normal(size=(4, H, W))
        T1 = np.zeros((4, H, pw + 2))
        dPad = np.zeros((4, ph + 2, pw + 2))
        rc = host_lib.eklt_host_gather_seg(dims, 4, _p(dU), _p(T1), _p(dPad))
        assert rc == 0, (H, W, patch, rc)

        def tri(n_out, offset, n_cells):          # [cells, pixels] triangle weights of the padded axis
            u = (np.arange(n_out) + offset + 0.5) / patch - 0.5
            return np.clip(1.0 - np.abs(u[None, :] - np.arange(n_cells)[:, None]), 0.0, None)

        R, C = tri(H, geo["h1"], ph + 2), tri(W, geo["w1"], pw + 2)
        ref = np.einsum("ai,cij,bj->cab", R, dU, C)
        assert np.abs(dPad - ref).max() <= 1e-12 * np.abs(ref).max(), (H, W, patch)
        # and it is the adjoint the oracle uses (after folding the replicate padding)
        folded = np.zeros((4, ph, pw))
        rr = np.clip(np.arange(ph + 2) - 1, 0, ph - 1)
        cc = np.clip(np.arange(pw + 2) - 1, 0, pw - 1)
        np.add.at(folded, (slice(None), rr[:, None], cc[None, :]), dPad)
        assert np.abs(folded - E.upsample_patch_adjoint(dU, patch, ph, pw)).max() <= 1e-12 * np.abs(ref).max()
