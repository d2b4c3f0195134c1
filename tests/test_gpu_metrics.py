"""Rows f-3 / f-4 of SURVEY.md section 8 on the GPU: flow-error metrics (`ebos_flow_error`) and the Gaussian-blurred
IWE (`ebos_blur3`, forward + exact adjoint; numpy branch through `ebos_sepconv2d`) against goldens produced by the
unmodified reference (tests/golden/reference_metrics_v1.npz, oracle/make_golden_metrics.py) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_metrics_v1.npz")
KEYS = ("EPE", "1PE", "2PE", "3PE", "5PE", "10PE", "20PE", "AE")


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def _vec(d):
    return np.array([float(d[k]) for k in KEYS], dtype=np.float64)


def test_flow_error_matches_reference(g):
    import event_based_bos_b200 as ebos

    for name in list(g["cases"]) + ["inf64"]:
        gt, pred, mask = g[f"{name}/gt"], g[f"{name}/pred"], g.get(f"{name}/mask")
        tol = 2e-6 if gt.dtype == np.float32 else 1e-12
        res = ebos.utils.calculate_flow_error_numpy(gt, pred, mask)
        assert tuple(res.keys()) == KEYS and all(isinstance(v, float) or np.isscalar(v) for v in res.values())
        np.testing.assert_allclose(_vec(res), g[f"{name}/numpy"], rtol=tol, atol=0, equal_nan=True, err_msg=name)
        if name == "inf64":
            continue
        res_t = ebos.utils.calculate_flow_error_tensor(
            torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(),
            event_mask=None if mask is None else torch.from_numpy(mask).cuda(),
            time_scale=torch.from_numpy(g[f"{name}/time_scale"]).cuda())
        assert all(isinstance(v, torch.Tensor) and v.is_cuda and v.dim() == 0 for v in res_t.values())
        np.testing.assert_allclose(_vec(res_t), g[f"{name}/tensor"], rtol=max(tol, 2e-7), atol=0, equal_nan=True, err_msg=name)
    # the NaN that upstream produces for an infinite ground-truth value is kept, not "fixed"
    res = ebos.utils.calculate_flow_error_numpy(g["inf64/gt"], g["inf64/pred"])
    assert np.isnan(res["EPE"]) and np.isnan(res["AE"]) and not np.isnan(res["1PE"])
    with pytest.raises(AssertionError):
        ebos.utils.calculate_flow_error_numpy(g["inf64/gt"][0], g["inf64/pred"][0])


def test_solver_flow_error_with_event_mask(g):
    """SolverBase.calculate_flow_error (src/solver/base.py:289-317): event mask from the imager, cropped to the ROI."""
    import event_based_bos_b200 as ebos

    ev = g["evmask/events"]
    H, W = g["evmask/mask"].shape[-2:]
    imager = ebos.EventImageConverter((H, W))
    assert np.array_equal(imager.create_eventmask(ev), g["evmask/mask"])
    assert np.array_equal(imager.create_eventmask(torch.from_numpy(ev).cuda()).cpu().numpy(), g["evmask/mask"])
    roi = {"xmin": 4, "xmax": 26, "ymin": 6, "ymax": 40}
    cfg = {"filter": {"filters": None, "parameters": roi}, "outer_padding": 0}
    slv = ebos.solver.SolverBase((H, W), (22, 34), {}, cfg, None)
    rng = np.random.default_rng(3)
    gt = rng.uniform(-4, 4, (2, 22, 34))
    pred = gt + rng.normal(0, 1.5, gt.shape)
    got = slv.calculate_flow_error(pred, gt, events=ev, roi=roi)
    ref = spec.flow_error(gt[None], pred[None], g["evmask/mask"][:, 4:26, 6:40])
    np.testing.assert_allclose(_vec(got), _vec(ref), rtol=1e-12)
    got = slv.calculate_flow_error(pred, gt)
    np.testing.assert_allclose(_vec(got), _vec(spec.flow_error(gt[None], pred[None])), rtol=1e-12)


def test_flow_error_full_size_vs_oracle():
    import event_based_bos_b200 as ebos

    rng = np.random.default_rng(9)
    B, H, W = 2, 720, 1280
    gt = rng.uniform(-25, 25, (B, 2, H, W))
    gt[:, :, ::7, ::5] = 0.0
    pred = gt + rng.normal(0, 3.0, gt.shape)
    mask = rng.uniform(size=(B, 1, H, W)) > 0.5
    for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-6)):
        got = ebos.utils.calculate_flow_error_numpy(gt.astype(dt), pred.astype(dt), mask)
        ref = spec.flow_error(gt.astype(dt), pred.astype(dt), mask)
        np.testing.assert_allclose(_vec(got), _vec(ref), rtol=tol)


def test_blurred_iwe_matches_reference(g):
    import event_based_bos_b200 as ebos
    from event_based_bos_b200 import ops

    for name in g["blur_cases"]:
        ev = torch.from_numpy(g[f"{name}/events"]).cuda().requires_grad_()
        H, W = g[f"{name}/iwe"].shape
        sigma = float(g[f"{name}/sigma"])
        f32 = ev.dtype == torch.float32
        tol = 2e-6 if f32 else 1e-13
        imager = ebos.EventImageConverter((H, W))
        img = imager.create_image_from_events_tensor(ev, "bilinear_vote", sigma=sigma)
        scale = np.abs(g[f"{name}/blurred"]).max()
        assert img.shape == (H, W)
        assert np.abs(img.detach().cpu().numpy() - g[f"{name}/blurred"]).max() <= tol * scale, name
        (img * torch.from_numpy(g[f"{name}/probe"]).cuda()).sum().backward()
        ge = g[f"{name}/grad_events"]
        assert np.abs(ev.grad.cpu().numpy() - ge).max() <= (2e-5 if f32 else 1e-12) * np.abs(ge).max(), name
        # the adjoint kernel alone against autograd through torchvision's gaussian_blur
        I = torch.from_numpy(g[f"{name}/iwe"]).cuda().requires_grad_()
        (ops.blur3(I, sigma) * torch.from_numpy(g[f"{name}/probe"]).cuda()).sum().backward()
        adj = g[f"{name}/adjoint"]
        assert np.abs(I.grad.cpu().numpy() - adj).max() <= tol * np.abs(adj).max(), name
    # <A x, y> = <x, A^T y> on an odd shape, batched
    x = torch.randn(3, 1, 37, 53, device="cuda", dtype=torch.float64)
    y = torch.randn_like(x)
    yg = y.clone()
    xr = x.clone().requires_grad_()
    (ops.blur3(xr, 1.3) * y).sum().backward()
    lhs = float((ops.blur3(x, 1.3) * y).sum())
    assert abs(lhs - float((x * xr.grad).sum())) <= 1e-12 * abs(lhs)
    with pytest.raises(ValueError):
        ops.blur3(torch.zeros(1, 5, device="cuda"), 1.0)


def test_numpy_branch_gaussian_matches_scipy():
    """numpy branch, sigma > 0: scipy.ndimage.gaussian_filter semantics (4-sigma truncation, 'reflect', every axis)."""
    from scipy.ndimage import gaussian_filter

    import event_based_bos_b200 as ebos

    rng = np.random.default_rng(4)
    H, W, n = 40, 56, 4000
    ev = np.stack([rng.uniform(0, H - 1, n), rng.uniform(0, W - 1, n), np.sort(rng.uniform(0, 1, n)), rng.integers(0, 2, n)], 1)
    imager = ebos.EventImageConverter((H, W))
    plain = imager.create_image_from_events_numpy(ev, "bilinear_vote", sigma=0)
    for sigma in (1, 2.5):
        got = imager.create_image_from_events_numpy(ev, "bilinear_vote", sigma=sigma)
        assert np.abs(got - gaussian_filter(plain, sigma)).max() <= 1e-12 * np.abs(plain).max()
    pol = imager.create_image_from_events_numpy(ev, "polarity", sigma=0)
    got = imager.create_image_from_events_numpy(ev, "polarity", sigma=1)
    assert got.shape == (2, H, W)
    assert np.abs(got - gaussian_filter(pol, 1)).max() <= 1e-12 * np.abs(pol).max()   # upstream also blurs ACROSS the two channels


def test_cost_registry_matches_reference(g):
    """The reference's whole cost registry through HybridCost with the weights of configs/hot_plate1.yaml:
    total, per-term history and gradients against the reference's autograd."""
    import event_based_bos_b200 as ebos

    names = [str(n) for n in g["costs/names"]]
    cww = dict(zip(names, (float(v) for v in g["costs/weight_values"])))
    assert set(names) <= set(ebos.costs.functions)
    hybrid = ebos.costs.HybridCost("minimize", cww, store_history=True)
    assert set(hybrid.required_keys) == {"prediction", "measurement", "flow", "omit_boundary", "pxy"}
    pred = torch.from_numpy(g["costs/prediction"]).cuda().requires_grad_()
    flow = torch.from_numpy(g["costs/flow"]).cuda().requires_grad_()
    pxy = torch.from_numpy(g["costs/pxy"]).cuda().requires_grad_()
    arg = {"prediction": pred, "measurement": torch.from_numpy(g["costs/measurement"]).cuda(),
           "weights": torch.from_numpy(g["costs/weights"]).cuda(), "flow": flow, "omit_boundary": False, "pxy": pxy}
    loss = hybrid.calculate(arg)
    loss.backward()
    assert abs(float(loss) - float(g["costs/loss"])) <= 1e-12 * abs(float(g["costs/loss"]))
    hist = hybrid.get_history()
    np.testing.assert_allclose([hist[k][0] for k in names], g["costs/terms"], rtol=1e-12)
    assert len(hist["loss"]) == 1
    for t, key in ((pred, "prediction"), (flow, "flow"), (pxy, "pxy")):
        ref = g[f"costs/grad_{key}"]
        assert np.abs(t.grad.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max(), key
    # numpy inputs give floats (upstream's numpy branches never flip the sign)
    got = [ebos.costs.functions["diff_norm"]("minimize").calculate({"prediction": g["costs/prediction"],
                                                                     "measurement": g["costs/measurement"], "weights": None}),
           ebos.costs.functions["flow_norm_pxy"]("minimize").calculate({"pxy": g["costs/pxy"]}),
           ebos.costs.functions["flow_norm"]("maximize").calculate({"flow": g["costs/flow"]})]
    np.testing.assert_allclose(got, g["costs/numpy_terms"], rtol=1e-12)
    with pytest.raises(KeyError):
        ebos.costs.functions["diff_norm"]().calculate({"prediction": pred, "measurement": pred})   # `weights` is read (:42)
    with pytest.raises(KeyError):
        ebos.costs.HybridCost("minimize", {"no_such_cost": 1.0})
    hybrid.update_weight({k: 2.0 for k in names})
    assert hybrid.cost_func[names[0]]["weight"] == 2.0 and hybrid.cost_func[names[0]]["func"].name == names[0]


@pytest.mark.parametrize("cost", ["gradient_magnitude", "image_variance"])
def test_fused_objective_on_blurred_iwe_vs_oracle(cost):
    """`iwe.blur_sigma` inside the fused entries (blur kernel -> cost -> adjoint blur -> backward): value and gradient
    against autograd through the oracle's restatement of create_iwe(..., sigma) (fp32 1e-5, fp64 1e-11), eager, as a
    captured graph, and through the solver (`solver.iwe.blur_sigma`)."""
    from event_based_bos_b200 import ops, solver

    H, W, n, sigma = 40, 56, 30000, 1.5
    ev = torch.from_numpy(spec.synthetic_events(n, (H, W), seed=6))
    flow = torch.from_numpy(spec.synthetic_flow((H, W), seed=6, max_val=4.0))
    for dt, tol in ((torch.float32, 1e-5), (torch.float64, 1e-11)):
        for omit, pad in ((False, 0), (True, 2)):
            kw = dict(cost=cost, tv_weight=0.5, data_weight=1.0, omit_boundary=omit, outer_padding=(pad, pad), blur_sigma=sigma)
            ref_loss, ref_grad = spec.cmax_value_and_grad(ev.to(dt), flow.to(dt), (H, W), **kw)
            win = ops.PreparedWindow(ev.to(dt).cuda(), (H, W), "first", True, dtype=dt)
            ws = ops.CmaxWorkspace(H, W, (pad, pad), "cuda", dt)
            for _ in range(2):     # twice: the clean-workspace protocol must hold with the blur planes in play
                loss, grad = ops.cmax_value_and_grad(win, flow.to(dt).cuda(), cost, 1.0, 0.5, None, omit, (pad, pad), ws,
                                                     blur_sigma=sigma)
                assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss)), (dt, omit)
                err = float((grad.cpu().double() - ref_grad.double()).abs().max() / ref_grad.double().abs().max())
                assert err <= tol, (dt, omit, err)
            cap = ops.CmaxGraph(win, flow.to(dt).cuda(), cost, 1.0, 0.5, None, omit, (pad, pad), blur_sigma=sigma)
            l2, g2 = cap.replay()
            assert abs(float(l2) - float(ref_loss)) <= tol * abs(float(ref_loss))
            assert float((g2.cpu().double() - ref_grad.double()).abs().max() / ref_grad.double().abs().max()) <= tol
    # solver: 20 iterations with the blurred objective in fp64 against the oracle loop
    cfg = {"outer_padding": 0, "warp_direction": "first", "iwe": {"method": "bilinear_vote", "blur_sigma": sigma},
           "optimizer": {"method": "Adam", "n_iter": 20},
           "cmax": {"cost_with_weight": {cost: 1.0, "image_gradient": 0.5}, "lr": 0.05, "precision": "64"}}
    flow0 = np.random.default_rng(1).uniform(-1, 1, (2, H, W))
    evd = ev.double()
    x0 = torch.from_numpy(flow0.copy()).requires_grad_()      # (a copy: Adam updates the leaf in place)
    opt = torch.optim.Adam([x0], lr=0.05)
    for _ in range(20):
        opt.zero_grad()
        spec.cmax_loss(evd, x0, (H, W), cost=cost, tv_weight=0.5, blur_sigma=sigma).backward()
        opt.step()
    for fused in (True, False):
        cfg["cmax"]["fused"] = fused
        slv = solver.collections["contrast_maximization"]((H, W), (H, W), {}, cfg, None)
        got = slv.estimate(evd.numpy(), flow0=flow0)
        assert float(np.sqrt(np.mean((got - x0.detach().numpy()) ** 2))) <= 1e-8, fused
