"""The CPU oracle (oracle/spec.py) against outputs of the unmodified reference (tests/golden/).
This is what pins the oracle: the reference itself has no tests for the path."""
import os

import numpy as np
import pytest
import torch

from oracle import spec
from tests.conftest import parse_direction


def _case(golden, name):
    H, W, pad, _ = golden[f"{name}/meta"]
    return int(H), int(W), int(pad), parse_direction(str(golden[f"{name}/direction"]))


def test_warp_and_vote_bit_exact(golden):
    for name in golden["warp_cases"]:
        H, W, pad, direction = _case(golden, name)
        ev = torch.from_numpy(golden[f"{name}/events"])
        flow = torch.from_numpy(golden[f"{name}/flow"])
        warped = spec.warp_dense_flow(ev, flow, (H, W), direction, normalize_t=True)
        assert np.array_equal(warped.numpy(), golden[f"{name}/warped"]), name
        inds, mask, vals = spec.vote_taps(warped[:, :2], (H + 2 * pad, W + 2 * pad), (pad, pad))
        assert np.array_equal(inds.numpy(), golden[f"{name}/inds"]), name
        assert np.array_equal(vals.numpy(), golden[f"{name}/vals"], equal_nan=True), name
        for sequential in (False, True):
            iwe = spec.bilinear_vote(warped, (H, W), (pad, pad), sequential=sequential)
            assert np.array_equal(iwe.numpy(), golden[f"{name}/iwe"]), (name, sequential)
        assert np.array_equal((iwe != 0).numpy()[None], golden[f"{name}/mask"]), name


def test_warp_not_normalised_and_2dof(golden):
    ev = torch.from_numpy(golden["nonorm/events"])
    warped = spec.warp_dense_flow(ev, torch.from_numpy(golden["nonorm/flow"]), (32, 48), "first", normalize_t=False)
    assert np.array_equal(warped.numpy(), golden["nonorm/warped"])
    w2 = spec.warp_2dof(ev, torch.from_numpy(golden["twodof/theta"]), "first", normalize_t=True)
    assert np.array_equal(w2.numpy(), golden["twodof/warped"])


def test_batched_equals_rows(golden):
    ev, flow = torch.from_numpy(golden["batched/events"]), torch.from_numpy(golden["batched/flow"])
    for b in range(2):
        warped = spec.warp_dense_flow(ev[b], flow[b], (32, 48), "middle", True)
        assert np.array_equal(warped.numpy(), golden["batched/warped"][b])
        assert np.array_equal(spec.bilinear_vote(warped, (32, 48)).numpy(), golden["batched/iwe"][b])


def test_weighted_and_blurred(golden):
    ev = torch.from_numpy(golden["weighted/events"])
    iwe = spec.bilinear_vote(ev, (48, 64), weight=torch.from_numpy(golden["weighted/weight"]))
    assert np.array_equal(iwe.numpy(), golden["weighted/iwe"])
    base = spec.bilinear_vote(ev, (48, 64))
    for sigma in (1, 3):
        blurred = spec.gaussian_blur3(base, sigma)
        np.testing.assert_allclose(blurred.numpy(), golden[f"sigma/iwe_sigma{sigma}"], rtol=1e-6, atol=1e-7)


def test_total_variation(golden):
    for tag in ("f32", "f64"):
        flow = torch.from_numpy(golden[f"tv_{tag}/flow"])
        w = torch.from_numpy(golden[f"tv_{tag}/weights"])
        tol = 1e-6 if tag == "f32" else 1e-13
        np.testing.assert_allclose(spec.total_variation(flow, w).numpy(), golden[f"tv_{tag}/loss"], rtol=tol)
        atol = tol * np.abs(golden[f"tv_{tag}/grad"]).max()  # sums of +-terms: tolerance relative to the gradient scale
        np.testing.assert_allclose(spec.total_variation_grad(flow, w).numpy(), golden[f"tv_{tag}/grad"], rtol=tol, atol=atol)
        np.testing.assert_allclose(spec.total_variation(flow, 1.0).numpy(), golden[f"tv_{tag}/loss_w1"], rtol=tol)
        np.testing.assert_allclose(spec.total_variation_grad(flow, 1.0).numpy(), golden[f"tv_{tag}/grad_w1"], rtol=tol, atol=atol)


def test_composed_loss_and_gradient(golden):
    for name in golden["comp_cases"]:
        H, W, omit, tvw, pad = golden[f"{name}/cfg"]
        H, W, pad, omit = int(H), int(W), int(pad), bool(omit)
        kind = str(golden[f"{name}/kind"])
        ev, flow = torch.from_numpy(golden[f"{name}/events"]), torch.from_numpy(golden[f"{name}/flow"])
        loss, grad = spec.cmax_value_and_grad(ev, flow, (H, W), cost=kind, tv_weight=float(tvw), omit_boundary=omit,
                                              outer_padding=(pad, pad))
        f64 = flow.dtype == torch.float64
        np.testing.assert_allclose(loss.numpy(), golden[f"{name}/loss"], rtol=1e-12 if f64 else 2e-6, err_msg=name)
        scale = np.abs(golden[f"{name}/grad"]).max()
        np.testing.assert_allclose(grad.numpy(), golden[f"{name}/grad"], rtol=1e-9 if f64 else 1e-4,
                                   atol=(1e-14 if f64 else 1e-6) * scale, err_msg=name)
        # the analytic backward used by the CUDA kernels == what autograd derived in the reference
        if f64:
            warped = spec.warp_dense_flow(ev, flow, (H, W))
            iwe = spec.bilinear_vote(warped, (H, W), (pad, pad)).requires_grad_()
            spec.DATA_COSTS[kind](iwe, omit).backward()
            dflow = spec.warp_vote_backward(ev, flow, iwe.grad, (H, W), outer_padding=(pad, pad))
            if tvw:
                dflow = dflow + float(tvw) * spec.total_variation_grad(flow, 1.0)
            np.testing.assert_allclose(dflow.numpy(), golden[f"{name}/grad"], rtol=1e-9, atol=1e-14 * scale, err_msg=name)


def test_adam_solve_matches_reference_loop(golden):
    for tag, tol in (("f64", 1e-9), ("f32", 2e-3)):
        H, W, iters, lr, tvw = golden[f"solve_{tag}/cfg"]
        ev = torch.from_numpy(golden[f"solve_{tag}/events"])
        flow, hist = spec.solve_dense_flow(ev, (int(H), int(W)), int(iters), "gradient_magnitude", float(tvw), float(lr),
                                           return_history=True)
        rms = float(np.sqrt(np.mean((flow.numpy() - golden[f"solve_{tag}/flow"]) ** 2)))
        assert rms <= tol, (tag, rms)
        np.testing.assert_allclose(hist[:5], golden[f"solve_{tag}/history"][:5], rtol=1e-5)


def test_adam_solve_from_tie_free_start(golden):
    """The well-conditioned solve goldens (random initial flow): this is the 1e-3 px RMS parity gate."""
    for tag, tol in (("f64", 1e-9), ("f64_long", 1e-8), ("f32", 1e-3)):
        H, W, iters, lr, tvw = golden[f"solve_init_{tag}/cfg"]
        ev = torch.from_numpy(golden[f"solve_init_{tag}/events"])
        flow0 = torch.from_numpy(golden[f"solve_init_{tag}/flow0"])
        flow = spec.solve_dense_flow(ev, (int(H), int(W)), int(iters), "gradient_magnitude", float(tvw), float(lr), flow0=flow0)
        rms = float(np.sqrt(np.mean((flow.numpy() - golden[f"solve_init_{tag}/flow"]) ** 2)))
        assert rms <= tol, (tag, rms)
    # conditioning: the reference's own fp32 and fp64 runs agree to < 1e-3 px from this start,
    # but are 1.6e-2 px apart from the all-zero start (exact ties; see oracle/make_golden.py)
    gap_init = float(np.sqrt(np.mean((golden["solve_init_f32/flow"] - golden["solve_init_f64/flow"]) ** 2)))
    gap_zero = float(np.sqrt(np.mean((golden["solve_f32/flow"] - golden["solve_f64/flow"]) ** 2)))
    assert gap_init < 1e-3 < gap_zero


def test_adam_update_equals_torch_optim():
    torch.manual_seed(0)
    p = torch.randn(64, dtype=torch.float64)
    q = p.clone().requires_grad_()
    opt = torch.optim.Adam([q], lr=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 8):
        g = torch.randn(64, dtype=torch.float64)
        q.grad = g.clone()
        opt.step()
        spec.adam_update(p, g, m, v, step)
        np.testing.assert_allclose(p.numpy(), q.detach().numpy(), rtol=1e-12, atol=1e-14)


def test_edge_cases():
    ev = torch.tensor([[1.0, 2.0, 0.5, 1.0], [3.0, 1.0, 0.5, 0.0]])
    flow = torch.ones(2, 4, 4)
    warped = spec.warp_dense_flow(ev, flow, (4, 4))  # single timestamp: 0/0 = NaN, not guarded
    assert torch.isnan(warped[:, :3]).all()
    iwe = spec.bilinear_vote(warped, (4, 4))
    assert torch.isnan(iwe[0, 0]) and torch.isfinite(iwe.reshape(-1)[1:]).all()
    with pytest.raises(ValueError):
        spec.reference_time(ev[:, 2], 1)  # an int direction is rejected like upstream
    with pytest.raises(IndexError):
        spec.warp_dense_flow(torch.tensor([[9.0, 0.0, 0.0, 1.0], [0.0, 0.0, 1.0, 1.0]]), flow, (4, 4))
    # a coordinate in (-1, 0) truncates to pixel 0
    neg = torch.tensor([[-0.5, -0.5, 0.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    assert int(spec.origin_pixel_index(neg[:, 0], neg[:, 1], 4)[0]) == 0


# ---- rows f-3 / f-4: flow-error metrics and the blurred IWE (tests/golden/reference_metrics_v1.npz) ----------------
METRICS_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_metrics_v1.npz")
ERROR_KEYS = ("EPE", "1PE", "2PE", "3PE", "5PE", "10PE", "20PE", "AE")


@pytest.fixture(scope="module")
def metrics_golden():
    with np.load(METRICS_GOLDEN, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def test_flow_error_oracle_matches_reference(metrics_golden):
    g = metrics_golden
    for name in list(g["cases"]) + ["inf64"]:
        mask = g.get(f"{name}/mask")
        got = spec.flow_error(g[f"{name}/gt"], g[f"{name}/pred"], mask)
        tol = 1e-6 if g[f"{name}/gt"].dtype == np.float32 else 1e-13
        np.testing.assert_allclose([got[k] for k in ERROR_KEYS], g[f"{name}/numpy"], rtol=tol, atol=0, equal_nan=True)
        if name != "inf64":
            got = spec.flow_error(g[f"{name}/gt"], g[f"{name}/pred"], mask, g[f"{name}/time_scale"], tensor_variant=True)
            np.testing.assert_allclose([got[k] for k in ERROR_KEYS], g[f"{name}/tensor"], rtol=max(tol, 1e-12), atol=0,
                                       equal_nan=True)
    assert np.isnan(g["inf64/numpy"][0]) and np.isnan(g["inf64/numpy"][7])   # upstream: inf * 0 = nan poisons EPE and AE


def test_blurred_iwe_oracle_matches_reference(metrics_golden):
    g = metrics_golden
    for name in g["blur_cases"]:
        iwe = torch.from_numpy(g[f"{name}/iwe"])
        out = spec.gaussian_blur3(iwe, float(g[f"{name}/sigma"]))
        tol = 1e-6 if iwe.dtype == torch.float32 else 1e-14
        np.testing.assert_allclose(out.numpy(), g[f"{name}/blurred"], rtol=tol, atol=tol * np.abs(g[f"{name}/blurred"]).max())
