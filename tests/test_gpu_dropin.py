"""The drop-in classes used the way the reference's call sites use them:
numpy arrays, CPU tensors and CUDA tensors in; same container out; reference goldens as the answer."""
import numpy as np
import pytest
import torch

import event_based_bos_b200 as ebos
from oracle import spec

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_reference_composition_with_tensors(golden):
    name = "c0_f32_first"
    H, W = 48, 64
    ev, flow = torch.from_numpy(golden[f"{name}/events"]), torch.from_numpy(golden[f"{name}/flow"])
    warper = ebos.Warp((H, W), normalize_t=True)
    for det in (False, True):
        imager = ebos.EventImageConverter((H, W), deterministic=det)
        for dev in ("cpu", "cuda"):
            warped, feat = warper.warp_event(ev.to(dev), flow.to(dev), "dense-flow", direction="first")
            assert warped.device.type == dev and warped.shape == (len(ev), 4)
            assert set(feat) == {"determinant", "trace", "divergence", "straint", "absement"}
            assert all(v["value"] is None for v in feat.values())
            assert np.array_equal(warped.cpu().numpy(), golden[f"{name}/warped"])
            iwe = imager.create_iwe(warped, method="bilinear_vote", sigma=0)
            assert iwe.device.type == dev and iwe.shape == (H, W)
            if det:
                assert np.array_equal(iwe.cpu().numpy(), golden[f"{name}/iwe"])
            else:
                assert rel_err(iwe.cpu().numpy(), golden[f"{name}/iwe"]) <= 1e-5
            mask = imager.create_eventmask(warped)
            assert mask.shape == (1, H, W) and np.array_equal(mask.cpu().numpy(), golden[f"{name}/mask"])


def test_numpy_inputs_follow_the_numpy_branch(golden):
    H, W = 32, 48
    ev, flow = golden["numpy/events"], golden["numpy/flow"]
    warped, _ = ebos.Warp((H, W), normalize_t=True).warp_event(ev, flow, "dense-flow", "first")
    assert isinstance(warped, np.ndarray) and np.array_equal(warped, golden["numpy/warped"])
    imager = ebos.EventImageConverter((H, W), deterministic=True)
    assert np.array_equal(imager.create_iwe(warped, "bilinear_vote", sigma=0), golden["numpy/iwe_sigma0"])
    np.testing.assert_allclose(imager.create_iwe(warped, "bilinear_vote", sigma=1), golden["numpy/iwe_sigma1"], rtol=1e-12)
    assert np.array_equal(imager.create_iwe(warped, "polarity", sigma=0), golden["numpy/iwe_polarity"])
    assert np.array_equal(imager.create_eventmask(warped), golden["numpy/eventmask"])
    atomic = ebos.EventImageConverter((H, W)).create_iwe(warped, "bilinear_vote", sigma=0)
    assert atomic.dtype == np.float64 and rel_err(atomic, golden["numpy/iwe_sigma0"]) <= 1e-12


def test_blur_padding_weights_batch_2dof(golden):
    ev = torch.from_numpy(golden["weighted/events"]).cuda()
    imager = ebos.EventImageConverter((48, 64), deterministic=True)
    for sigma in (1, 3):
        out = imager.create_iwe(ev, "bilinear_vote", sigma=sigma)
        np.testing.assert_allclose(out.cpu().numpy(), golden[f"sigma/iwe_sigma{sigma}"], rtol=2e-6, atol=1e-7)
    wt = torch.from_numpy(golden["weighted/weight"]).cuda()
    out = imager.create_image_from_events_tensor(ev, "bilinear_vote", weight=wt, sigma=0)
    assert np.array_equal(out.cpu().numpy(), golden["weighted/iwe"])
    # padded imager: golden c3 (outer_padding 3)
    H, W, pad, _ = (int(v) for v in golden["c3_f32_frac_pad/meta"])
    pimg = ebos.EventImageConverter((H, W), outer_padding=pad, deterministic=True)
    assert pimg.image_size == (H + 2 * pad, W + 2 * pad)
    out = pimg.create_iwe(torch.from_numpy(golden["c3_f32_frac_pad/warped"]).cuda(), "bilinear_vote", sigma=0)
    assert np.array_equal(out.cpu().numpy(), golden["c3_f32_frac_pad/iwe"])
    # batched events x batched flows
    evb, flb = torch.from_numpy(golden["batched/events"]).cuda(), torch.from_numpy(golden["batched/flow"]).cuda()
    wb, _ = ebos.Warp((32, 48), normalize_t=True).warp_event(evb, flb, "dense-flow", "middle")
    assert np.array_equal(wb.cpu().numpy(), golden["batched/warped"])
    ib = ebos.EventImageConverter((32, 48), deterministic=True).create_iwe(wb, "bilinear_vote", sigma=0)
    assert np.array_equal(ib.cpu().numpy(), golden["batched/iwe"])
    # 2-dof translation
    w2, _ = ebos.Warp((32, 48), normalize_t=True).warp_event(torch.from_numpy(golden["nonorm/events"]).cuda(),
                                                             torch.from_numpy(golden["twodof/theta"]).cuda(),
                                                             "2d-translation", "first")
    assert np.array_equal(w2.cpu().numpy(), golden["twodof/warped"])
    fl = ebos.Warp((6, 8)).get_flow_from_motion(np.array([1.5, -2.0]), "2d-translation")
    assert fl.shape == (2, 6, 8) and np.allclose(fl[0], -1.5) and np.allclose(fl[1], 2.0)


def test_costs_and_hybrid_through_autograd(golden):
    """loss.backward() through Warp -> EventImageConverter -> HybridCost, like a reference solver would."""
    for name in golden["comp_cases"]:
        if golden[f"{name}/events"].dtype != np.float32:
            continue
        H, W, omit, tvw, pad = golden[f"{name}/cfg"]
        H, W, pad, omit = int(H), int(W), int(pad), bool(omit)
        kind = str(golden[f"{name}/kind"])
        ev = torch.from_numpy(golden[f"{name}/events"]).cuda()
        flow = torch.from_numpy(golden[f"{name}/flow"]).cuda().requires_grad_()
        warped, _ = ebos.Warp((H, W), normalize_t=True).warp_event(ev, flow, "dense-flow", direction="first")
        iwe = ebos.EventImageConverter((H, W), outer_padding=pad).create_iwe(warped, "bilinear_vote", sigma=0)
        cost = ebos.costs.HybridCost("minimize", {kind: 1.0, "image_gradient": float(tvw)}, store_history=True)
        loss = cost.calculate({"iwe": iwe, "flow": flow, "weights": 1.0, "omit_boundary": omit})
        loss.backward()
        assert abs(float(loss) - float(golden[f"{name}/loss"])) <= 1e-5 * abs(float(golden[f"{name}/loss"])), name
        assert rel_err(flow.grad.cpu().numpy(), golden[f"{name}/grad"]) <= 1e-5, name
        hist = cost.get_history()
        assert len(hist["loss"]) == 1 and len(hist[kind]) == 1 and len(hist["image_gradient"]) == 1
    # direction handling of the individual costs
    img = torch.rand(20, 30).cuda()
    var_min = ebos.costs.functions["image_variance"]("minimize").calculate({"iwe": img, "omit_boundary": False})
    var_nat = ebos.costs.functions["image_variance"]("natural").calculate({"iwe": img, "omit_boundary": False})
    assert float(var_min) == -float(var_nat) and abs(float(var_nat) - float(torch.var(img))) < 1e-6
    with pytest.raises(KeyError):
        ebos.costs.functions["gradient_magnitude"]().calculate({"omit_boundary": False})


def test_error_behaviour():
    flow = torch.zeros(2, 4, 4).cuda()
    warper = ebos.Warp((4, 4), normalize_t=True)
    with pytest.raises(RuntimeError, match="out of bounds"):
        warper.warp_event(torch.tensor([[7.0, 0, 0, 0], [0, 0, 1.0, 0]]).cuda(), flow, "dense-flow")
    with pytest.raises(ValueError):
        warper.warp_event(torch.zeros(3, 4).cuda(), flow, "dense-flow", direction=2)
    with pytest.raises(ebos.warp.MotionModelKeyError):
        warper.warp_event(torch.zeros(3, 4).cuda(), flow, "homography")
    with pytest.raises(RuntimeError):
        ebos.EventImageConverter((4, 4)).create_iwe([1, 2, 3])
    with pytest.raises(NotImplementedError):
        ebos.EventImageConverter((4, 4)).create_iwe(torch.zeros(3, 4).cuda(), method="polarity")
    with pytest.raises(RuntimeError):  # upstream's tensor 'count' is broken and raises
        ebos.EventImageConverter((4, 4)).create_iwe(torch.zeros(3, 4).cuda(), method="count")


def test_image_gradient_weight_shapes_broadcast_like_upstream():
    """`weights` may be anything that broadcasts against the [2,H,W] flow gradient upstream
    (src/costs/image_gradient.py:69-70): [H,W], [1,H,W] (a `mask[None]` ROI weight), per-channel [2,H,W], scalars."""
    import event_based_bos_b200 as ebos
    from oracle import spec

    rng = np.random.default_rng(21)
    H, W = 19, 27
    flow = torch.from_numpy(rng.uniform(-3, 3, (2, H, W)))
    cost = ebos.costs.functions["image_gradient"]("minimize")
    for w in (torch.from_numpy(rng.uniform(0.1, 2, (H, W))), torch.from_numpy(rng.uniform(0.1, 2, (1, H, W))),
              torch.from_numpy(rng.uniform(0.1, 2, (2, H, W))), torch.from_numpy(rng.uniform(0.1, 2, (2, 1, 1))), 0.7):
        f = flow.clone().cuda().requires_grad_()
        loss = cost.calculate({"flow": f, "omit_boundary": False, "weights": w.cuda() if isinstance(w, torch.Tensor) else w})
        loss.backward()
        fr = flow.clone().requires_grad_()
        ref = torch.mean(torch.abs(torch.gradient(fr, dim=1)[0] * w) + torch.abs(torch.gradient(fr, dim=2)[0] * w))
        ref.backward()
        assert abs(float(loss) - float(ref)) <= 1e-12 * abs(float(ref))
        assert float((f.grad.cpu() - fr.grad).abs().max()) <= 1e-14
