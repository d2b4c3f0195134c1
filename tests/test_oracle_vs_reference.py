"""The restated CPU oracle (oracle/spec.py) against the LIVE reference on fresh random inputs.

Runs only where the reference tree is mounted (the build container: /root/reference, override with
EBOS_REFERENCE_ROOT); on the GPU box the committed goldens (tests/test_oracle_golden.py) pin the same functions.
Bit-exact for the warp, the tap indices / masks / values and the IWE; 1e-12 for the float64 scalars."""
import numpy as np
import pytest
import torch

from oracle import ref_import, spec

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


@pytest.mark.parametrize("seed", [11, 12, 13])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_warp_and_vote_bit_exact_vs_live_reference(ref, seed, dtype):
    rng = np.random.default_rng(seed)
    H, W, n, pad = 30 + seed, 44 + seed, 3000, seed % 3
    ev = torch.from_numpy(np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.02, n)),
                                    rng.integers(0, 2, n)], 1)).to(dtype)
    flow = torch.from_numpy(rng.uniform(-6, 6, (2, H, W))).to(dtype)
    for direction in ("first", "middle", "last", 0.3, "before", "after"):
        warped, _ = ref.warp.Warp((H, W), normalize_t=True).warp_event(ev, flow, "dense-flow", direction=direction)
        mine = spec.warp_dense_flow(ev, flow, (H, W), direction, True)
        assert torch.equal(warped, mine), direction
    imager = ref.event_image_converter.EventImageConverter((H, W), outer_padding=pad)
    iwe = imager.create_image_from_events_tensor(warped, "bilinear_vote", sigma=0)
    assert torch.equal(iwe, spec.bilinear_vote(mine, (H, W), (pad, pad)))
    wts = torch.from_numpy(rng.uniform(0.2, 2.0, n)).to(dtype)
    assert torch.equal(imager.bilinear_vote_tensor(warped, weight=wts), spec.bilinear_vote(mine, (H, W), (pad, pad), weight=wts))
    blurred = imager.create_image_from_events_tensor(warped, "bilinear_vote", sigma=2)
    mine_b = spec.gaussian_blur3(spec.bilinear_vote(mine, (H, W), (pad, pad)), 2.0)
    assert float((blurred - mine_b).abs().max()) <= (1e-5 if dtype == torch.float32 else 1e-13) * float(blurred.abs().max())


def test_total_variation_and_loop_vs_live_reference(ref):
    rng = np.random.default_rng(5)
    H, W = 21, 33
    flow = torch.from_numpy(rng.uniform(-3, 3, (2, H, W))).requires_grad_()
    wts = torch.from_numpy(rng.uniform(0.1, 2.0, (H, W)))
    cost = ref.costs.ImageGradient(direction="minimize")
    loss = cost.calculate({"flow": flow, "omit_boundary": False, "weights": wts})
    loss.backward()
    mine = flow.detach().clone().requires_grad_()
    ml = spec.total_variation(mine, wts)
    ml.backward()
    assert abs(float(loss) - float(ml)) <= 1e-14 * abs(float(loss))
    assert float((flow.grad - mine.grad).abs().max()) <= 1e-15
    assert float((spec.total_variation_grad(flow.detach(), wts) - flow.grad).abs().max()) <= 1e-15


def test_flow_error_vs_live_reference(ref):
    rng = np.random.default_rng(0)
    gt, pred = rng.uniform(-30, 30, (2, 2, 16, 20)), rng.uniform(-30, 30, (2, 2, 16, 20))
    gt[0, :, 3, 4] = 0.0
    mask = rng.uniform(size=(2, 1, 16, 20)) > 0.3
    for m in (None, mask):
        a = spec.flow_error(gt, pred, m)
        b = ref.utils.calculate_flow_error_numpy(gt, pred, m)
        assert a.keys() == b.keys()
        for k in a:
            assert a[k] == pytest.approx(b[k], rel=1e-13), k
        ts = rng.uniform(0.5, 2, (2, 1))
        a = spec.flow_error(gt, pred, m, ts, tensor_variant=True)
        b = ref.utils.calculate_flow_error_tensor(torch.from_numpy(gt), torch.from_numpy(pred),
                                                  None if m is None else torch.from_numpy(m), torch.from_numpy(ts))
        for k in a:
            assert a[k] == pytest.approx(float(b[k]), rel=2e-7), k
