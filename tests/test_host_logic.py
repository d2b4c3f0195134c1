"""Host-side logic that needs no GPU: registries, argument mapping, config parsing, utilities."""
import numpy as np
import pytest
import torch

import event_based_bos_b200 as ebos
from event_based_bos_b200 import _capi, ops, sharding, utils
from event_based_bos_b200.solver.contrast_maximization import split_cost_weights


def test_direction_mapping_matches_reference_rules():
    assert ops.direction_code("first") == (_capi.DIR_FIRST, 0.0)
    assert ops.direction_code("last") == (_capi.DIR_LAST, 0.0)
    assert ops.direction_code("middle") == (_capi.DIR_FRAC, 0.5)
    assert ops.direction_code("before") == (_capi.DIR_FRAC, -1.0)
    assert ops.direction_code("after") == (_capi.DIR_FRAC, 2.0)
    assert ops.direction_code(0.25) == (_capi.DIR_FRAC, 0.25)
    kind, frac = ops.direction_code("random")
    assert kind == _capi.DIR_FRAC and 0.0 <= frac <= 1.0
    for bad in (1, "sideways", None):
        with pytest.raises(ValueError):
            ops.direction_code(bad)


def test_registries_and_cost_contract():
    assert set(ebos.costs.functions) == {"image_gradient", "image_variance", "gradient_magnitude", "diff_norm", "flow_norm",
                                         "flow_norm_pxy"}
    assert "contrast_maximization" in ebos.solver.collections
    with pytest.raises(ValueError):
        ebos.costs.functions["image_gradient"](direction="sideways")
    tv = ebos.costs.functions["image_gradient"](store_history=True)
    assert tv.required_keys == ["flow", "omit_boundary"] and tv.get_history() == {"loss": []}
    with pytest.raises(KeyError):  # `weights` is required although upstream does not list it
        tv.calculate({"flow": torch.zeros(2, 4, 4), "omit_boundary": False})
    with pytest.raises(AttributeError):  # upstream has no numpy implementation
        tv.calculate({"flow": np.zeros((2, 4, 4)), "omit_boundary": False, "weights": 1.0})
    hy = ebos.costs.HybridCost("minimize", {"gradient_magnitude": 1.0, "image_gradient": 0.5}, store_history=True)
    assert sorted(hy.required_keys) == ["flow", "iwe", "omit_boundary", "omit_boundary"]
    assert set(hy.get_history()) == {"loss", "gradient_magnitude", "image_gradient"}
    hy.update_weight({"gradient_magnitude": 2.0, "image_gradient": 0.1})
    assert hy.cost_func["image_gradient"]["weight"] == 0.1


def test_warp_class_bookkeeping():
    w = ebos.Warp((6, 8), normalize_t=True)
    assert w.get_key_names("dense-flow") == ["trans_x", "trans_y"]
    assert w.get_key_names("scaler") == ["scaler"]
    with pytest.raises(ebos.warp.MotionModelKeyError):
        w.get_key_names("affine")
    with pytest.raises(ebos.warp.MotionModelKeyError):
        w.warp_event(np.zeros((3, 4)), np.zeros(2), "affine")
    ev = np.array([[0, 0, 1.0, 0], [1, 1, 3.0, 1], [2, 2, 2.0, 0]])
    assert w.calculate_reftime(ev, "first") == 1.0 and w.calculate_reftime(ev, "last") == 3.0
    assert w.calculate_reftime(ev, "middle") == 2.0 and w.calculate_reftime(ev, "before") == -1.0
    np.testing.assert_allclose(w.calculate_dt(ev, 1.0), [0.0, 1.0, 0.5])
    imager = ebos.EventImageConverter((6, 8), outer_padding=2)
    assert imager.image_size == (10, 12) and imager.outer_padding == (2, 2)
    imager.update_property(outer_padding=1)  # upstream quirk: adds p, not 2p
    assert imager.image_size == (11, 13)
    with pytest.raises(NotImplementedError):
        imager.create_image_from_events_tensor(torch.zeros(3, 4), method="polarity")


def test_solver_config_parsing():
    assert split_cost_weights({"gradient_magnitude": 1.0, "image_gradient": 0.5}) == ("gradient_magnitude", 1.0, 0.5)
    assert split_cost_weights({"image_variance": 2.0}) == ("image_variance", 2.0, 0.0)
    for bad in ({"image_gradient": 0.5}, {"image_variance": 1, "gradient_magnitude": 1}, {"image_variance": 1, "diff_norm": 1}):
        with pytest.raises(ValueError):
            split_cost_weights(bad)
    cfg = {"outer_padding": 0, "filter": {"filters": None, "parameters": {"xmin": 0, "xmax": 720, "ymin": 320, "ymax": 960}},
           "warp_direction": "first", "optimizer": {"method": "Adam", "n_iter": 600}}
    slv = ebos.solver.collections["contrast_maximization"]((720, 1280), (720, 640), {}, cfg, None)
    assert slv.n_iter == 600 and slv.lr == 0.05 and slv.data_cost == "gradient_magnitude" and slv.tv_weight == 0.5
    ev = utils.synthetic_events(5000, (720, 1280), seed=0, dtype=np.float64)
    kept, span = slv.preprocess(ev)
    assert (kept[:, 1] >= 320).all() and (kept[:, 1] < 960).all() and span == ev[:, 2].max() - ev[:, 2].min()
    assert slv._roi_mask().sum() == 720 * 640
    with pytest.raises(ValueError):
        ebos.solver.collections["contrast_maximization"]((8, 8), (8, 8), {}, {"optimizer": {"method": "BFGS"}}, None)


def test_utils_generators_and_metrics():
    ev = utils.synthetic_events(1000, (720, 1280), seed=3)
    assert ev.shape == (1000, 4) and ev.dtype == np.float32 and (np.diff(ev[:, 2]) >= 0).all()
    assert ev[:, 0].max() < 720 and ev[:, 1].max() < 1280 and set(np.unique(ev[:, 3])) <= {0.0, 1.0}
    np.random.seed(0)
    g = utils.generate_events(100, 20, 30, 0.0, 0.5)
    assert g.shape == (100, 4) and g[:, 2].max() <= 0.5
    assert utils.generate_uniform_optical_flow((4, 5), 2, 3)[1, 0, 0] == 3
    fl = utils.smooth_flow((64, 96), seed=1)
    assert fl.shape == (2, 64, 96) and abs(np.abs(fl).max() - 3.0) < 1e-5
    bos = utils.synthetic_bos_events(4000, (64, 96), fl, seed=1)
    assert bos.shape == (4000, 4) and (np.diff(bos[:, 2]) >= 0).all()
    # absolute sensor time survives the fp32 cast only after rebasing
    t = 12.0 + np.sort(np.random.default_rng(0).uniform(0, 1 / 120, 1000))
    raw = np.stack([np.zeros(1000), np.zeros(1000), t, np.zeros(1000)], 1)
    reb = utils.rebase_time(raw)
    assert reb[:, 2].min() == 0.0 and len(np.unique(reb[:, 2].astype(np.float32))) > 0.99 * len(np.unique(t))
    # the raw cast quantises to ~1 us at t = 12 s: timestamps collide and dt loses ~3 digits
    assert len(np.unique(raw[:, 2].astype(np.float32))) < len(np.unique(reb[:, 2].astype(np.float32)))
    err_raw = np.abs(raw[:, 2].astype(np.float32).astype(np.float64) - t).max()
    err_reb = np.abs(reb[:, 2].astype(np.float32).astype(np.float64) - (t - t.min())).max()
    assert err_reb < 1e-3 * err_raw


def test_shard_assignment():
    for n, R in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        seen = sorted(w for r in range(R) for w in sharding.shard_windows(n, r, R))
        assert seen == list(range(n))
        sizes = [len(sharding.shard_windows(n, r, R)) for r in range(R)]
        assert max(sizes) - min(sizes) <= 1
        spans = [sharding.shard_events(n, r, R) for r in range(R)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_windows(4, 5, 4)
