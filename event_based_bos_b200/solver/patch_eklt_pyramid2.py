"""`PatchEkltPyramid2` -- the solver configs/hot_plate1.yaml selects (`method: patch_eklt_pyramid2`), with the inner
loop on the GPU.

Mirrors the constructor contract, config keys, start values, level schedule and return convention of
src/solver/patch_eklt_pyramid2.py (and of its bases patch_eklt_dependent.py, patch_eklt.py,
generative_max_likelihood.py).  Supported `generative_ml` switches -- every combination of

    poisson_model, optimize_warp, no_polarity, weight_loss_by_event_hist, weight_loss_by_inverse_event_hist,
    use_log_intensity, model_image in {current, black, background}

with `cost_with_weight` drawn from {diff_norm, image_gradient, flow_norm_pxy} and `optimizer.method: Adam` (the shipped
config: poisson + warp, all three costs).  What selects a DIFFERENT algorithm upstream (angle model, scipy / optuna
optimisers, other cost terms, 5x5 Sobel) raises NotImplementedError instead of silently computing something else.
Visualisation / video hooks of the upstream class are not reproduced.
"""
import logging
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _capi, eklt
from .base import SolverBase

logger = logging.getLogger(__name__)

AVAILABLE_MODEL_IMAGES = ["background", "current", "black"]
SUPPORTED_COSTS = ("diff_norm", "image_gradient", "flow_norm_pxy")


def _flag(cfg: dict, key: str) -> bool:
    return key in cfg and bool(cfg[key])


class PatchEkltPyramid2(SolverBase):
    """Pyramidal (coarse-to-fine) EKLT estimation; patches of 64 -> 8 px (src/solver/patch_eklt_pyramid2.py:22-51)."""

    def __init__(self, orig_image_shape: tuple, crop_image_shape: tuple, calibration_parameter: dict = {},
                 solver_config: dict = {}, visualize_module=None) -> None:
        super().__init__(orig_image_shape, crop_image_shape, calibration_parameter, solver_config, visualize_module)
        self._frame = None
        self._opt_config = self.slv_config["optimizer"]
        self._opt_method = self._opt_config["method"]
        self._gml_config = self.slv_config["generative_ml"]
        assert self._gml_config["model_image"] in AVAILABLE_MODEL_IMAGES, \
            f"the setting 'mode_image' must be in {AVAILABLE_MODEL_IMAGES}."
        self.is_angle_model = _flag(self._gml_config, "angle_model")
        self.is_poisson_model = _flag(self._gml_config, "poisson_model")
        self.do_weight_inverse = _flag(self._gml_config, "weight_loss_by_inverse_event_hist")
        self.cost_weight = self.slv_config["cost_with_weight"]
        unsupported = []
        if self._opt_method != "Adam":
            unsupported.append(f"optimizer.method={self._opt_method!r} (only 'Adam')")
        if self.is_angle_model:
            unsupported.append("generative_ml.angle_model (asserted against upstream as well: pyramid2:300)")
        self.optimize_warp = _flag(self._gml_config, "optimize_warp")
        self.no_polarity = _flag(self._gml_config, "no_polarity")
        self.weight_by_hist = _flag(self._gml_config, "weight_loss_by_event_hist")
        if _flag(self._gml_config, "px-py_as-angle-magnitude"):
            unsupported.append("px-py_as-angle-magnitude is optuna-only upstream")
        if self._gml_config.get("sobel_ksize", 3) != 3:
            unsupported.append("generative_ml.sobel_ksize must be 3")
        extra = sorted(set(self.cost_weight) - set(SUPPORTED_COSTS))
        if extra:
            unsupported.append(f"cost terms {extra} (supported: {SUPPORTED_COSTS})")
        if unsupported:
            raise NotImplementedError("event_based_bos_b200.PatchEkltPyramid2 does not implement: " + "; ".join(unsupported))
        if "flow_norm_pxy" in self.cost_weight and not self.optimize_warp:
            # upstream: FlowNormPxy.calculate fails on the missing "pxy" key at the first iteration (costs/base.py)
            raise KeyError("pxy: cost term flow_norm_pxy needs generative_ml.optimize_warp")
        ekc = self.slv_config.get("eklt", {}) or {}
        self._dtype = torch.float32 if str(ekc.get("precision", "64")) == "32" else torch.float64
        self.use_cuda_graph = bool(ekc.get("cuda_graph", True))
        self.store_history = bool(ekc.get("store_history", False))
        self.history = {}
        self._staging = {}
        self._replays = {}          # (stream, pyramid level) -> ops.ReplaySlot, kept across windows
        self._streams = []
        self.last_many_stats = {}
        self.levels = eklt.pyramid_levels(tuple(self.orig_image_shape), 64, 8)
        self.coarest_scale, self.finest_scale = 1, len(self.levels) + 1        # upstream's spelling
        self.iter_cnt = 0
        self.estimate_mask_dense_numpy = np.zeros(self.orig_image_shape)
        self.estimate_mask_dense_numpy[self.crop_xmin:self.crop_xmax, self.crop_ymin:self.crop_ymax] = 1

    # -- per-window preprocessing ---------------------------------------------------------------------------------
    def _set_frame(self, frame: np.ndarray) -> None:
        """Frame gradients (src/solver/generative_max_likelihood.py:194-213)."""
        _capi.require_device()
        self._frame = frame
        f = torch.as_tensor(np.asarray(frame), device="cuda").to(self._dtype)
        self._gradient_x_torch, self._gradient_y_torch = eklt.frame_gradients(
            f, use_log_intensity=_flag(self._gml_config, "use_log_intensity"))

    def calculate_iwe_cache(self, events: np.ndarray) -> None:
        """Polarity histogram -> measured increment and weight_inverse (src/solver/patch_eklt.py:271-304)."""
        _capi.require_device()
        ev = torch.as_tensor(np.asarray(events), device="cuda")
        hist = eklt.polarity_histogram(ev, tuple(self.orig_image_shape), self.no_polarity).to(self._dtype)
        roi = (self.crop_xmin, self.crop_xmax, self.crop_ymin, self.crop_ymax)
        self.cache_measured, self.weight_inverse, self.cache_weights = eklt.measurement_and_weights(
            hist, roi, iwe_sigma=self._gml_config["iwe_sigma"], weight_inverse=self.do_weight_inverse,
            weight_sigma=self._gml_config["weight_sigma"] if self.weight_by_hist else 0.0)

    def _initialize_velocity(self) -> np.ndarray:
        """src/solver/generative_max_likelihood.py:425-450: random intensity (one np.random draw) or zero velocity,
        plus a zero translation with optimize_warp."""
        head = [np.random.random() * 2.0 - 1] if self.is_poisson_model else [0.0, 0.0]
        return np.array(head + ([0.0, 0.0] if self.optimize_warp else []), dtype=np.float64)

    def _start_parameters(self, level: int, previous: Optional[torch.Tensor]) -> torch.Tensor:
        """x0 of a level (src/solver/patch_eklt_pyramid2.py:225-248), drawing from np.random exactly like upstream:
        one draw to measure the parameter dimension, then -- at the coarsest level -- one per patch; the per-patch
        tuples are concatenated and reshaped to [n_dim,ph,pw] as upstream does (which interleaves them)."""
        _, ph, pw = self.levels[level]
        self.n_parameter_dim = len(self._initialize_velocity())
        if level == 0:
            x0 = np.concatenate([self._initialize_velocity() for _ in range(ph * pw)]).reshape(
                (self.n_parameter_dim, ph, pw))
            return torch.as_tensor(x0, device="cuda").to(self._dtype)
        return eklt.resize_params(previous, (ph, pw))

    # -- main entry --------------------------------------------------------------------------------------------------
    def estimate(self, events: np.ndarray, *args, **kwargs) -> np.ndarray:
        """events [n,4] (row, col, t, p) + `frame=` (and `background=`)  ->  dense flow [2,H,W] float64, zero outside
        the ROI (src/solver/patch_eklt_pyramid2.py:129-190)."""
        return self._enqueue(events, **kwargs).cpu().numpy()

    def estimate_many(self, windows: Sequence[np.ndarray], frames: Optional[Sequence[np.ndarray]] = None,
                      backgrounds: Optional[Sequence[np.ndarray]] = None, concurrency: int = 4) -> List[np.ndarray]:
        """`estimate` for a sequence of independent windows (upstream keeps no state between windows apart from the
        `np.random` stream of the start values, which is drawn here in window order exactly as consecutive `estimate`
        calls would) with up to `concurrency` solves in flight on separate CUDA streams.  One evaluation of this
        objective is ten short launches that leave most of the GPU idle (ncu: issue slots ~50 %, DRAM 6-22 %), so
        overlapping windows raises windows/s without touching any window's result.  Rolling schedule as in
        `ContrastMaximizationDense.estimate_many`: a slot's whole solve is queued in one go, the result lands in the
        slot's pinned staging buffer and is collected when the slot comes round again."""
        n_slots = max(1, min(int(concurrency), len(windows)))
        if self.store_history:
            n_slots = 1          # the loss history is read back iteration by iteration
        kw = lambda i: {k: v[i] for k, v in (("frame", frames), ("background", backgrounds)) if v is not None}
        if n_slots == 1:
            return [self.estimate(w, **kw(i)) for i, w in enumerate(windows)]
        _capi.require_device()
        device = torch.device("cuda", torch.cuda.current_device())
        if self._frame is None and self._gml_config["model_image"] == "background":
            self._set_frame(backgrounds[0])       # shared by every window: ready before the slots start
            torch.cuda.current_stream().synchronize()
        while len(self._streams) < n_slots:      # slot streams (and their executables) live as long as the solver
            self._streams.append(torch.cuda.Stream(device=device))
        streams = self._streams
        results: List[Optional[np.ndarray]] = [None] * len(windows)
        busy: List[Optional[tuple]] = [None] * n_slots
        t_plan = t_wait = 0.0

        def retire(slot: int) -> None:
            nonlocal t_wait
            idx, stage, done, _keep = busy[slot]
            t0 = time.perf_counter()
            done.synchronize()
            t_wait += time.perf_counter() - t0
            results[idx] = stage.numpy().copy()
            busy[slot] = None

        for idx, events in enumerate(windows):
            slot = idx % n_slots
            if busy[slot] is not None:
                retire(slot)
            t0 = time.perf_counter()
            with torch.cuda.stream(streams[slot]):
                out = self._enqueue(events, **kw(idx))
                stage = self._staging.get(slot)
                if stage is None or stage.shape != out.shape:
                    stage = self._staging[slot] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                stage.copy_(out, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
            busy[slot] = (idx, stage, done, out)
            t_plan += time.perf_counter() - t0
        for slot in sorted((s for s in range(n_slots) if busy[s] is not None), key=lambda s: busy[s][0]):
            retire(slot)
        k = 1e3 / len(windows)
        self.last_many_stats = {"windows": len(windows), "slots": n_slots, "host_plan_and_queue_ms": t_plan * k,
                                "host_wait_for_gpu_ms": t_wait * k}
        return results

    def _enqueue(self, events: np.ndarray, **kwargs) -> torch.Tensor:
        """Queue one window's whole solve on the current CUDA stream; returns the device tensor [2,H,W] float64 the
        masked flow will be in."""
        if self._gml_config["model_image"] == "current":
            self._set_frame(kwargs["frame"])
        elif self._gml_config["model_image"] == "black":
            self._set_frame(np.zeros_like(kwargs["frame"]))
        elif self._frame is None and self._gml_config["model_image"] == "background":
            self._set_frame(kwargs["background"])
        self.calculate_iwe_cache(events)
        roi = (self.crop_xmin, self.crop_xmax, self.crop_ymin, self.crop_ymax)
        weights = tuple(float(self.cost_weight.get(k, 0.0)) for k in SUPPORTED_COSTS)
        planes = (self._gradient_x_torch, self._gradient_y_torch, self.cache_measured, self.weight_inverse)
        problem = eklt.EkltProblem(*planes, roi, weights, poisson=self.is_poisson_model, warp=self.optimize_warp,
                                   no_polarity=self.no_polarity, weights=self.cache_weights)
        theta = None
        self.best_params_per_scale = {}
        self.history = {}
        for li, (patch, ph, pw) in enumerate(self.levels):
            scale = li + 1
            logger.info(f"Scale {scale}, patch num {(ph, pw)}, patch shape {(patch, patch)}")
            x0 = self._start_parameters(li, theta)
            iters = self._opt_config["n_iter"] // (self.finest_scale - scale + 1)
            hist = [] if self.store_history else None
            key = (torch.cuda.current_stream().cuda_stream, li)
            if key not in self._replays:
                from .. import ops
                self._replays[key] = ops.ReplaySlot()
            theta = problem.level(patch).solve(x0, iters, lr=0.05, cuda_graph=self.use_cuda_graph, history=hist,
                                               replay=self._replays[key])
            self.best_params_per_scale[scale] = theta
            if hist is not None:
                self.history[scale] = hist
        patch = self.levels[-1][0]
        patch_flow = eklt.patch_flow(theta[0]) if self.is_poisson_model else theta[:2].contiguous()
        dense = eklt.upsample(patch_flow, patch, tuple(self.orig_image_shape))
        self.iter_cnt += 1
        mask = torch.zeros(tuple(self.orig_image_shape), dtype=torch.float64, device=dense.device)
        mask[self.crop_xmin:self.crop_xmax, self.crop_ymin:self.crop_ymax] = 1
        return dense.double() * mask
