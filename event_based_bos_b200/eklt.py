"""Functional layer of the EKLT inner loop: tensors in, tensors out, every numeric step a kernel of
libebos.so (csrc/ebos_eklt.cu).  There is no CPU path.

Reference interfaces mirrored (paths into the reference tree):
  patch_flow                 PatchEkltDependent.poisson_to_flow                    src/solver/patch_eklt_dependent.py:259-281
  upsample                   PatchEklt.interpolate_dense_flow_from_patch_tensor    src/solver/patch_eklt.py:173-204
  EkltLevel.value_and_grad   PatchEkltPyramid2._objective_scipy + loss.backward()  src/solver/patch_eklt_pyramid2.py:267-285, 368-392
  EkltLevel.solve            the Adam loop of run_estimation_per_scale             src/solver/patch_eklt_pyramid2.py:253-288
  frame_gradients            GenerativeMaximumLikelihood._set_frame                src/solver/generative_max_likelihood.py:194-213
  measurement_and_weights    PatchEklt.calculate_iwe_cache + _make_measured_increment   src/solver/patch_eklt.py:271-304
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import check, current_stream, ptr
from .ops import _check_cuda, dtype_code


# ---- geometry (host integers only) ------------------------------------------------------------------------------
def patch_grid(image_size: Tuple[int, int], patch: int) -> Tuple[int, int]:
    """Patch-image shape: centres `arange(0, H, patch) + patch/2` (src/solver/patch_eklt_pyramid2.py:83-110)."""
    H, W = image_size
    return (-(-H // patch), -(-W // patch))


def pyramid_levels(image_size: Tuple[int, int], coarsest: int = 64, finest: int = 8) -> List[Tuple[int, int, int]]:
    """[(patch, ph, pw)] coarse to fine (src/solver/patch_eklt_pyramid2.py:53-81)."""
    n = int(np.log2(coarsest / finest)) + 2
    return [(coarsest // (2 ** (i - 1)),) + patch_grid(image_size, coarsest // (2 ** (i - 1))) for i in range(1, n)]


# ---- small operators ----------------------------------------------------------------------------------------------
def patch_flow(intensity: torch.Tensor) -> torch.Tensor:
    """Sobel/8 with replicate padding of the intensity patch grid [ph,pw] -> [2,ph,pw] (row, column derivative)."""
    _check_cuda(intensity)
    p = intensity.contiguous()
    if p.dim() != 2:
        raise ValueError(f"intensity must be [ph,pw], got {tuple(p.shape)}")
    out = torch.empty((2,) + tuple(p.shape), dtype=p.dtype, device=p.device)
    check(_capi.load().ebos_eklt_patch_flow(ptr(p), p.shape[0], p.shape[1], dtype_code(p), ptr(out), current_stream()),
          "ebos_eklt_patch_flow")
    return out


def upsample(patch_values: torch.Tensor, patch: int, image_size: Tuple[int, int]) -> torch.Tensor:
    """[C,ph,pw] -> [C,H,W]: replicate pad 1, bilinear x`patch`, centre crop (the reference's 24-row offset at
    720 rows included)."""
    _check_cuda(patch_values)
    p = patch_values.contiguous()
    if p.dim() != 3:
        raise ValueError(f"patch_values must be [C,ph,pw], got {tuple(p.shape)}")
    H, W = int(image_size[0]), int(image_size[1])
    if tuple(p.shape[1:]) != patch_grid((H, W), patch):
        raise ValueError(f"patch grid {tuple(p.shape[1:])} does not match image {H}x{W} at patch size {patch}")
    out = torch.empty((p.shape[0], H, W), dtype=p.dtype, device=p.device)
    check(_capi.load().ebos_eklt_upsample(ptr(p), p.shape[0], H, W, p.shape[1], p.shape[2], int(patch), dtype_code(p),
                                          ptr(out), current_stream()), "ebos_eklt_upsample")
    return out


def sepconv2d(image: torch.Tensor, taps_rows: Sequence[float], taps_cols: Sequence[float], border: str = "reflect101"
              ) -> torch.Tensor:
    """Separable correlation with mirrored borders ('reflect101' = cv2 default, 'reflect' = scipy.ndimage)."""
    _check_cuda(image)
    img = image.contiguous()
    if img.dim() != 2:
        raise ValueError(f"image must be [H,W], got {tuple(img.shape)}")
    modes = {"reflect101": 0, "reflect": 1}
    if border not in modes:
        raise ValueError(f"border must be one of {sorted(modes)}")
    tr = np.ascontiguousarray(taps_rows, dtype=np.float64)
    tc = np.ascontiguousarray(taps_cols, dtype=np.float64)
    tmp, out = torch.empty_like(img), torch.empty_like(img)
    check(_capi.load().ebos_sepconv2d(ptr(img), img.shape[0], img.shape[1], tr.ctypes.data_as(ctypes.c_void_p), len(tr),
                                      tc.ctypes.data_as(ctypes.c_void_p), len(tc), modes[border], dtype_code(img),
                                      ptr(tmp), ptr(out), current_stream()), "ebos_sepconv2d")
    return out


def gaussian_taps_cv2(sigma: float) -> np.ndarray:
    """Taps of cv2.GaussianBlur(ksize=None, sigmaX=sigma) on a float64 image: ksize = round(8 sigma + 1) | 1."""
    k = int(round(sigma * 8 + 1)) | 1
    x = np.arange(k) - (k - 1) / 2.0
    w = np.exp(-0.5 * x * x / (sigma * sigma))
    return w / w.sum()


def gaussian_taps_scipy(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """Taps of scipy.ndimage.gaussian_filter(sigma): radius int(truncate * sigma + 0.5)."""
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return w / w.sum()


def frame_gradients(frame: torch.Tensor, use_log_intensity: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row derivative, column derivative) of the frame: cv2.Sobel(ksize=3) with its reflect-101 border."""
    f = torch.log(frame + 1) if use_log_intensity else frame
    d, s = (-1.0, 0.0, 1.0), (1.0, 2.0, 1.0)
    return sepconv2d(f, d, s), sepconv2d(f, s, d)


def polarity_histogram(events: torch.Tensor, image_size: Tuple[int, int], no_polarity: bool = False) -> torch.Tensor:
    """`create_iwe(events, method="polarity")[0] - [1]` on the device (their sum with `no_polarity`): the positive and the
    negative events (p > 0 / not) are voted bilinearly with the numpy branch's arithmetic (float64, floor bias 1e-8;
    src/event_image_converter.py:355-363, 503-560) by the `ebos_iwe_splat` kernel.  src/solver/patch_eklt.py:277-281."""
    from . import ops
    from .event_image_converter import NUMPY_FLOOR_BIAS

    _check_cuda(events)
    ev = events.to(torch.float64)
    pos = ev[:, 3] > 0
    H, W = int(image_size[0]), int(image_size[1])
    planes = []
    for sel in (pos, ~pos):
        part = ev[sel].contiguous()
        planes.append(ops.iwe_splat(part, (H, W), (0, 0), 1.0, False, NUMPY_FLOOR_BIAS) if part.shape[0] > 0
                      else torch.zeros((H, W), dtype=torch.float64, device=ev.device))
    return planes[0] + planes[1] if no_polarity else planes[0] - planes[1]


def measurement_and_weights(histogram: torch.Tensor, roi: Tuple[int, int, int, int], iwe_sigma: float = 2.0,
                            weight_inverse: bool = True, inverse_sigma: float = 10.0, weight_sigma: float = 0.0
                            ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """From the polarity histogram (positive minus negative IWE; their sum with no_polarity) to
    (measured increment * ROI mask, weight_inverse, event-histogram weights * ROI mask or None).

    weight_sigma > 0 (weight_loss_by_event_hist): weights = GaussianBlur(|hist|, weight_sigma) are multiplied into the
    blurred histogram before the normalisation and returned for the prediction (src/solver/patch_eklt.py:283-286,
    src/solver/patch_eklt_pyramid2.py:333-337, :262-263).

    cache_histogram = GaussianBlur(hist, sigma) / ||.||_F;  weight_inverse = 1 - 0.95 * clip(G10(|hist|)) / max
    (src/solver/patch_eklt.py:283-304).  The reductions (norm, mean, std, max) are one-off torch reductions on the
    device; the filters are the ebos_sepconv2d kernel."""
    _check_cuda(histogram)
    h = histogram.contiguous()
    if iwe_sigma:
        g = gaussian_taps_cv2(float(iwe_sigma))
        blurred = sepconv2d(h, g, g, "reflect101")
    else:
        blurred = h.clone()
    weights = None
    if weight_sigma:
        gw = gaussian_taps_cv2(float(weight_sigma))
        weights = sepconv2d(h.abs(), gw, gw, "reflect101")
        blurred = weights * blurred
    meas = blurred / torch.linalg.norm(blurred)
    mask = torch.zeros_like(meas)
    mask[roi[0]:roi[1], roi[2]:roi[3]] = 1
    if weight_inverse:
        s = gaussian_taps_scipy(float(inverse_sigma))
        wi = sepconv2d(h.abs(), s, s, "reflect")
        # (bounds as device tensors: a Python-float bound would be a host read-back in the middle of the window's queue)
        wi = torch.minimum(torch.clamp(wi, min=0), wi.mean() + wi.std(unbiased=False) / 2.0)
        wi = 1.0 - 0.95 * (wi / wi.max())
    else:
        wi = torch.ones_like(h)
    return meas * mask, wi, (None if weights is None else weights * mask)


# ---- one pyramid level --------------------------------------------------------------------------------------------
class EkltProblem:
    """The per-window constants of the objective, resident on the device."""

    def __init__(self, grad_x: torch.Tensor, grad_y: torch.Tensor, measured: torch.Tensor, weight_inverse: torch.Tensor,
                 roi: Tuple[int, int, int, int], cost_weights: Tuple[float, float, float] = (1.0, 0.5, 0.1),
                 poisson: bool = True, warp: bool = True, no_polarity: bool = False,
                 weights: Optional[torch.Tensor] = None):
        """`poisson`, `warp`, `no_polarity`, `weights` mirror generative_ml.poisson_model / optimize_warp / no_polarity /
        weight_loss_by_event_hist (hot_plate1: True, True, False, None); they fix the layout of theta:
        [(1 if poisson else 2) + (2 if warp else 0), ph, pw]."""
        _check_cuda(grad_x, grad_y, measured, weight_inverse, weights)
        self.dtype = grad_x.dtype
        self.code = dtype_code(grad_x)
        self.H, self.W = (int(v) for v in grad_x.shape)
        for name, t in (("grad_y", grad_y), ("measured", measured), ("weight_inverse", weight_inverse),
                        ("weights", weights)):
            if t is None:
                continue
            if tuple(t.shape) != (self.H, self.W) or t.dtype != self.dtype:
                raise ValueError(f"{name} must be [{self.H},{self.W}] {self.dtype}, got {tuple(t.shape)} {t.dtype}")
        self.grad_x, self.grad_y = grad_x.contiguous(), grad_y.contiguous()
        self.measured, self.weight_inverse = measured.contiguous(), weight_inverse.contiguous()
        self.weights = None if weights is None else weights.contiguous()
        self.flags = ((_capi.EKLT_POISSON if poisson else 0) | (_capi.EKLT_WARP if warp else 0)
                      | (_capi.EKLT_NO_POLARITY if no_polarity else 0))
        self.channels = (1 if poisson else 2) + (2 if warp else 0)
        x0, x1, y0, y1 = (int(v) for v in roi)
        if not (0 <= x0 <= x1 <= self.H and 0 <= y0 <= y1 <= self.W):
            raise ValueError(f"roi {roi} outside the {self.H}x{self.W} image")
        self.roi = (x0, x1, y0, y1)
        self.w_data, self.w_tv, self.w_pxy = (float(v) for v in cost_weights)
        self.device = grad_x.device

    def level(self, patch: int) -> "EkltLevel":
        return EkltLevel(self, int(patch))

    def update_(self, grad_x: torch.Tensor, grad_y: torch.Tensor, measured: torch.Tensor, weight_inverse: torch.Tensor,
                weights: Optional[torch.Tensor] = None) -> None:
        """Refresh the planes IN PLACE for the next window (same shapes, dtype and switches), so that levels and CUDA
        graphs built on this problem stay valid."""
        if (weights is None) != (self.weights is None):
            raise ValueError("update_ cannot add or remove the event-histogram weights plane")
        for dst, src in ((self.grad_x, grad_x), (self.grad_y, grad_y), (self.measured, measured),
                         (self.weight_inverse, weight_inverse), (self.weights, weights)):
            if dst is not None:
                dst.copy_(src)


class EkltLevel:
    """Objective of one pyramid level: theta [C,ph,pw] = (intensity | v_row, v_col) (+ p_row, p_col)."""

    def __init__(self, problem: EkltProblem, patch: int):
        self.p = problem
        self.patch = patch
        self.ph, self.pw = patch_grid((problem.H, problem.W), patch)
        lib = _capi.load()
        self.ws_bytes = int(lib.ebos_eklt_workspace_bytes(problem.H, problem.W, self.ph, self.pw, patch, problem.code))
        if self.ws_bytes == 0:
            raise RuntimeError("ebos_eklt_workspace_bytes rejected the geometry")
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=problem.device)
        self.loss = torch.zeros(1, dtype=problem.dtype, device=problem.device)
        self.grad = torch.zeros((problem.channels, self.ph, self.pw), dtype=problem.dtype, device=problem.device)

    def _check_theta(self, theta: torch.Tensor) -> None:
        _check_cuda(theta)
        c = self.p.channels
        if tuple(theta.shape) != (c, self.ph, self.pw) or theta.dtype != self.p.dtype or not theta.is_contiguous():
            raise ValueError(f"theta must be a contiguous [{c},{self.ph},{self.pw}] {self.p.dtype} tensor, "
                             f"got {tuple(theta.shape)} {theta.dtype}")

    def _common_args(self):
        p = self.p
        return (p.flags, ptr(p.grad_x), ptr(p.grad_y), ptr(p.measured), ptr(p.weight_inverse), ptr(p.weights), p.H, p.W,
                self.ph, self.pw, self.patch, *p.roi, p.w_data, p.w_tv, p.w_pxy, p.code, ptr(self.workspace),
                self.ws_bytes)

    def value_and_grad(self, theta: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(loss [1], dL/dtheta like theta); both alias buffers of this level (overwritten by the next call)."""
        self._check_theta(theta)
        check(_capi.load().ebos_eklt_value_and_grad(ptr(theta), *self._common_args(), ptr(self.loss), ptr(self.grad),
                                                    current_stream()), "ebos_eklt_value_and_grad")
        return self.loss, self.grad

    def loss_terms(self) -> dict:
        """Un-weighted terms of the LAST evaluation (device -> host read; diagnostics)."""
        # acc sits behind the 256-byte aligned TV accumulators (csrc/ebos_eklt.cu: carve)
        off = ((_capi.ACC_DOUBLES * 8 + 255) // 256) * 256
        acc = self.workspace[off:off + 64 * 8].view(torch.float64).cpu()
        # csrc/ebos_eklt.cu: [3] loss [4] data [5] tv [6] pxy; [16:32] spread slots of sum q^2
        return {"norm": math.sqrt(float(acc[16:32].sum())), "data": float(acc[4]), "tv": float(acc[5]),
                "pxy": float(acc[6]), "loss": float(acc[3])}

    def adam_iteration(self, theta: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                       step_dev: torch.Tensor, lr: float = 0.05, betas: Tuple[float, float] = (0.9, 0.999),
                       eps: float = 1e-8) -> torch.Tensor:
        """One solver iteration (objective, gradient, Adam update of `theta` in place); returns the loss BEFORE it."""
        self._check_theta(theta)
        _check_cuda(exp_avg, exp_avg_sq, step_dev)
        check(_capi.load().ebos_eklt_adam_iteration(
            ptr(theta), *self._common_args(), ptr(self.loss), ptr(self.grad), ptr(exp_avg), ptr(exp_avg_sq), float(lr),
            float(betas[0]), float(betas[1]), float(eps), ptr(step_dev), current_stream()), "ebos_eklt_adam_iteration")
        return self.loss

    def solve(self, theta0: torch.Tensor, n_iter: int, lr: float = 0.05, cuda_graph: bool = True,
              history: Optional[list] = None, replay=None) -> torch.Tensor:
        """`n_iter` Adam iterations from theta0; returns the FINAL iterate (upstream's `best_x` aliases the leaf).
        With `cuda_graph`, 10 iterations are captured once and replayed; a loss `history` forces eager mode.  `replay`:
        an `ops.ReplaySlot` that outlives this level (the solver keeps one per pyramid level and stream slot), so that
        its executable is updated for the new window instead of being built and destroyed (see ops.ReplaySlot)."""
        from . import ops

        theta = theta0.to(device=self.p.device, dtype=self.p.dtype).contiguous().clone()
        self._check_theta(theta)
        m, v = torch.zeros_like(theta), torch.zeros_like(theta)
        step_dev = torch.zeros(1, dtype=torch.int32, device=theta.device)
        if n_iter <= 0:
            return theta

        def iteration():
            self.adam_iteration(theta, m, v, step_dev, lr)

        if not cuda_graph or history is not None or n_iter < 4:
            for _ in range(n_iter):
                iteration()
                if history is not None:
                    history.append(float(self.loss[0]))
            return theta
        unroll = next(u for u in (10, 8, 6, 5, 4, 3, 2, 1) if n_iter % u == 0)
        backup = theta.clone()
        if replay is None:
            replay = ops.ReplaySlot()
        cur = torch.cuda.current_stream()

        def run():
            iteration()                      # warm-up outside capture
            theta.copy_(backup)
            m.zero_()
            v.zero_()
            step_dev.zero_()
            replay.capture(lambda: [iteration() for _ in range(unroll)])
            for _ in range(n_iter // unroll):
                replay.launch()

        if cur == torch.cuda.default_stream(theta.device):
            # the default stream cannot be captured: one side stream per process
            global _SIDE
            if _SIDE is None or _SIDE.device != theta.device:
                _SIDE = torch.cuda.Stream(device=theta.device)
            _SIDE.wait_stream(cur)
            with torch.cuda.stream(_SIDE):
                run()
            cur.wait_stream(_SIDE)
        else:
            run()
        self._keep = (replay, m, v, step_dev, backup)
        return theta


_SIDE = None


def resize_params(theta: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """Seed of a finer level: torchvision `resize` of the coarser result (src/solver/patch_eklt_pyramid2.py:243-246) =
    bilinear, align_corners=False.  A [3,h,w] -> [3,2h(-1),2w(-1)] interpolation of at most 90x160 values, once per
    level: done with torch's own interpolate on the device (host-side plumbing, not part of the iteration)."""
    return torch.nn.functional.interpolate(theta[None], size=tuple(int(v) for v in out_hw), mode="bilinear",
                                           align_corners=False)[0].contiguous()
