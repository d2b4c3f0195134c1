"""Functional layer over the C-ABI: tensors in, tensors out, autograd wired to the analytic
backward kernels.  Everything here runs on a CUDA device through libebos.so; there is no CPU path.

Reference interfaces mirrored (paths into the reference tree):
  warp_dense_flow / warp_2dof      Warp.warp_event                      src/warp.py:193-383
  iwe_splat                        EventImageConverter.bilinear_vote_tensor   src/event_image_converter.py:562-620
  PreparedWindow + cmax_value_and_grad   fused warp->IWE->cost->backward (SURVEY.md section 3b composition)
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _capi
from ._capi import check, current_stream, ptr

Direction = Union[str, float]

COST_KINDS = {"image_variance": _capi.COST_VARIANCE, "gradient_magnitude": _capi.COST_GRADMAG}


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _capi.EBOS_F32
    if t.dtype == torch.float64:
        return _capi.EBOS_F64
    raise TypeError(f"event_based_bos_b200 kernels support float32/float64 tensors, got {t.dtype}")


def direction_code(direction: Direction) -> Tuple[int, float]:
    """Map Warp.calculate_reftime's `direction` (src/warp.py:230-262) to (kind, fraction).
    Like upstream, only a genuine `float` selects the fractional form (an int raises)."""
    if type(direction) is float:
        return _capi.DIR_FRAC, direction
    if direction == "first":
        return _capi.DIR_FIRST, 0.0
    if direction == "middle":
        return _capi.DIR_FRAC, 0.5
    if direction == "last":
        return _capi.DIR_LAST, 0.0
    if direction == "random":
        return _capi.DIR_FRAC, float(np.random.uniform(low=0.0, high=1.0))
    if direction == "before":
        return _capi.DIR_FRAC, -1.0
    if direction == "after":
        return _capi.DIR_FRAC, 2.0
    raise ValueError(f"direction argument should be first, middle, last. Or float. {direction}")


def _as_batched(events: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    if events.dim() == 2:
        return events[None], False
    if events.dim() == 3:
        return events, True
    raise ValueError(f"events must be [n,4] or [b,n,4], got {tuple(events.shape)}")


def _check_cuda(*tensors: Optional[torch.Tensor]) -> None:
    _capi.require_device()
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("event_based_bos_b200.ops expects CUDA tensors (the drop-in classes move inputs for you)")


def time_stats(events: torch.Tensor) -> torch.Tensor:
    """[b,2] (min t, max t) per batch row.  Replaces nt_min/nt_max (src/types/__init__.py:20-47)."""
    ev, _ = _as_batched(events)
    _check_cuda(ev)
    ev = ev.contiguous()
    if ev.shape[1] == 0:
        raise RuntimeError("min()/max() of an empty event array (the reference raises here too)")
    out = torch.empty((ev.shape[0], 2), dtype=ev.dtype, device=ev.device)
    check(_capi.load().ebos_time_stats(ptr(ev), ev.shape[1], ev.shape[0], dtype_code(ev), ptr(out), current_stream()),
          "ebos_time_stats")
    return out


class _WarpDenseFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, events, flow, H, W, dir_kind, dir_frac, normalize_t, validate):
        ev = events.contiguous()
        fl = flow.contiguous()
        b, n = ev.shape[0], ev.shape[1]
        shared_flow = fl.shape[0] == 1 and b > 1
        tstats = time_stats(ev)
        warped = torch.empty_like(ev)
        status = torch.zeros(1, dtype=torch.int32, device=ev.device)
        check(_capi.load().ebos_warp_dense_flow(
            ptr(ev), n, b, ptr(fl), 0 if shared_flow else 2 * H * W, H, W, ptr(tstats), dir_kind, dir_frac,
            int(normalize_t), dtype_code(ev), ptr(warped), ptr(status), current_stream()), "ebos_warp_dense_flow")
        if validate and int(status.item()) & _capi.STATUS_PIXEL_OOB:
            raise RuntimeError("index out of bounds: an event's integer pixel lies outside the flow grid "
                               "(the reference's torch.gather raises for the same input)")
        ctx.save_for_backward(ev, tstats)
        ctx.meta = (H, W, dir_kind, dir_frac, int(normalize_t), fl.shape, shared_flow)
        return warped

    @staticmethod
    def backward(ctx, grad_warped):
        ev, tstats = ctx.saved_tensors
        H, W, dir_kind, dir_frac, normalize_t, flow_shape, shared_flow = ctx.meta
        b, n = ev.shape[0], ev.shape[1]
        gw = grad_warped.contiguous()
        dflow = torch.zeros(flow_shape, dtype=ev.dtype, device=ev.device)
        check(_capi.load().ebos_warp_dense_flow_bwd(
            ptr(ev), n, b, H, W, ptr(tstats), dir_kind, dir_frac, normalize_t, dtype_code(ev), ptr(gw), ptr(dflow),
            0 if shared_flow else 2 * H * W, current_stream()), "ebos_warp_dense_flow_bwd")
        return None, dflow, None, None, None, None, None, None


def warp_dense_flow(events: torch.Tensor, flow: torch.Tensor, image_size: Tuple[int, int],
                    direction: Direction = "first", normalize_t: bool = False, validate: bool = True) -> torch.Tensor:
    """Warped events (x', y', dt, p), same leading shape as `events`; differentiable w.r.t. `flow`."""
    ev, batched = _as_batched(events)
    fl = flow if flow.dim() == 4 else flow[None]
    _check_cuda(ev, fl)
    H, W = int(image_size[0]), int(image_size[1])
    if fl.shape[-3] != 2 or fl.shape[-2] * fl.shape[-1] < 1:
        raise ValueError(f"flow must be [(b,)2,H,W], got {tuple(flow.shape)}")
    if fl.dtype != ev.dtype:
        raise TypeError(f"events ({ev.dtype}) and flow ({fl.dtype}) must share a dtype")
    kind, frac = direction_code(direction)
    # k = trunc(x)*W + trunc(y) uses image_size[1]; the bound is the flow plane's own size, like
    # torch.gather on flow.reshape(b, 2, -1) (src/warp.py:333-336).
    plane = fl.shape[-2] * fl.shape[-1]
    if plane % W:
        raise ValueError(f"flow plane {tuple(fl.shape[-2:])} is not a multiple of image width {W}")
    out = _WarpDenseFlow.apply(ev, fl, plane // W, W, kind, frac, normalize_t, validate)
    return out if batched else out[0]


def warp_2dof(events: torch.Tensor, theta: torch.Tensor, direction: Direction = "first",
              normalize_t: bool = False) -> torch.Tensor:
    """x' = x + dt*theta0, y' = y + dt*theta1 (src/warp.py:344-383).  Forward only."""
    ev, batched = _as_batched(events)
    _check_cuda(ev, theta)
    ev = ev.contiguous()
    th = theta.to(ev.dtype).contiguous()
    kind, frac = direction_code(direction)
    tstats = time_stats(ev)
    out = torch.empty_like(ev)
    check(_capi.load().ebos_warp_2dof(ptr(ev), ev.shape[1], ev.shape[0], ptr(th), ptr(tstats), kind, frac,
                                      int(normalize_t), dtype_code(ev), ptr(out), current_stream()), "ebos_warp_2dof")
    return out if batched else out[0]


class _IweSplat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, events, weight, Hp, Wp, pad_h, pad_w, deterministic, floor_bias):
        ev = events.contiguous()
        w = None if weight is None else weight.contiguous()
        b, n = ev.shape[0], ev.shape[1]
        image = torch.empty((b, Hp, Wp), dtype=ev.dtype, device=ev.device)
        lib = _capi.load()
        mode = 1 if deterministic else 0
        ws_bytes = lib.ebos_splat_workspace_bytes(n, b, Hp, Wp, dtype_code(ev), mode)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ev.device) if ws_bytes else None
        check(lib.ebos_iwe_splat(ptr(ev), n, b, Hp, Wp, pad_h, pad_w, ptr(w), floor_bias, dtype_code(ev), mode,
                                 ptr(image), 0, 0, ptr(ws), ws_bytes, current_stream()), "ebos_iwe_splat")
        ctx.save_for_backward(ev, w)
        ctx.meta = (Hp, Wp, pad_h, pad_w, floor_bias)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        ev, w = ctx.saved_tensors
        Hp, Wp, pad_h, pad_w, floor_bias = ctx.meta
        b, n = ev.shape[0], ev.shape[1]
        g = grad_image.contiguous()
        gev = torch.empty_like(ev)
        gw = torch.empty_like(w) if (w is not None and ctx.needs_input_grad[1]) else None
        check(_capi.load().ebos_iwe_splat_bwd(ptr(ev), n, b, Hp, Wp, pad_h, pad_w, ptr(w), floor_bias, dtype_code(ev),
                                              ptr(g), ptr(gev), ptr(gw), current_stream()), "ebos_iwe_splat_bwd")
        return gev, gw, None, None, None, None, None, None


def iwe_splat(events: torch.Tensor, padded_size: Tuple[int, int], outer_padding: Tuple[int, int] = (0, 0),
              weight: Union[float, torch.Tensor] = 1.0, deterministic: bool = False,
              floor_bias: float = 1e-6) -> torch.Tensor:
    """Bilinear-vote image [(b,)Hp,Wp]; differentiable w.r.t. the event coordinates and `weight`.

    `deterministic=True` reproduces the reference's sequential tap-major accumulation bit for bit.
    `floor_bias`: 1e-6 (tensor branch, src/event_image_converter.py:586) or 1e-8 (numpy branch, :528)."""
    ev, batched = _as_batched(events)
    _check_cuda(ev)
    w: Optional[torch.Tensor]
    if isinstance(weight, torch.Tensor):
        assert weight.shape == events.shape[:-1]  # same assertion as src/event_image_converter.py:576-577
        w = weight.to(ev.dtype).reshape(ev.shape[0], ev.shape[1])
    elif float(weight) == 1.0:
        w = None
    else:
        w = torch.full((ev.shape[0], ev.shape[1]), float(weight), dtype=ev.dtype, device=ev.device)
    img = _IweSplat.apply(ev, w, int(padded_size[0]), int(padded_size[1]), int(outer_padding[0]), int(outer_padding[1]),
                          bool(deterministic), float(floor_bias))
    return img if batched else img[0]


def iwe_splat_debug(events: torch.Tensor, padded_size: Tuple[int, int], outer_padding: Tuple[int, int] = (0, 0),
                    deterministic: bool = False, floor_bias: float = 1e-6):
    """(image, inds int64 [4n], mask bool [4n]) -- the reference's `inds` / `inds_mask`
    (src/event_image_converter.py:592-617), for bit-exact parity checks.  Unbatched."""
    _check_cuda(events)
    ev = events.contiguous()[None]
    n = ev.shape[1]
    Hp, Wp = int(padded_size[0]), int(padded_size[1])
    image = torch.empty((1, Hp, Wp), dtype=ev.dtype, device=ev.device)
    inds = torch.empty(4 * n, dtype=torch.int64, device=ev.device)
    mask = torch.empty(4 * n, dtype=torch.uint8, device=ev.device)
    lib = _capi.load()
    mode = 1 if deterministic else 0
    ws_bytes = lib.ebos_splat_workspace_bytes(n, 1, Hp, Wp, dtype_code(ev), mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ev.device) if ws_bytes else None
    check(lib.ebos_iwe_splat(ptr(ev), n, 1, Hp, Wp, int(outer_padding[0]), int(outer_padding[1]), 0, float(floor_bias),
                             dtype_code(ev), mode, ptr(image), ptr(inds), ptr(mask), ptr(ws), ws_bytes, current_stream()),
          "ebos_iwe_splat")
    return image[0], inds, mask.bool()


# ------------------------------------------------------------------------------------------------------
# fused path
# ------------------------------------------------------------------------------------------------------
def _float_dtype(dtype) -> torch.dtype:
    if dtype in (torch.float32, "32", 32, "float32"):
        return torch.float32
    if dtype in (torch.float64, "64", 64, "float64"):
        return torch.float64
    raise TypeError(f"the fused path runs in float32 or float64, got {dtype!r}")


class PreparedWindow:
    """Events of one time window, prepared once for many objective evaluations: time-normalised
    (src/warp.py:264-288) and stably sorted by origin pixel (src/warp.py:334) into an SoA buffer.
    `events` is [n,4] on the GPU (x=row, y=col, t, p); `dtype` float32 (fast path, default) or float64
    (the dtype the reference's solvers run in)."""

    def __init__(self, events: torch.Tensor, image_size: Tuple[int, int], direction: Direction = "first",
                 normalize_t: bool = True, weight: Optional[torch.Tensor] = None, validate: bool = True,
                 t_min_max: Optional[torch.Tensor] = None, dtype=torch.float32, allow_packed: bool = True):
        _check_cuda(events, weight, t_min_max)
        if events.dim() != 2 or events.shape[1] != 4:
            raise ValueError(f"events must be [n,4], got {tuple(events.shape)}")
        if events.shape[0] == 0:
            raise RuntimeError("min()/max() of an empty event array (the reference raises here too)")
        self.dtype = _float_dtype(dtype)
        ev = events.to(self.dtype).contiguous()
        self.device = ev.device
        self.n = int(ev.shape[0])
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.direction = direction
        self.normalize_t = bool(normalize_t)
        self.has_weight = weight is not None
        w = None if weight is None else weight.to(self.dtype).contiguous()
        kind, frac = direction_code(direction)
        lib = _capi.load()
        self.code = dtype_code(ev)
        self.buffer = torch.empty(lib.ebos_window_bytes(self.n, self.H, self.W, self.code), dtype=torch.uint8, device=ev.device)
        ws_bytes = lib.ebos_window_workspace_bytes(self.n, self.H, self.W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ev.device)
        status = torch.zeros(1, dtype=torch.int32, device=ev.device)
        # t_min_max: [2] on the device -- the GLOBAL (min t, max t) when this window is one shard of a
        # larger event set (event-sharded multi-GPU path).
        tmm = None if t_min_max is None else t_min_max.to(self.dtype).contiguous()
        check(lib.ebos_window_prepare(ptr(ev), self.n, self.H, self.W, kind, frac, int(self.normalize_t), ptr(w),
                                      ptr(tmm), int(bool(allow_packed)), self.code, ptr(self.buffer), ptr(ws), ws_bytes,
                                      ptr(status), current_stream()), "ebos_window_prepare")
        # One 4-byte read-back per window: which layout prepare chose (packed (row,col,dt) for integer
        # coordinates) and whether an event's pixel lies outside the grid.
        st = int(status.item())
        self.packed = bool(st & _capi.STATUS_PACKED)
        self.flags = (_capi.WIN_HAS_WEIGHT if self.has_weight else 0) | (_capi.WIN_PACKED if self.packed else 0)
        if validate and st & _capi.STATUS_PIXEL_OOB:
            raise RuntimeError("index out of bounds: an event's integer pixel lies outside the flow grid "
                               "(the reference's torch.gather raises for the same input)")

    def permutation(self) -> torch.Tensor:
        """int32 [n]: sorted position -> index of the event in the array the window was built from."""
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        check(_capi.load().ebos_window_info(ptr(self.buffer), self.n, self.H, self.W, self.code, ptr(out), 0, current_stream()),
              "ebos_window_info")
        return out

    def time_info(self) -> torch.Tensor:
        """float64 [4]: t_ref, period, t_min, t_max as computed on the device (exact values of `dtype`)."""
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        check(_capi.load().ebos_window_info(ptr(self.buffer), self.n, self.H, self.W, self.code, 0, ptr(out), current_stream()),
              "ebos_window_info")
        return out


def _check_flow(window: PreparedWindow, flow: torch.Tensor) -> None:
    if flow.dtype != window.dtype or tuple(flow.shape) != (2, window.H, window.W) or not flow.is_contiguous():
        raise ValueError(f"flow must be a contiguous {window.dtype} [2,{window.H},{window.W}] tensor, got "
                         f"{flow.dtype} {tuple(flow.shape)}")


def window_splat(window: PreparedWindow, flow: torch.Tensor, outer_padding: Tuple[int, int] = (0, 0),
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """IWE [Hp,Wp] of the window warped by `flow` [2,H,W] (fused warp + bilinear vote, atomic mode)."""
    _check_cuda(flow)
    _check_flow(window, flow)
    ph, pw = int(outer_padding[0]), int(outer_padding[1])
    if out is None:
        out = torch.empty((window.H + 2 * ph, window.W + 2 * pw), dtype=window.dtype, device=flow.device)
    check(_capi.load().ebos_window_splat(ptr(window.buffer), window.n, window.flags, ptr(flow), window.H,
                                         window.W, ph, pw, window.code, ptr(out), current_stream()), "ebos_window_splat")
    return out


class CmaxWorkspace:
    """Scratch planes for `cmax_value_and_grad`, allocated once per (H, W, padding, dtype)."""

    def __init__(self, H: int, W: int, outer_padding: Tuple[int, int] = (0, 0), device="cuda", dtype=torch.float32):
        ph, pw = int(outer_padding[0]), int(outer_padding[1])
        self.H, self.W, self.ph, self.pw = H, W, ph, pw
        self.dtype = _float_dtype(dtype)
        # accumulators and IWE back to back in one allocation: the fused entry then zeroes both with one memset node
        Hp, Wp = H + 2 * ph, W + 2 * pw
        acc_bytes = _capi.ACC_DOUBLES * 8
        item = torch.empty((), dtype=self.dtype).element_size()
        self._acc_iwe = torch.zeros(acc_bytes + Hp * Wp * item, dtype=torch.uint8, device=device)
        self.acc = self._acc_iwe[:acc_bytes].view(torch.float64)
        self.iwe = self._acc_iwe[acc_bytes:].view(self.dtype).view(Hp, Wp)
        self.grad_iwe = torch.empty_like(self.iwe)
        self.dflow = torch.empty((2, H, W), dtype=self.dtype, device=device)
        self.loss = torch.zeros(1, dtype=self.dtype, device=device)
        # "clean workspace" protocol of the fused entries: accumulators and IWE are all zero between calls (they are
        # zero-filled for the NEXT evaluation concurrently with the backward of the current one).  Cleared for good by a
        # call that keeps its IWE (`keep_iwe=True`) or by anyone writing into `acc` / `iwe` directly.
        self.clean = True
        self.dflow_clean = False     # `dflow` all zero between calls (contract of cmax_adam_iteration_fused_tv)
        self._blur = None

    def zero_dflow(self) -> None:
        self.dflow.zero_()
        self.dflow_clean = True

    def blur_plane(self) -> torch.Tensor:
        """Scratch plane of the blurred-IWE objective (`blur_sigma > 0`), allocated on first use."""
        if self._blur is None:
            self._blur = torch.empty_like(self.iwe)
        return self._blur


def cmax_value_and_grad(window: PreparedWindow, flow: torch.Tensor, cost: str = "gradient_magnitude",
                        data_weight: float = 1.0, tv_weight: float = 0.0, tv_weights: Optional[torch.Tensor] = None,
                        omit_boundary: bool = False, outer_padding: Tuple[int, int] = (0, 0),
                        workspace: Optional[CmaxWorkspace] = None, keep_iwe: bool = False,
                        blur_sigma: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss [1], dL/dflow [2,H,W]) of   data_weight * L_cost(IWE(warp(events, flow))) + tv_weight * TV(flow).

    One C call, five kernels, no host synchronisation and no materialised warped events or [4N]
    temporaries.  The returned tensors alias `workspace` (overwritten by the next call).  `keep_iwe=True` leaves the
    IWE of this evaluation in `workspace.iwe` (the workspace then zero-fills at the start of every later call).
    `blur_sigma > 0`: the data objective is taken on torchvision's `gaussian_blur(IWE, kernel_size=3, sigma)`
    (src/event_image_converter.py:399-404) and differentiated through it."""
    _check_cuda(flow, tv_weights)
    _check_flow(window, flow)
    if cost not in COST_KINDS:
        raise KeyError(f"unknown data cost {cost!r}; available: {sorted(COST_KINDS)}")
    ws = workspace or CmaxWorkspace(window.H, window.W, outer_padding, flow.device, window.dtype)
    if ws.dtype != window.dtype:
        raise TypeError(f"workspace dtype {ws.dtype} does not match the window's {window.dtype}")
    tvw = None
    if tv_weights is not None:
        tvw = tv_weights.to(window.dtype).contiguous()
        if tuple(tvw.shape) != (window.H, window.W):
            raise ValueError(f"tv_weights must be [{window.H},{window.W}], got {tuple(tvw.shape)}")
    if keep_iwe:
        ws.clean = False
    ws.dflow_clean = False
    check(_capi.load().ebos_cmax_value_and_grad(
        ptr(window.buffer), window.n, window.flags, ptr(flow), window.H, window.W, ws.ph, ws.pw,
        COST_KINDS[cost], int(bool(omit_boundary)), float(data_weight), float(tv_weight), ptr(tvw), window.code,
        ptr(ws.iwe), ptr(ws.grad_iwe), ptr(ws.dflow), ptr(ws.loss), ptr(ws.acc), int(ws.clean), float(blur_sigma),
        ptr(ws.blur_plane()) if blur_sigma > 0 else 0, current_stream()), "ebos_cmax_value_and_grad")
    return ws.loss, ws.dflow


class CmaxGraph:
    """One fused evaluation (`cmax_value_and_grad`) captured as a CUDA graph.

    `replay()` re-runs the captured kernels on the CURRENT contents of `flow` (the same tensor object: update it in
    place) and returns `(loss, dflow)` aliasing the workspace.  The C entry is capture-safe (no allocation, no host
    synchronisation; its TV | splat fork/join becomes graph dependencies), and a replay costs one launch instead of
    seven plus four cross-stream event operations."""

    def __init__(self, window: PreparedWindow, flow: torch.Tensor, cost: str = "gradient_magnitude",
                 data_weight: float = 1.0, tv_weight: float = 0.0, tv_weights: Optional[torch.Tensor] = None,
                 omit_boundary: bool = False, outer_padding: Tuple[int, int] = (0, 0),
                 workspace: Optional[CmaxWorkspace] = None, blur_sigma: float = 0.0):
        self.window, self.flow = window, flow
        self.ws = workspace or CmaxWorkspace(window.H, window.W, outer_padding, flow.device, window.dtype)
        self._tvw = None if tv_weights is None else tv_weights.to(window.dtype).contiguous()
        if blur_sigma > 0:
            self.ws.blur_plane()                          # allocate outside the capture
        args = (window, flow, cost, data_weight, tv_weight, self._tvw, omit_boundary, outer_padding, self.ws, False, blur_sigma)
        cur = torch.cuda.current_stream(flow.device)
        side = torch.cuda.Stream(device=flow.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            cmax_value_and_grad(*args)   # warm-up outside capture (module loading), also validates the arguments
            self.graph = torch.cuda.CUDAGraph()
            self.graph.capture_begin()
            cmax_value_and_grad(*args)
            self.graph.capture_end()
        cur.wait_stream(side)

    def replay(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self.graph.replay()
        return self.ws.loss, self.ws.dflow


def cmax_adam_iteration(window: PreparedWindow, flow: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                        step_dev: torch.Tensor, workspace: CmaxWorkspace, cost: str = "gradient_magnitude",
                        data_weight: float = 1.0, tv_weight: float = 0.0, tv_weights: Optional[torch.Tensor] = None,
                        omit_boundary: bool = False, lr: float = 0.05, betas: Tuple[float, float] = (0.9, 0.999),
                        eps: float = 1e-8, blur_sigma: float = 0.0) -> torch.Tensor:
    """One solver iteration in one C call: objective + gradient (like `cmax_value_and_grad`) and the Adam update of
    `flow` in place.  `workspace.acc` and `workspace.iwe` must be zero on entry (a fresh or `clean` CmaxWorkspace; they
    are left zero), `step_dev` (int32 [1]) counts the completed iterations.  Returns `workspace.loss` (the objective
    before the update)."""
    _check_cuda(flow, tv_weights, exp_avg, exp_avg_sq, step_dev)
    _check_flow(window, flow)
    ws = workspace
    if ws.dtype != window.dtype or exp_avg.dtype != window.dtype or exp_avg_sq.dtype != window.dtype:
        raise TypeError("cmax_adam_iteration: window, workspace, flow and Adam moments must share a dtype")
    tvw = None
    if tv_weights is not None:
        tvw = tv_weights.to(window.dtype).contiguous()
    if not ws.clean:
        raise RuntimeError("cmax_adam_iteration needs a clean CmaxWorkspace (accumulators and IWE zero on entry)")
    ws.dflow_clean = False
    check(_capi.load().ebos_cmax_adam_iteration(
        ptr(window.buffer), window.n, window.flags, ptr(flow), window.H, window.W, ws.ph, ws.pw, COST_KINDS[cost],
        int(bool(omit_boundary)), float(data_weight), float(tv_weight), ptr(tvw), window.code, ptr(ws.iwe),
        ptr(ws.grad_iwe), ptr(ws.dflow), ptr(ws.loss), ptr(ws.acc), ptr(exp_avg), ptr(exp_avg_sq), float(lr),
        float(betas[0]), float(betas[1]), float(eps), ptr(step_dev), float(blur_sigma),
        ptr(ws.blur_plane()) if blur_sigma > 0 else 0, current_stream()), "ebos_cmax_adam_iteration")
    return ws.loss


def fused_tv_supported(window: PreparedWindow, tv_weights: Optional[torch.Tensor] = None) -> bool:
    """Whether `cmax_adam_iteration_fused_tv` applies: fp32, unit TV weights, W % 4 == 0, W >= 12, H >= 5."""
    return (window.dtype == torch.float32 and tv_weights is None and window.W % 4 == 0 and window.W >= 12
            and window.H >= 5)


def cmax_adam_iteration_fused_tv(window: PreparedWindow, flow_in: torch.Tensor, flow_out: torch.Tensor,
                                 exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step_dev: torch.Tensor,
                                 workspace: CmaxWorkspace, cost: str = "gradient_magnitude", data_weight: float = 1.0,
                                 tv_weight: float = 0.0, omit_boundary: bool = False, lr: float = 0.05,
                                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                                 blur_sigma: float = 0.0) -> torch.Tensor:
    """`cmax_adam_iteration` with the TV term folded into the Adam kernel (one launch and 2 of 18 plane passes less per
    iteration): reads `flow_in`, writes the updated flow to `flow_out` (another [2,H,W] tensor; callers alternate the
    two).  On top of the clean-workspace contract, `workspace.dflow` must be zero on entry (`workspace.zero_dflow()`
    once) and is left zero.  Returns `workspace.loss` (the objective before the update)."""
    _check_cuda(flow_in, flow_out, exp_avg, exp_avg_sq, step_dev)
    _check_flow(window, flow_in)
    _check_flow(window, flow_out)
    ws = workspace
    if not fused_tv_supported(window):
        raise ValueError("cmax_adam_iteration_fused_tv needs an fp32 window with W % 4 == 0, W >= 12, H >= 5")
    if flow_in.data_ptr() == flow_out.data_ptr():
        raise ValueError("cmax_adam_iteration_fused_tv: flow_in and flow_out must be different tensors")
    if ws.dtype != window.dtype or exp_avg.dtype != window.dtype or exp_avg_sq.dtype != window.dtype:
        raise TypeError("cmax_adam_iteration_fused_tv: window, workspace, flows and Adam moments must share a dtype")
    if not ws.clean or not ws.dflow_clean:
        raise RuntimeError("cmax_adam_iteration_fused_tv needs a clean CmaxWorkspace with a zeroed gradient plane "
                           "(workspace.zero_dflow())")
    check(_capi.load().ebos_cmax_adam_iteration_fused_tv(
        ptr(window.buffer), window.n, window.flags, ptr(flow_in), ptr(flow_out), window.H, window.W, ws.ph, ws.pw,
        COST_KINDS[cost], int(bool(omit_boundary)), float(data_weight), float(tv_weight), window.code, ptr(ws.iwe),
        ptr(ws.grad_iwe), ptr(ws.dflow), ptr(ws.loss), ptr(ws.acc), ptr(exp_avg), ptr(exp_avg_sq), float(lr),
        float(betas[0]), float(betas[1]), float(eps), ptr(step_dev), float(blur_sigma),
        ptr(ws.blur_plane()) if blur_sigma > 0 else 0, current_stream()), "ebos_cmax_adam_iteration_fused_tv")
    return ws.loss


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float = 0.05, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
              step_dev: Optional[torch.Tensor] = None) -> None:
    """torch.optim.Adam update in place.  With `step_dev` (int32 [1] on the device) the step counter
    lives on the GPU so that the call can be captured in a CUDA graph."""
    _check_cuda(param, grad, exp_avg, exp_avg_sq)
    code = dtype_code(param)
    if not (grad.dtype == exp_avg.dtype == exp_avg_sq.dtype == param.dtype):
        raise TypeError("adam_step: param, grad and moments must share a dtype")
    lib = _capi.load()
    if step_dev is not None:
        check(lib.ebos_adam_step_graph(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, betas[0],
                                       betas[1], eps, ptr(step_dev), code, current_stream()), "ebos_adam_step_graph")
    else:
        check(lib.ebos_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, betas[0],
                                 betas[1], eps, int(step), code, current_stream()), "ebos_adam_step")


# operator-level cost kernels (used by the CostBase-style classes) ----------------------------------------
class _IweCost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, iwe, kind, omit_boundary):
        img = iwe.contiguous()
        Hp, Wp = img.shape
        code = dtype_code(img)
        acc = torch.zeros(_capi.ACC_DOUBLES, dtype=torch.float64, device=img.device)
        grad = torch.empty_like(img)
        loss = torch.empty(1, dtype=img.dtype, device=img.device)
        lib = _capi.load()
        st = current_stream()
        check(lib.ebos_iwe_cost(kind, ptr(img), Hp, Wp, int(omit_boundary), 1.0, code, ptr(acc), ptr(grad), st),
              "ebos_iwe_cost")
        check(lib.ebos_loss_finalize(kind, ptr(acc), Hp, Wp, 1, 1, int(omit_boundary), 1.0, 0.0, code, ptr(loss), st),
              "ebos_loss_finalize")
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


def iwe_cost(iwe: torch.Tensor, cost: str, omit_boundary: bool = False) -> torch.Tensor:
    """Scalar data objective on an IWE plane (fp32/fp64), differentiable (analytic gradient kernel)."""
    _check_cuda(iwe)
    if iwe.dim() != 2:
        raise ValueError(f"iwe must be [H,W], got {tuple(iwe.shape)}")
    dtype_code(iwe)
    return _IweCost.apply(iwe, COST_KINDS[cost], bool(omit_boundary))


class _FlowTv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, weights):
        fl = flow.contiguous()
        _, H, W = fl.shape
        acc = torch.zeros(_capi.ACC_DOUBLES, dtype=torch.float64, device=fl.device)
        dflow = torch.empty_like(fl)
        check(_capi.load().ebos_flow_tv(ptr(fl), ptr(weights), H, W, 1.0, dtype_code(fl), ptr(acc), ptr(dflow),
                                        current_stream()), "ebos_flow_tv")
        ctx.save_for_backward(dflow)
        return ((acc[3] + acc[24:40].sum()) / (2.0 * H * W)).to(fl.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (dflow,) = ctx.saved_tensors
        return dflow * grad_out, None


def flow_total_variation(flow: torch.Tensor, weights: Union[float, torch.Tensor, None] = None) -> torch.Tensor:
    """mean(|d flow/d row * w| + |d flow/d col * w|)  (src/costs/image_gradient.py:60-70), differentiable
    w.r.t. `flow` (fp32/fp64)."""
    _check_cuda(flow)
    if flow.dim() != 3 or flow.shape[0] != 2:
        raise ValueError(f"flow must be [2,H,W], got {tuple(flow.shape)}")
    if flow.shape[1] < 2 or flow.shape[2] < 2:
        raise RuntimeError("torch.gradient expected each dimension size to be at least edge_order+1")
    dtype_code(flow)
    w = None
    if isinstance(weights, torch.Tensor):
        w = weights.to(device=flow.device, dtype=flow.dtype)
        # upstream multiplies `torch.gradient(flow)[0]` [2,H,W] by `weights`, i.e. anything that broadcasts against it
        # (src/costs/image_gradient.py:69-70): scalars, [H,W], [1,H,W] (a `mask[None]` ROI weight) and per-channel [2,H,W]
        if w.dim() == 3 and w.shape[0] == 2:
            zero = torch.zeros_like(flow[0])
            per_channel = w.expand(2, flow.shape[1], flow.shape[2])
            return (_FlowTv.apply(torch.stack([flow[0], zero]), per_channel[0].contiguous())
                    + _FlowTv.apply(torch.stack([zero, flow[1]]), per_channel[1].contiguous()))
        if w.dim() == 3:
            w = w[0]
        w = w.expand(flow.shape[1], flow.shape[2]).contiguous()
    elif weights is not None and float(weights) != 1.0:
        w = torch.full(flow.shape[1:], float(weights), dtype=flow.dtype, device=flow.device)
    return _FlowTv.apply(flow, w)


class _Blur3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, sigma):
        img = image.contiguous()
        H, W = img.shape[-2:]
        batch = img.numel() // (H * W)
        out = torch.empty_like(img)
        check(_capi.load().ebos_blur3(ptr(img), batch, H, W, sigma, 0, dtype_code(img), ptr(out), current_stream()), "ebos_blur3")
        ctx.sigma = sigma
        return out

    @staticmethod
    def backward(ctx, grad_out):
        g = grad_out.contiguous()
        H, W = g.shape[-2:]
        out = torch.empty_like(g)
        check(_capi.load().ebos_blur3(ptr(g), g.numel() // (H * W), H, W, ctx.sigma, 1, dtype_code(g), ptr(out), current_stream()),
              "ebos_blur3(adjoint)")
        return out, None


def blur3(image: torch.Tensor, sigma: float) -> torch.Tensor:
    """torchvision `gaussian_blur(kernel_size=3, sigma)` over the last two dimensions of `image` [...,H,W] (reflect
    padding; src/event_image_converter.py:399-404), differentiable: forward and exact adjoint are CUDA stencil kernels."""
    _check_cuda(image)
    if image.dim() < 2 or image.shape[-1] < 2 or image.shape[-2] < 2:
        raise ValueError(f"blur3 needs [...,H,W] with H, W >= 2 (reflect padding), got {tuple(image.shape)}")
    if not sigma > 0:
        raise ValueError("sigma must be positive")
    dtype_code(image)
    return _Blur3.apply(image, float(sigma))


class ReplaySlot:
    """A replayable launch sequence that outlives the window it was captured for (include/ebos.h:
    ebos_capture_begin / ebos_capture_end_exec / ebos_exec_launch).

    `capture(fn)` records what `fn()` enqueues on the CURRENT stream (which must not be the default stream; `fn` must be
    capture-safe: no allocation, no host synchronisation) and makes it this slot's executable: the first call
    instantiates it, later calls update it in place to the new pointers / sizes -- no instantiation, no destruction and
    hence no device-wide synchronisation per window (a torch.cuda.CUDAGraph per window costs two: its private memory
    pool and its executable are freed with device-synchronising calls, which serialised estimate_many's windows).
    `launch()` enqueues one replay on the current stream."""

    def __init__(self):
        import ctypes

        self._exec = ctypes.c_void_p(0)
        self.updates = 0            # captures that updated the executable in place
        self.instantiations = 0     # captures that had to (re)build it

    def capture(self, fn) -> None:
        import ctypes

        lib = _capi.load()
        stream = current_stream()
        check(lib.ebos_capture_begin(stream), "ebos_capture_begin")
        updated = ctypes.c_int32(0)
        try:
            fn()
        except BaseException:
            lib.ebos_capture_end_count(stream, None, None)      # leave capture mode, discard the partial graph
            raise
        check(lib.ebos_capture_end_exec(stream, ctypes.byref(self._exec), ctypes.byref(updated)), "ebos_capture_end_exec")
        if updated.value:
            self.updates += 1
        else:
            self.instantiations += 1

    def launch(self) -> None:
        check(_capi.load().ebos_exec_launch(self._exec, current_stream()), "ebos_exec_launch")

    def close(self) -> None:
        if self._exec:
            _capi.load().ebos_exec_destroy(self._exec)
            self._exec.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def count_launches(fn) -> Tuple[int, int]:
    """(kernel launches, other graph nodes such as memsets) that `fn()` enqueues on the current stream, counted from a
    throw-away stream capture (nothing runs).  `fn` must be capture-safe: no allocation, no host synchronisation."""
    import ctypes

    _capi.require_device()
    lib = _capi.load()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    k, o = ctypes.c_int32(0), ctypes.c_int32(0)
    with torch.cuda.stream(side):
        check(lib.ebos_capture_begin(side.cuda_stream), "ebos_capture_begin")
        try:
            fn()
        finally:
            check(lib.ebos_capture_end_count(side.cuda_stream, ctypes.byref(k), ctypes.byref(o)), "ebos_capture_end_count")
    torch.cuda.current_stream().wait_stream(side)
    return int(k.value), int(o.value)
