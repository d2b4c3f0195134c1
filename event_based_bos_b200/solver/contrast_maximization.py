"""Per-window dense-flow contrast maximisation on the fused CUDA path.

The reference's configs name this solver (`configs/README.md:59,73` -> `contrast_maximization`) but the
snapshot does not contain it.  It is built here out of the reference's own parts:
the operators `Warp.warp_event("dense-flow")` + `EventImageConverter.create_iwe("bilinear_vote")`, a
`CostBase`-style objective, and the Adam loop idiom of src/solver/patch_eklt_pyramid2.py:259-288
(Adam lr 0.05, StepLR(step=iters, gamma=0.1) -- which never decays inside the loop -- zero_grad / loss /
backward / step; because `best_x` aliases the optimised leaf upstream, the FINAL iterate is returned).
"""
import logging
import time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import costs, ops, utils
from .base import SolverBase

logger = logging.getLogger(__name__)

DATA_COSTS = ("image_variance", "gradient_magnitude")


def split_cost_weights(cost_with_weight: Dict[str, float]) -> Tuple[str, float, float]:
    """{'gradient_magnitude': 1.0, 'image_gradient': 0.5} -> ('gradient_magnitude', 1.0, 0.5)."""
    data = [(k, v) for k, v in cost_with_weight.items() if k in DATA_COSTS]
    extra = [k for k in cost_with_weight if k not in DATA_COSTS and k != "image_gradient"]
    if len(data) != 1 or extra:
        raise ValueError(f"contrast maximisation needs exactly one data cost of {DATA_COSTS} plus an optional "
                         f"'image_gradient' regulariser; got {dict(cost_with_weight)}")
    return data[0][0], float(data[0][1]), float(cost_with_weight.get("image_gradient", 0.0))


class ContrastMaximizationDense(SolverBase):
    """Pixel-wise flow by contrast maximisation with a TV regulariser.

    Reads from `solver_config` (yaml `solver:` block): `warp_direction`, `outer_padding`, `optimizer.n_iter`,
    `optimizer.method` ("Adam"), and the sub-dict `cmax` with `cost_with_weight` (default
    {gradient_magnitude: 1.0, image_gradient: 0.5}), `lr` (0.05), `omit_boundary` (False), `fused` (True),
    `cuda_graph` (True), `fold_tv` (False), `store_history` (False), `precision` ("32" fast path | "64" = the dtype the
    reference's solvers run in, src/solver/patch_eklt_pyramid2.py:253).
    """

    def __init__(self, orig_image_shape: tuple, crop_image_shape: tuple, calibration_parameter: dict = {},
                 solver_config: dict = {}, visualize_module=None):
        super().__init__(orig_image_shape, crop_image_shape, calibration_parameter, solver_config, visualize_module)
        cm = dict(self.slv_config.get("cmax", {}))
        self.cost_with_weight = cm.get("cost_with_weight", {"gradient_magnitude": 1.0, "image_gradient": 0.5})
        self.data_cost, self.data_weight, self.tv_weight = split_cost_weights(self.cost_with_weight)
        self.lr = float(cm.get("lr", 0.05))
        self.omit_boundary = bool(cm.get("omit_boundary", False))
        self.fused = bool(cm.get("fused", True))
        self.use_cuda_graph = bool(cm.get("cuda_graph", True))
        self.store_history = bool(cm.get("store_history", False))
        # TV term inside the Adam kernel (ebos_cmax_adam_iteration_fused_tv).  Off by default: measured on B200 (r02r) the
        # fused kernel is L2-bandwidth bound on its 3x re-read of the flow rows (17 us against 10.7 us for Adam alone,
        # with the separate TV kernel hidden beside the splat): 41.0 vs 49.0 windows/s with one window in flight,
        # 56.6 vs 66.2 with eight
        self.fold_tv = bool(cm.get("fold_tv", False))
        self.precision = str(cm.get("precision", "32"))
        if self.precision not in ("32", "64"):
            raise ValueError(f"cmax.precision must be '32' or '64', got {self.precision!r}")
        self._dtype = torch.float64 if self.precision == "64" else torch.float32
        self.warp_direction = self.slv_config.get("warp_direction", "first")
        # `iwe: {method: bilinear_vote, blur_sigma: s}` of the solver configs (configs/hot_plate1.yaml): the data cost is
        # taken on the 3x3-Gaussian-blurred IWE (src/event_image_converter.py:399-404)
        iwe_cfg = self.slv_config.get("iwe", {}) or {}
        if iwe_cfg.get("method", "bilinear_vote") != "bilinear_vote":
            raise NotImplementedError(f"iwe.method {iwe_cfg['method']!r}: only 'bilinear_vote' has a tensor branch upstream")
        self.blur_sigma = float(iwe_cfg.get("blur_sigma", 0) or 0)
        opt = self.slv_config.get("optimizer", {})
        self._opt_method = opt.get("method", "Adam")
        self.n_iter = int(opt.get("n_iter", 600))
        if self._opt_method != "Adam":
            raise ValueError(f"ContrastMaximizationDense supports optimizer.method = 'Adam', got {self._opt_method!r}")
        self.history: Dict[str, List[float]] = {"loss": []}
        self._staging: Dict[int, torch.Tensor] = {}
        self._hist_dev = None
        self.last_many_stats: Dict[str, float] = {}
        self._side = None                       # capture stream of `estimate` on the default stream
        self._replay = None                     # ops.ReplaySlot of `estimate`
        self._replays: List = []                # ... of the estimate_many slots
        self._streams: List = []
        self._join = lambda: None
        self.cost_func = costs.HybridCost("minimize", self.cost_with_weight, store_history=self.store_history)

    # ------------------------------------------------------------------------------------------
    def estimate(self, events: np.ndarray, *args, flow0: Optional[np.ndarray] = None, **kwargs) -> np.ndarray:
        """[n,4] events (x=row, y=col, t [s], p) -> flow [2,H,W] float64 (pixel displacement over the window)."""
        return self._download(self._enqueue(events, flow0))

    def _download(self, out: torch.Tensor, slot: int = 0) -> np.ndarray:
        """Device result -> numpy through a persistent pinned staging tensor (one per concurrency slot): a pageable
        `.cpu()` of the 14.7 MB float64 flow took 7 ms, a quarter of a 500 k-event solve; allocating pinned memory per
        call is worse (cudaHostAlloc synchronises the device)."""
        stage = self._stage(out, slot)
        stage.copy_(out, non_blocking=True)
        torch.cuda.current_stream(out.device).synchronize()
        return stage.numpy().copy()

    def _stage(self, out: torch.Tensor, slot: int = 0) -> torch.Tensor:
        stage = self._staging.get(slot)
        if stage is None or stage.shape != out.shape or stage.dtype != out.dtype:
            stage = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            self._staging[slot] = stage
        return stage

    def estimate_many(self, windows: Sequence[np.ndarray], concurrency: int = 4,
                      flow0: Optional[Sequence[Optional[np.ndarray]]] = None) -> List[np.ndarray]:
        """Solve independent time windows (no inter-window state upstream) with up to `concurrency`
        solves in flight on separate CUDA streams.  One small window (~0.5 M events) leaves most of the 148 SMs idle
        -- its splat is ~120 CTAs -- so overlapping windows raises windows/s without touching the per-window result:
        every window runs exactly the kernels and the order of `estimate` (results are identical).

        Rolling schedule: window i goes to slot i mod `concurrency`.  A slot's whole solve (H2D, prepare, the graph
        replays, ROI mask, D2H into the slot's pinned staging buffer) is queued in one go; the host then moves on to
        the next slot and comes back to this one `concurrency` windows later, when it retires the result (event wait +
        one host copy) and queues the next window.  The other slots keep the GPU busy while the host plans a window
        (~4 ms), and the solves run staggered (one window's Adam next to another's splat) instead of in lock step.
        `self.last_many_stats` holds the host-side time per window of the last call."""
        n_slots = max(1, min(int(concurrency), len(windows)))
        if not self.fused or self.store_history or any(len(w) >= (1 << 21) for w in windows):
            # large windows fill the GPU on their own (and take the eager, non-graph path); a loss history is kept per
            # solver object, i.e. for one window at a time
            n_slots = 1
        if n_slots == 1:
            return [self.estimate(w, flow0=None if flow0 is None else flow0[i]) for i, w in enumerate(windows)]
        while len(self._streams) < n_slots:      # slot streams and executables live as long as the solver
            self._streams.append(torch.cuda.Stream(device=self._device))
            self._replays.append(ops.ReplaySlot())
        streams = self._streams
        results: List[Optional[np.ndarray]] = [None] * len(windows)
        busy: List[Optional[tuple]] = [None] * n_slots
        t_plan = t_wait = t_copy = t_free = 0.0
        t_begin = time.perf_counter()

        def retire(slot: int) -> None:
            nonlocal t_wait, t_copy, t_free
            idx, stage, done, keep = busy[slot]
            busy[slot] = None
            t0 = time.perf_counter()
            done.synchronize()
            t1 = time.perf_counter()
            results[idx] = stage.numpy().copy()
            t2 = time.perf_counter()
            del keep            # the window's CUDA graph and buffers go here
            t_wait += t1 - t0
            t_copy += t2 - t1
            t_free += time.perf_counter() - t2

        for idx, events in enumerate(windows):
            slot = idx % n_slots
            if busy[slot] is not None:
                retire(slot)
            t0 = time.perf_counter()
            with torch.cuda.stream(streams[slot]):
                x0 = self._upload_flow0(None if flow0 is None else flow0[idx])
                advance, n_calls = self._plan_fused(self._upload_events(events), x0, self._replays[slot])
                for _ in range(n_calls):
                    advance()
                out = self._finish(x0)
                stage = self._stage(out, slot)
                stage.copy_(out, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
            busy[slot] = (idx, stage, done, (advance, out))   # (the graph and its buffers live until the slot retires)
            t_plan += time.perf_counter() - t0
        for slot in sorted((s for s in range(n_slots) if busy[s] is not None), key=lambda s: busy[s][0]):
            retire(slot)
        k = 1e3 / len(windows)
        self.last_many_stats = {"windows": len(windows), "slots": n_slots, "host_plan_and_queue_ms": t_plan * k,
                                "host_wait_for_gpu_ms": t_wait * k, "host_result_copy_ms": t_copy * k,
                                "host_release_ms": t_free * k, "host_total_ms": (time.perf_counter() - t_begin) * k}
        return results

    def _upload_events(self, events: np.ndarray) -> torch.Tensor:
        from .. import _capi

        _capi.require_device()
        # Absolute sensor time -> window-relative IN FLOAT64, before the cast to the solver dtype (the fp32 ulp at
        # t = 10 s is 1 us; see utils.rebase_time).  Done on the device: one H2D copy of the raw float64 events.
        raw = torch.from_numpy(np.ascontiguousarray(events, dtype=np.float64)).to(self._device, non_blocking=True)
        if raw.shape[0]:
            raw[:, 2] -= raw[:, 2].min()
        return raw.to(self._dtype)

    def _upload_flow0(self, flow0: Optional[np.ndarray]) -> torch.Tensor:
        H, W = self.orig_image_shape
        x0 = torch.zeros((2, H, W), dtype=self._dtype, device=self._device)
        if flow0 is not None:
            x0.copy_(torch.from_numpy(np.asarray(flow0)).to(self._dtype))
        return x0

    def _finish(self, flow: torch.Tensor) -> torch.Tensor:
        """ROI mask and float64 conversion on the device (one D2H copy of the result follows)."""
        H, W = self.orig_image_shape
        out = flow.detach().to(torch.float64)
        if (self.crop_xmin, self.crop_ymin, self.crop_xmax, self.crop_ymax) != (0, 0, H, W):
            mask = torch.zeros((H, W), dtype=torch.float64, device=out.device)
            mask[self.crop_xmin:self.crop_xmax, self.crop_ymin:self.crop_ymax] = 1.0
            out = out * mask
        return out

    def _enqueue(self, events: np.ndarray, flow0: Optional[np.ndarray] = None) -> torch.Tensor:
        """Queue one solve on the current CUDA stream; returns the device tensor the result will be in."""
        ev = self._upload_events(events)
        x0 = self._upload_flow0(flow0)
        self.history = {"loss": []}
        self._hist_dev = None
        flow = self._solve_fused(ev, x0) if self.fused else self._solve_operators(ev, x0)
        return self._finish(flow)

    def _roi_mask(self) -> np.ndarray:
        mask = np.zeros(self.orig_image_shape)
        mask[self.crop_xmin:self.crop_xmax, self.crop_ymin:self.crop_ymax] = 1.0
        return mask[None]

    # -- fused CUDA path ---------------------------------------------------------------------------
    def _plan_fused(self, ev: torch.Tensor, x0: torch.Tensor, replay=None) -> Tuple[Callable[[], None], int]:
        """Prepare one window for the fused path on the current stream.  Returns (advance, n_calls): calling
        `advance()` n_calls times on that stream performs exactly n_iter solver iterations on `x0` in place."""
        H, W = self.orig_image_shape
        self._join = lambda: None
        window = ops.PreparedWindow(ev, (H, W), self.warp_direction, self.normalize_t_in_batch, dtype=self._dtype)
        pad = (self.padding, self.padding)
        ws = ops.CmaxWorkspace(H, W, pad, x0.device, self._dtype)
        if self.blur_sigma > 0:
            ws.blur_plane()                               # allocate outside the graph capture
        m, v = torch.zeros_like(x0), torch.zeros_like(x0)
        step_dev = torch.zeros(1, dtype=torch.int32, device=x0.device)
        hist = torch.zeros(max(self.n_iter, 1), dtype=self._dtype, device=x0.device) if self.store_history else None
        self._hist_dev = hist
        count = [0, 0]     # (loss-history slot, iterations issued: parity selects the source flow plane)

        # opt-in: TV term folded into the Adam kernel (fp32, W % 4 == 0; an even iteration count, so that the two flow
        # planes the kernel alternates between end on `x0`): one launch and 2 of 18 plane passes less per iteration
        fold = self.fold_tv and ops.fused_tv_supported(window) and self.n_iter % 2 == 0
        flows = [x0, torch.empty_like(x0)] if fold else [x0]
        if fold:
            ws.zero_dflow()

        def iteration():
            if fold:
                # one C call: splat -> cost -> backward -> Adam + TV (+ loss, accumulator reset), five graph nodes
                src = count[1] & 1
                ops.cmax_adam_iteration_fused_tv(window, flows[src], flows[1 - src], m, v, step_dev, ws, self.data_cost,
                                                 self.data_weight, self.tv_weight, self.omit_boundary, self.lr,
                                                 blur_sigma=self.blur_sigma)
                count[1] += 1
            else:
                # one C call: TV | splat -> cost -> backward -> Adam (+ loss, accumulator reset), six graph nodes
                ops.cmax_adam_iteration(window, x0, m, v, step_dev, ws, self.data_cost, self.data_weight, self.tv_weight,
                                        None, self.omit_boundary, self.lr, blur_sigma=self.blur_sigma)
            if hist is not None:
                hist[count[0]].copy_(ws.loss[0])
                count[0] += 1

        # Replayed launch sequence: worth its capture cost when iterations are short (launch-bound); for large windows
        # (>= ~2 Mi events, >100 us of GPU work per iteration) eager launches run ahead of the GPU anyway.
        if not (self.use_cuda_graph and not self.store_history and self.n_iter > 2 and window.n < (1 << 21)):
            return iteration, self.n_iter
        # `unroll` captured iterations per executable, replayed n_iter / unroll times; the Adam step counter lives on
        # the device.  Fewer, longer launches keep the host ahead when several solves are in flight.  The executable
        # belongs to `replay` (one per estimate_many slot / per solver) and is UPDATED for each new window, never
        # destroyed while windows are in flight (ops.ReplaySlot).
        unroll = next(u for u in ((10, 8, 6, 4, 2) if fold else (10, 8, 6, 5, 4, 3, 2, 1)) if self.n_iter % u == 0)
        backup = x0.clone()
        cur = torch.cuda.current_stream()
        side = None
        if cur == torch.cuda.default_stream(x0.device):
            # the default stream cannot be captured: one persistent side stream per solver
            if self._side is None:
                self._side = torch.cuda.Stream(device=x0.device)
            side = self._side
            side.wait_stream(cur)
        if replay is None:
            if self._replay is None:
                self._replay = ops.ReplaySlot()
            replay = self._replay

        def capture():
            iteration()        # warm-up outside capture (module loading, the auxiliary lane of this stream)
            x0.copy_(backup)   # the solve must perform exactly n_iter updates
            m.zero_()
            v.zero_()
            step_dev.zero_()
            ws.acc.zero_()
            count[1] = 0
            replay.capture(lambda: [iteration() for _ in range(unroll)])

        if side is None:
            capture()
        else:
            with torch.cuda.stream(side):
                capture()
        keep = (window, ws, m, v, step_dev, backup, flows)   # buffers the executable points into

        def advance(_keep=keep):
            if side is None:
                replay.launch()
            else:
                with torch.cuda.stream(side):
                    replay.launch()

        def join():
            if side is not None:
                cur.wait_stream(side)

        self._join = join
        return advance, self.n_iter // unroll

    def _solve_fused(self, ev: torch.Tensor, x0: torch.Tensor) -> torch.Tensor:
        advance, n_calls = self._plan_fused(ev, x0)
        for _ in range(n_calls):
            advance()
        self._join()
        if self._hist_dev is not None:
            self.history["loss"] = self._hist_dev.cpu().tolist()
        return x0

    # -- operator-level path: the reference composition with torch autograd + torch.optim.Adam ------------
    def _solve_operators(self, ev: torch.Tensor, x0: torch.Tensor) -> torch.Tensor:
        x0 = x0.clone().requires_grad_()
        optimizer = torch.optim.Adam([x0], lr=self.lr)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, max(self.n_iter, 1), 0.1)
        imager = self.orig_imager if self.padding == 0 else type(self.orig_imager)(self.orig_image_shape, self.padding)
        for _ in range(self.n_iter):
            optimizer.zero_grad()
            warped, _ = self.orig_warper.warp_event(ev, x0, "dense-flow", direction=self.warp_direction)
            iwe = imager.create_iwe(warped, method="bilinear_vote", sigma=self.blur_sigma)
            loss = self.cost_func.calculate({"iwe": iwe, "flow": x0, "weights": 1.0, "omit_boundary": self.omit_boundary})
            if self.store_history:
                self.history["loss"].append(float(loss))
            loss.backward()
            optimizer.step()
            scheduler.step()
        return x0.detach()
