"""Multi-GPU execution of the path: one process per GPU (torchrun), `torch.distributed` for plumbing.

Two regimes:

* **Window sharding** -- time windows of a sequence are independent (the reference carries no state
  between windows: src/solver/patch_eklt_pyramid2.py:187-190, bos_event.py:257-258 are commented out).
  Rank r solves windows r, r+R, ...; there is NO data-path collective, only an optional final gather.

* **Event sharding** of one giant window -- rank r owns a contiguous slice of the events; every
  objective evaluation builds a partial IWE, all-reduces it (sum, H*W fp32 = 3.7 MB at 1280x720),
  evaluates cost and dL/dIWE redundantly, runs the backward over its own events into a partial flow
  gradient and all-reduces that (7.4 MB), so that every rank applies the identical Adam step.
  The time normalisation uses the GLOBAL (min t, max t), exchanged once per window.

The numerical stages are injectable so that the host-side logic runs under `gloo` on CPU in the tests
(with oracle stages); in production they are the CUDA kernels of `ops`.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank_world() -> Tuple[int, int]:
    return (dist.get_rank(), dist.get_world_size()) if is_distributed() else (0, 1)


# ----------------------------------------------------------------------------------------------
# window sharding
# ----------------------------------------------------------------------------------------------
def shard_windows(n_windows: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """Indices of the windows rank `rank` solves: round robin (w mod R == r), which balances a
    sequence whose event rate drifts over time better than contiguous blocks."""
    r, R = rank_world()
    rank = r if rank is None else rank
    world = R if world is None else world
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_windows, world))


def solve_windows(solve_fn: Callable[[int], torch.Tensor], n_windows: int, gather: bool = False,
                  rank: Optional[int] = None, world: Optional[int] = None):
    """Run `solve_fn(window_index) -> flow [2,H,W]` for this rank's windows.

    Returns {window_index: flow} for the local windows; with `gather=True` every rank receives the
    flows of ALL windows (all_gather of equally sized blocks, padded with zeros for ranks that own
    one window fewer) -- the only collective of this regime, and off the solve path."""
    mine = shard_windows(n_windows, rank, world)
    local = {w: solve_fn(w) for w in mine}
    if not gather or not is_distributed():
        return local
    r, R = rank_world()
    per_rank = (n_windows + R - 1) // R
    proto = next(iter(local.values())) if local else None
    shape, dtype = _broadcast_shape(proto)
    device = proto.device if proto is not None else _default_device()
    block = torch.zeros((per_rank,) + shape, dtype=dtype, device=device)   # the solvers' own dtype (float64 flows stay float64)
    for j, w in enumerate(mine):
        block[j].copy_(local[w])
    blocks = [torch.empty_like(block) for _ in range(R)]
    dist.all_gather(blocks, block)
    out = {}
    for rr in range(R):
        for j, w in enumerate(range(rr, n_windows, R)):
            out[w] = blocks[rr][j]
    return out


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


_GATHER_DTYPES = (torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64)


def _broadcast_shape(proto: Optional[torch.Tensor]) -> Tuple[Tuple[int, ...], torch.dtype]:
    """All ranks must agree on the shape AND dtype of a window's result even if some own no window."""
    dev = proto.device if proto is not None else _default_device()
    shp = torch.zeros(6, dtype=torch.int64, device=dev)
    if proto is not None:
        if proto.dim() > 4 or proto.dtype not in _GATHER_DTYPES:
            raise TypeError(f"solve_windows(gather=True) cannot gather a {proto.dtype} tensor of rank {proto.dim()}")
        shp[0] = proto.dim()
        shp[1:1 + proto.dim()] = torch.tensor(proto.shape, device=dev)
        shp[5] = _GATHER_DTYPES.index(proto.dtype) + 1
    dist.all_reduce(shp, op=dist.ReduceOp.MAX)
    if int(shp[5]) == 0:
        raise RuntimeError("solve_windows(gather=True): no rank produced a result")
    return tuple(int(v) for v in shp[1:1 + int(shp[0])]), _GATHER_DTYPES[int(shp[5]) - 1]


# ----------------------------------------------------------------------------------------------
# event sharding
# ----------------------------------------------------------------------------------------------
def shard_events(n_events: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of the contiguous event slice owned by `rank` (sizes differ by at most one)."""
    r, R = rank_world()
    rank = r if rank is None else rank
    world = R if world is None else world
    base, rem = divmod(n_events, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_time_range(t_local: torch.Tensor) -> torch.Tensor:
    """[2] = (min t, max t) over the events of ALL ranks, in t's dtype (one MIN and one MAX all-reduce)."""
    kw = dict(device=t_local.device, dtype=t_local.dtype)
    lo = t_local.min().reshape(1) if t_local.numel() else torch.full((1,), float("inf"), **kw)
    hi = t_local.max().reshape(1) if t_local.numel() else torch.full((1,), float("-inf"), **kw)
    if is_distributed():
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return torch.cat([lo, hi])


class EventShardedObjective:
    """loss and dL/dflow of one window whose events are sharded over the ranks.

    Stages (callables, defaulting to the CUDA kernels):
        splat(flow)            -> partial IWE of the local events          [Hp,Wp]
        cost(iwe)              -> (data loss scalar tensor, dL/dIWE)        evaluated redundantly
        backward(flow, g_iwe)  -> partial dL/dflow of the local events      [2,H,W]
        regulariser(flow)      -> (tv loss scalar tensor, dTV/dflow)        identical on every rank
    """

    def __init__(self, splat: Callable, cost: Callable, backward: Callable, regulariser: Optional[Callable] = None):
        self.splat, self.cost, self.backward, self.regulariser = splat, cost, backward, regulariser

    def value_and_grad(self, flow: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        iwe = self.splat(flow)
        if is_distributed():
            dist.all_reduce(iwe, op=dist.ReduceOp.SUM)       # exchange 1: partial IWEs
        loss, g_iwe = self.cost(iwe)
        dflow = self.backward(flow, g_iwe)
        if is_distributed():
            dist.all_reduce(dflow, op=dist.ReduceOp.SUM)     # exchange 2: partial flow gradients
        if self.regulariser is not None:
            tv, dtv = self.regulariser(flow)
            loss = loss + tv
            dflow = dflow + dtv
        return loss, dflow


def cuda_event_sharded_objective(events_local: torch.Tensor, image_size: Tuple[int, int], cost: str = "gradient_magnitude",
                                 data_weight: float = 1.0, tv_weight: float = 0.0, omit_boundary: bool = False,
                                 direction="first", outer_padding: Tuple[int, int] = (0, 0),
                                 dtype=torch.float32) -> EventShardedObjective:
    """EventShardedObjective whose stages are the libebos kernels, for the local slice of the events."""
    from . import _capi, ops
    from ._capi import check, current_stream, ptr

    H, W = image_size
    ph, pw = outer_padding
    tmm = global_time_range(events_local[:, 2])
    window = ops.PreparedWindow(events_local, (H, W), direction, True, t_min_max=tmm, dtype=dtype)
    dev = events_local.device
    iwe = torch.empty((H + 2 * ph, W + 2 * pw), dtype=window.dtype, device=dev)
    g_iwe = torch.empty_like(iwe)
    dflow = torch.empty((2, H, W), dtype=window.dtype, device=dev)
    dtv = torch.empty_like(dflow)
    acc = torch.zeros(_capi.ACC_DOUBLES, dtype=torch.float64, device=dev)
    loss = torch.zeros(1, dtype=window.dtype, device=dev)
    kind = ops.COST_KINDS[cost]
    lib = _capi.load()

    def splat(flow):
        return ops.window_splat(window, flow, outer_padding, out=iwe)

    def cost_fn(img):
        st = current_stream()
        check(lib.ebos_iwe_cost(kind, ptr(img), H + 2 * ph, W + 2 * pw, int(omit_boundary), data_weight, window.code,
                                ptr(acc), ptr(g_iwe), st), "ebos_iwe_cost")
        acc[3] = 0.0
        acc[24:40] = 0.0
        check(lib.ebos_loss_finalize(kind, ptr(acc), H + 2 * ph, W + 2 * pw, H, W, int(omit_boundary), data_weight, 0.0,
                                     window.code, ptr(loss), st), "ebos_loss_finalize")
        return loss.clone(), g_iwe

    def backward(flow, g):
        dflow.zero_()
        check(lib.ebos_window_backward(ptr(window.buffer), window.n, window.flags, ptr(flow), H, W, ph, pw,
                                       window.code, ptr(g), kind, ptr(iwe), ptr(acc), int(omit_boundary), data_weight,
                                       ptr(dflow), current_stream()), "ebos_window_backward")
        return dflow

    def regulariser(flow):
        check(lib.ebos_flow_tv(ptr(flow), 0, H, W, tv_weight, window.code, ptr(acc), ptr(dtv), current_stream()),
              "ebos_flow_tv")
        return ((acc[3] + acc[24:40].sum()) * (tv_weight / (2.0 * H * W))).to(window.dtype).reshape(1), dtv

    # Exchange over peer memory (2..8 ranks of one NVLink domain): the partial IWE and the partial flow gradient live in
    # symmetric memory (every rank's buffer mapped into every process); cross-rank ordering by the symmetric-memory
    # device barriers.  Two forms:
    #   one-shot (R <= 2, gradient magnitude): the IWE reduction is fused into the cost kernel's tile load
    #            (ebos_iwe_cost_peers), the gradient reduction is one pass over the peers (ebos_sum_peers).  Measured on
    #            2 x B200 (profiles/tools/symm_probe.py): 24 / 30 us per exchange against 42 / 47 us for NCCL all-reduce
    #            at these sizes (3.7 / 7.4 MB, latency-bound).  Every rank reads R whole planes.
    #   two-shot (R >= 4, and the variance objective, which needs the reduced IWE materialised): reduce-scatter in place
    #            (rank r sums slice r of every peer's buffer into its own buffer: ebos_reduce_peers_slice), barrier,
    #            all-gather (ebos_gather_peers_slices): 2 (R-1)/R planes per rank instead of R-1 -- at R = 8 the
    #            one-shot gradient pass alone would pull 52 MB per rank over NVLink per evaluation.
    # EBOS_NO_P2P=1 or any failure to set it up falls back to NCCL; EBOS_P2P_FORM=1|2 forces a form (A/B runs).
    p2p = None
    iwe_full = None
    R0 = dist.get_world_size() if is_distributed() else 1
    if 2 <= R0 <= 8 and dev.type == "cuda" and dist.get_backend() == "nccl":
        import ctypes
        import os

        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            if os.environ.get("EBOS_NO_P2P"):
                raise RuntimeError("peer-memory exchange not requested")
            import torch.distributed._symmetric_memory as symm

            s_iwe = symm.empty(tuple(iwe.shape), dtype=window.dtype, device=dev)
            s_df = [symm.empty(tuple(dflow.shape), dtype=window.dtype, device=dev) for _ in range(2)]
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # every rank takes the same path
        if int(ok.item()):
            h_iwe = symm.rendezvous(s_iwe, dist.group.WORLD)
            h_df = [symm.rendezvous(b, dist.group.WORLD) for b in s_df]
            form = int(os.environ.get("EBOS_P2P_FORM", "0")) or (1 if R0 <= 2 else 2)
            if kind != _capi.COST_GRADMAG and form == 1:
                form = 2
            mc_iwe = int(getattr(h_iwe, "multicast_ptr", 0) or 0)
            mc_df = [int(getattr(h, "multicast_ptr", 0) or 0) for h in h_df]
            has_mc = (bool(mc_iwe and all(mc_df)) and window.dtype == torch.float32 and s_iwe.numel() % 4 == 0
                      and s_df[0].numel() % 4 == 0 and not os.environ.get("EBOS_NO_MULTIMEM"))
            if form == 3 and not has_mc:
                form = 2
            p2p = {"iwe": s_iwe, "df": s_df, "h_iwe": h_iwe, "h_df": h_df, "form": form, "count": 0,
                   "mc_iwe": mc_iwe, "mc_df": mc_df, "has_mc": has_mc,
                   "iwe_ptrs": (ctypes.c_void_p * R0)(*[int(v) for v in h_iwe.buffer_ptrs]),
                   "df_ptrs": [(ctypes.c_void_p * R0)(*[int(v) for v in h.buffer_ptrs]) for h in h_df]}
            iwe_full = iwe                                     # plain plane for the gathered IWE (two-shot form)
            iwe = s_iwe                                        # the splat writes the symmetric plane

    def _slice_bounds(n_elems: int, rank: int):
        """Slice of a plane owned by `rank` in the two-shot form (multiples of 4 elements: 16-byte vector accesses)."""
        per = ((n_elems + R0 - 1) // R0 + 3) // 4 * 4
        return per, min(rank * per, n_elems), min((rank + 1) * per, n_elems)

    side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def tv_beside_exchange(flow, target):
        """The TV kernel ((tv_weight / R) * dTV -> `target`, the buffer the backward accumulates on top of) needs only the
        flow: it runs on a side stream while this rank waits in the barriers of exchange 1, instead of in front of the
        splat.  Returns the stream to join before the backward."""
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            check(lib.ebos_flow_tv(ptr(flow), 0, H, W, tv_weight / max(R0, 1), window.code, ptr(acc), ptr(target),
                                   current_stream()), "ebos_flow_tv")
        return side

    class _Lean(EventShardedObjective):
        """Same result with fewer passes: the TV kernel writes (tv_weight / R) * dTV straight into the gradient buffer
        (every rank computes the identical TV term, the all-reduce over R ranks restores its full weight), the backward
        accumulates on top, and the scalar loss is formed once from the accumulators -- no separate TV plane, no
        zero-fill, no plane-sized add."""

        def value_and_grad(self, flow):
            if p2p is not None and p2p["form"] == 1:
                return self._value_and_grad_p2p(flow)
            if p2p is not None and p2p["form"] == 2:
                return self._value_and_grad_p2p_two_shot(flow)
            if p2p is not None and p2p["form"] == 3:
                return self._value_and_grad_multimem(flow)
            st = current_stream()      # form 0: NCCL all-reduce
            R = dist.get_world_size() if is_distributed() else 1
            Hp, Wp = H + 2 * ph, W + 2 * pw
            ops.window_splat(window, flow, outer_padding, out=iwe)
            tv_stream = tv_beside_exchange(flow, dflow)
            if R > 1:
                dist.all_reduce(iwe, op=dist.ReduceOp.SUM)       # exchange 1: partial IWEs
            check(lib.ebos_iwe_cost(kind, ptr(iwe), Hp, Wp, int(omit_boundary), data_weight, window.code, ptr(acc),
                                    ptr(g_iwe), st), "ebos_iwe_cost")
            torch.cuda.current_stream().wait_stream(tv_stream)
            check(lib.ebos_window_backward(ptr(window.buffer), window.n, window.flags, ptr(flow), H, W, ph, pw,
                                           window.code, ptr(g_iwe), kind, ptr(iwe), ptr(acc), int(omit_boundary),
                                           data_weight, ptr(dflow), st), "ebos_window_backward")
            if R > 1:
                dist.all_reduce(dflow, op=dist.ReduceOp.SUM)     # exchange 2: partial flow gradients (+ TV / R each)
            check(lib.ebos_loss_finalize(kind, ptr(acc), Hp, Wp, H, W, int(omit_boundary), data_weight, tv_weight,
                                         window.code, ptr(loss), st), "ebos_loss_finalize")
            return loss, dflow

        def _value_and_grad_p2p(self, flow):
            st = current_stream()
            R = R0
            Hp, Wp = H + 2 * ph, W + 2 * pw
            part, h_part, part_ptrs = p2p["df"][0], p2p["h_df"][0], p2p["df_ptrs"][0]
            ops.window_splat(window, flow, outer_padding, out=p2p["iwe"])
            tv_stream = tv_beside_exchange(flow, part)
            p2p["h_iwe"].barrier(channel=0)                   # every rank's partial IWE is complete
            check(lib.ebos_iwe_cost_peers(kind, p2p["iwe_ptrs"], R, Hp, Wp, int(omit_boundary), data_weight, window.code,
                                          ptr(acc), ptr(g_iwe), st), "ebos_iwe_cost_peers")
            torch.cuda.current_stream().wait_stream(tv_stream)
            check(lib.ebos_window_backward(ptr(window.buffer), window.n, window.flags, ptr(flow), H, W, ph, pw,
                                           window.code, ptr(g_iwe), kind, ptr(p2p["iwe"]), ptr(acc), int(omit_boundary),
                                           data_weight, ptr(part), st), "ebos_window_backward")
            # every partial gradient is complete -- and every rank is past its cost kernel, so the IWE planes are free
            h_part.barrier(channel=0)
            check(lib.ebos_sum_peers(part_ptrs, R, dflow.numel(), window.code, ptr(dflow), st), "ebos_sum_peers")
            h_part.barrier(channel=1)                         # nobody refills its gradient plane while a peer reads it
            check(lib.ebos_loss_finalize(kind, ptr(acc), Hp, Wp, H, W, int(omit_boundary), data_weight, tv_weight,
                                         window.code, ptr(loss), st), "ebos_loss_finalize")
            return loss, dflow

        def _value_and_grad_p2p_two_shot(self, flow):
            st = current_stream()
            R, rank = R0, dist.get_rank()
            Hp, Wp = H + 2 * ph, W + 2 * pw
            k = p2p["count"] % 2                              # the gradient planes alternate: a peer may still be
            p2p["count"] += 1                                 # gathering evaluation i while this rank starts i + 1
            part, h_part, part_ptrs = p2p["df"][k], p2p["h_df"][k], p2p["df_ptrs"][k]
            s_plane = p2p["iwe"]
            ops.window_splat(window, flow, outer_padding, out=s_plane)
            tv_stream = tv_beside_exchange(flow, part)
            # exchange 1 (partial IWEs): reduce-scatter in place, all-gather into a local plane
            n_iwe = s_plane.numel()
            per, b0, b1 = _slice_bounds(n_iwe, rank)
            p2p["h_iwe"].barrier(channel=0)                   # every rank's partial IWE is complete
            check(lib.ebos_reduce_peers_slice(p2p["iwe_ptrs"], R, b0, b1, window.code, ptr(s_plane), st), "ebos_reduce_peers_slice")
            p2p["h_iwe"].barrier(channel=1)                   # every slice is reduced
            check(lib.ebos_gather_peers_slices(p2p["iwe_ptrs"], R, n_iwe, per, window.code, ptr(iwe_full), st),
                  "ebos_gather_peers_slices")
            check(lib.ebos_iwe_cost(kind, ptr(iwe_full), Hp, Wp, int(omit_boundary), data_weight, window.code, ptr(acc),
                                    ptr(g_iwe), st), "ebos_iwe_cost")
            torch.cuda.current_stream().wait_stream(tv_stream)
            check(lib.ebos_window_backward(ptr(window.buffer), window.n, window.flags, ptr(flow), H, W, ph, pw,
                                           window.code, ptr(g_iwe), kind, ptr(iwe_full), ptr(acc), int(omit_boundary),
                                           data_weight, ptr(part), st), "ebos_window_backward")
            # exchange 2 (partial gradients + TV / R each); passing this barrier also means every rank has finished
            # gathering the IWE, so the symmetric IWE plane may be refilled by the next evaluation
            n_df = part.numel()
            per, b0, b1 = _slice_bounds(n_df, rank)
            h_part.barrier(channel=0)
            check(lib.ebos_reduce_peers_slice(part_ptrs, R, b0, b1, window.code, ptr(part), st), "ebos_reduce_peers_slice")
            h_part.barrier(channel=1)
            check(lib.ebos_gather_peers_slices(part_ptrs, R, n_df, per, window.code, ptr(dflow), st), "ebos_gather_peers_slices")
            check(lib.ebos_loss_finalize(kind, ptr(acc), Hp, Wp, H, W, int(omit_boundary), data_weight, tv_weight,
                                         window.code, ptr(loss), st), "ebos_loss_finalize")
            return loss, dflow

    def replayable(self_obj, flow, evaluations: int = 2):
        """A callable that replays `evaluations` captured evaluations of `flow` (a fixed tensor, updated in place by the
        caller between replays) as ONE executable graph -- the sharded evaluation is ~14 launches incl. four cross-rank
        barriers, issued from Python, and with eight ranks sharing a host the issue rate, not the GPU, can set the pace.
        Peer-memory forms only (the symmetric-memory barriers are ordinary kernels); `evaluations` must be even for the
        two-shot form (its gradient planes alternate; for the same reason do not put an ODD number of eager `value_and_grad`
        calls between two replays).  Every rank must call this; returns None on ALL ranks if any rank
        could not capture (the caller then keeps calling `value_and_grad`)."""
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        slot = None
        if p2p is None or p2p["form"] == 0 or evaluations % 2:
            ok.zero_()
        else:
            try:
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream())
                outs = []
                with torch.cuda.stream(cap):
                    slot = ops.ReplaySlot()
                    slot.capture(lambda: outs.extend(self_obj.value_and_grad(flow) for _ in range(evaluations)))
                torch.cuda.current_stream().wait_stream(cap)
            except Exception:
                slot = None
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            return None

        def replay():
            cur = torch.cuda.current_stream()
            cap.wait_stream(cur)
            with torch.cuda.stream(cap):
                slot.launch()
            cur.wait_stream(cap)
            return outs[-1]          # (loss, gradient) tensors of the last captured evaluation

        replay.slot = slot
        return replay

    def close(self_obj):
        """Release the symmetric-memory planes and their rendezvous handles NOW (every rank, after its last evaluation).
        The closures of this objective keep them alive until Python's cycle collector gets round to the local class --
        at an arbitrary later moment, and unmapping symmetric memory is illegal while any stream of the process is
        being captured (it aborted a later CUDA-graph capture of the same process in the 2-GPU bench)."""
        nonlocal iwe, iwe_full
        torch.cuda.synchronize(dev)
        if p2p is not None:
            p2p.clear()
        iwe = iwe_full = None
        self_obj.closed = True

    _Lean.replayable = replayable
    _Lean.close = close
    def _value_and_grad_multimem(self_obj, flow):
        """In-switch form (NVLS): rank r has the NVSwitch sum slice r of every rank's plane and write the sum back into every
        copy (ebos_multimem_allreduce_slice on the multicast mapping of the symmetric buffer): one pass between two barriers
        per exchange, no gather, 2/R of a plane per rank on the links.  The reduced IWE / gradient end up in the symmetric
        planes themselves."""
        st = current_stream()
        rank = dist.get_rank()
        Hp, Wp = H + 2 * ph, W + 2 * pw
        k = p2p["count"] % 2
        p2p["count"] += 1
        part, h_part, mc_part = p2p["df"][k], p2p["h_df"][k], p2p["mc_df"][k]
        s_plane = p2p["iwe"]
        ops.window_splat(window, flow, outer_padding, out=s_plane)
        tv_stream = tv_beside_exchange(flow, part)
        _, b0, b1 = _slice_bounds(s_plane.numel(), rank)
        p2p["h_iwe"].barrier(channel=0)                   # every rank's partial IWE is complete
        check(lib.ebos_multimem_allreduce_slice(p2p["mc_iwe"], b0, b1, window.code, st), "ebos_multimem_allreduce_slice")
        p2p["h_iwe"].barrier(channel=1)                   # every slice is reduced and written to every copy
        check(lib.ebos_iwe_cost(kind, ptr(s_plane), Hp, Wp, int(omit_boundary), data_weight, window.code, ptr(acc),
                                ptr(g_iwe), st), "ebos_iwe_cost")
        torch.cuda.current_stream().wait_stream(tv_stream)
        check(lib.ebos_window_backward(ptr(window.buffer), window.n, window.flags, ptr(flow), H, W, ph, pw,
                                       window.code, ptr(g_iwe), kind, ptr(s_plane), ptr(acc), int(omit_boundary),
                                       data_weight, ptr(part), st), "ebos_window_backward")
        _, b0, b1 = _slice_bounds(part.numel(), rank)
        h_part.barrier(channel=0)                         # every partial gradient is complete (and every rank is past its
        check(lib.ebos_multimem_allreduce_slice(mc_part, b0, b1, window.code, st), "ebos_multimem_allreduce_slice")   # IWE reads)
        h_part.barrier(channel=1)
        check(lib.ebos_loss_finalize(kind, ptr(acc), Hp, Wp, H, W, int(omit_boundary), data_weight, tv_weight,
                                     window.code, ptr(loss), st), "ebos_loss_finalize")
        return loss, part

    _Lean._value_and_grad_multimem = _value_and_grad_multimem
    obj = _Lean(splat, cost_fn, backward, regulariser if tv_weight else None)
    # Which exchange is fastest depends on the rank count and on what NCCL can do on the box (in-switch NVLS reductions):
    # measured on B200s, 128 Mi events -- 2 ranks: one-shot 0.593 / two-shot 0.607 / NCCL 0.625 ms; 4 ranks: one-shot
    # 0.415 / two-shot 0.371 / NCCL 0.368 ms.  Unless a form is forced, every available form is timed for a few
    # evaluations on this window (all ranks in lock-step, the slowest rank decides) and the fastest is kept.
    obj.exchange_autotune = None
    if p2p is not None and not __import__("os").environ.get("EBOS_P2P_FORM"):
        forms = ([1] if kind == _capi.COST_GRADMAG else []) + [2] + ([3] if p2p["has_mc"] else []) + [0]
        probe = torch.zeros((2, H, W), dtype=window.dtype, device=dev)
        timings = {}
        for form in forms:
            p2p["form"] = form
            for _ in range(2):
                obj.value_and_grad(probe)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier()
            a.record()
            for _ in range(5):
                obj.value_and_grad(probe)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timings[form] = float(t)
        best = min(timings, key=timings.get)
        if best == 0:
            # the peer-memory forms can be replayed as one executable graph (`replayable`: -2 % at 2 ranks, -6 % at 8, measured),
            # the NCCL form cannot: a peer form within 5 % of NCCL's eager time is the better choice
            peer = min((f for f in timings if f != 0), key=timings.get, default=None)
            if peer is not None and timings[peer] <= 1.05 * timings[0]:
                best = peer
        p2p["form"] = best
        names = {0: "nccl", 1: "peer one-shot", 2: "peer two-shot", 3: "in-switch multimem"}
        obj.exchange_autotune = {names[f]: round(v, 4) for f, v in timings.items()}
        if p2p["form"] == 0:
            pass
    if p2p is None or p2p["form"] == 0:
        obj.exchange = "nccl all-reduce (2 per evaluation)" if R0 > 1 else "none"
        if p2p is not None:
            obj.exchange += "; chosen by the start-up timing over the peer-memory forms"
    elif p2p["form"] == 3:
        obj.exchange = ("peer-memory in-switch (NVLS multimem.ld_reduce + multimem.st on the multicast mapping, one pass per plane; "
                        "4 device barriers per evaluation)")
    elif p2p["form"] == 1:
        obj.exchange = "peer-memory one-shot (IWE reduction fused into the cost kernel; 3 device barriers per evaluation)"
    else:
        obj.exchange = "peer-memory two-shot (reduce-scatter in place + all-gather for IWE and gradient; 4 device barriers per evaluation)"
    # kernels of this library per evaluation: TV, splat, cost, backward, loss + the peer kernels (1 one-shot, 4 two-shot);
    # the NCCL path uses two library all-reduces instead
    obj.launches_per_evaluation = 5 if (p2p is None or p2p["form"] == 0) else {1: 6, 2: 9, 3: 7}[p2p["form"]]
    return obj
