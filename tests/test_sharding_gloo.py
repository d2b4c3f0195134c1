"""world_size-2 `gloo` runs of the multi-GPU host logic on CPU.  The numerical stages are the oracle's
CPU functions (the CUDA stages are covered by the gpu tests); what is checked here is the sharding,
the exchange pattern and that R-rank results equal the 1-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from event_based_bos_b200 import sharding
from oracle import spec

H, W, N = 20, 28, 3000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_objective(ev_local, tmm, cost="gradient_magnitude", tv_weight=0.5):
    """EventShardedObjective with oracle stages; dt uses the GLOBAL time range like the CUDA path."""
    def with_dt(flow):
        t_ref = tmm[0]
        dt = (ev_local[:, 2] - t_ref) / ((tmm[1] - t_ref) - (tmm[0] - t_ref))
        k = spec.origin_pixel_index(ev_local[:, 0], ev_local[:, 1], W)
        f = flow.reshape(2, -1)
        out = ev_local.clone()
        out[:, 0] = ev_local[:, 0] - dt * f[0][k]
        out[:, 1] = ev_local[:, 1] - dt * f[1][k]
        out[:, 2] = dt
        return out, dt, k

    def splat(flow):
        return spec.bilinear_vote(with_dt(flow)[0], (H, W))

    def cost_fn(iwe):
        a = iwe.clone().requires_grad_()
        loss = spec.DATA_COSTS[cost](a, False)
        loss.backward()
        return loss.detach().reshape(1), a.grad

    def backward(flow, g):
        warped, dt, k = with_dt(flow)
        inds, mask, _ = spec.vote_taps(warped[:, :2], (H, W))
        gg = g.reshape(-1)[inds] * mask
        n = len(k)
        fl = torch.floor(warped[:, :2] + 1e-6)
        a, b = warped[:, 0] - fl[:, 0], warped[:, 1] - fl[:, 1]
        dx = (1 - b) * (gg[n:2 * n] - gg[:n]) + b * (gg[3 * n:] - gg[2 * n:3 * n])
        dy = (1 - a) * (gg[2 * n:3 * n] - gg[:n]) + a * (gg[3 * n:] - gg[n:2 * n])
        d = flow.new_zeros(2, H * W)
        d[0].scatter_add_(0, k, -dt * dx)
        d[1].scatter_add_(0, k, -dt * dy)
        return d.reshape(2, H, W)

    def reg(flow):
        return tv_weight * spec.total_variation(flow, 1.0).reshape(1), tv_weight * spec.total_variation_grad(flow, 1.0)

    return sharding.EventShardedObjective(splat, cost_fn, backward, reg)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        ev = torch.from_numpy(spec.synthetic_events(N, (H, W), seed=4, dtype=np.float64))
        flow = torch.from_numpy(spec.synthetic_flow((H, W), seed=4, dtype=np.float64))
        # --- event sharding: partial IWE + partial gradient, two all-reduces
        s, e = sharding.shard_events(N)
        local = ev[s:e]
        tmm = sharding.global_time_range(local[:, 2]).double()
        # float32 exchange of the range is what the CUDA path does; use exact values for the fp64 check
        lo, hi = local[:, 2].min().reshape(1), local[:, 2].max().reshape(1)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert abs(float(tmm[0]) - float(lo)) < 1e-9 and abs(float(tmm[1]) - float(hi)) < 1e-9
        loss, grad = _oracle_objective(local, torch.cat([lo, hi])).value_and_grad(flow)
        # --- window sharding: 5 windows over 2 ranks, gathered
        flows = sharding.solve_windows(lambda w: torch.full((2, 3, 4), float(w)), 5, gather=True)
        assert sorted(flows) == [0, 1, 2, 3, 4] and all(float(flows[w][0, 0, 0]) == w for w in flows)
        mine = sharding.solve_windows(lambda w: torch.full((2, 3, 4), float(w)), 5, gather=False)
        assert sorted(mine) == list(range(rank, 5, world))
        # the gathered flows keep the solvers' dtype (float64 results must not be squeezed through float32), also when
        # a rank owns no window at all (1 window over 2 ranks)
        f64 = sharding.solve_windows(lambda w: torch.full((2, 3, 4), 1.0 + 2.0 ** -40, dtype=torch.float64), 1, gather=True)
        assert f64[0].dtype == torch.float64 and float(f64[0][0, 0, 0]) == 1.0 + 2.0 ** -40
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.numpy(), grad=grad.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ev = torch.from_numpy(spec.synthetic_events(N, (H, W), seed=4, dtype=np.float64))
    flow = torch.from_numpy(spec.synthetic_flow((H, W), seed=4, dtype=np.float64))
    ref_loss, ref_grad = spec.cmax_value_and_grad(ev, flow, (H, W), cost="gradient_magnitude", tv_weight=0.5)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_allclose(z["loss"][0], float(ref_loss), rtol=1e-12)
        np.testing.assert_allclose(z["grad"], ref_grad.numpy(), rtol=1e-9, atol=1e-15)
    a, b = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(a["grad"], b["grad"])  # every rank applies the identical update
