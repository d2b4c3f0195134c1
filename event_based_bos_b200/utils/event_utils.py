"""Event helpers (host side, numpy)."""
from typing import Optional, Tuple

import numpy as np


def generate_events(n_events: int, height: int, width: int, tmin: float = 0.0, tmax: float = 0.5,
                    dist: str = "uniform") -> np.ndarray:
    """Random events [n,4] (x=row, y=col, t sorted, p in {0,1}) drawn from numpy's global RNG.
    Same distribution and signature as src/utils/event_utils.py:18-47."""
    x = np.random.randint(0, height, n_events)
    y = np.random.randint(0, width, n_events)
    t = np.sort(np.random.uniform(tmin, tmax, n_events))
    p = np.random.randint(0, 2, n_events)
    return np.stack([x, y, t, p], axis=1).astype(np.float64)


def synthetic_events(n: int, image_size: Tuple[int, int], seed: int = 0, t_max: float = 1.0 / 120.0,
                     dtype=np.float32) -> np.ndarray:
    """Seeded uniform events of SURVEY.md section 8d (one Basler frame interval of Prophesee-shaped data)."""
    H, W = image_size
    rng = np.random.default_rng(seed)
    x = rng.integers(0, H, n)
    y = rng.integers(0, W, n)
    t = np.sort(rng.uniform(0.0, t_max, n))
    p = rng.integers(0, 2, n)
    return np.stack([x, y, t, p], axis=1).astype(dtype)


def synthetic_bos_events(n: int, image_size: Tuple[int, int], flow: np.ndarray, seed: int = 0, events_per_edge: int = 8,
                         t_max: float = 1.0 / 120.0, dtype=np.float32) -> np.ndarray:
    """Clustered, flow-consistent events ("BOS-like"): `n/events_per_edge` edge
    pixels of a random blob texture each emit `events_per_edge` events along x(t) = x0 + that*flow(x0),
    rounded to the pixel grid, so that the true flow is a contrast maximiser."""
    H, W = image_size
    rng = np.random.default_rng(seed)
    n_edges = max(1, n // events_per_edge)
    # blob texture: thresholded low-pass noise; its boundary pixels are the edges
    coarse = rng.standard_normal((H // 8 + 2, W // 8 + 2))
    tex = np.kron(coarse, np.ones((8, 8)))[:H, :W] > 0.0
    edge = np.zeros((H, W), dtype=bool)
    edge[:, 1:] |= tex[:, 1:] != tex[:, :-1]
    edge[1:, :] |= tex[1:, :] != tex[:-1, :]
    er, ec = np.nonzero(edge)
    pick = rng.integers(0, len(er), n_edges)
    r0 = np.repeat(er[pick], events_per_edge).astype(np.float64)
    c0 = np.repeat(ec[pick], events_per_edge).astype(np.float64)
    that = rng.uniform(0.0, 1.0, n_edges * events_per_edge)
    f0 = flow[0][er[pick], ec[pick]].repeat(events_per_edge)
    f1 = flow[1][er[pick], ec[pick]].repeat(events_per_edge)
    x = np.clip(np.rint(r0 + that * f0), 0, H - 1)
    y = np.clip(np.rint(c0 + that * f1), 0, W - 1)
    t = that * t_max
    order = np.argsort(t, kind="stable")
    p = rng.integers(0, 2, len(t))
    ev = np.stack([x, y, t, p], axis=1)[order]
    return ev[:n].astype(dtype)


def crop_event(events, x0: int, x1: int, y0: int, y1: int):
    """Events with x0 <= x < x1 and y0 <= y < y1; coordinates are NOT shifted
    (src/utils/event_utils.py:109-129)."""
    mask = (x0 <= events[..., 0]) * (events[..., 0] < x1) * (y0 <= events[..., 1]) * (events[..., 1] < y1)
    return events[mask]


def rebase_time(events: np.ndarray, t0: Optional[float] = None) -> np.ndarray:
    """Subtract the window start from the timestamps IN FLOAT64 (before any cast to fp32).

    Loader timestamps are absolute seconds (10-14 s in hot_plate1) with microsecond resolution; the
    fp32 ulp at 10 s is ~1 us, so a raw cast destroys the ordering inside a window.  The warp only
    depends on t through (t - t_ref)/period, which is invariant to this shift."""
    out = np.array(events, dtype=np.float64, copy=True)
    if len(out):
        out[:, 2] -= out[:, 2].min() if t0 is None else t0
    return out
