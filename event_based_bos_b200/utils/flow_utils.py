"""Flow helpers (host side, numpy)."""
from typing import Optional, Tuple

import numpy as np


def generate_dense_optical_flow(image_size: tuple, max_val: int = 30) -> np.ndarray:
    """Random flow [2,H,W] from numpy's global RNG (src/utils/flow_utils.py:20-30)."""
    return np.random.uniform(-max_val, max_val, (2,) + tuple(image_size))


def generate_uniform_optical_flow(image_size: tuple, x: int = 30, y: int = 30) -> np.ndarray:
    """Constant flow [2,H,W]; x is the row component (src/utils/flow_utils.py:33-45)."""
    return np.ones((2,) + tuple(image_size)) * np.array([x, y])[:, None, None]


def synthetic_flow(image_size: Tuple[int, int], seed: int = 0, max_val: float = 3.0, dtype=np.float32) -> np.ndarray:
    """Seeded incoherent flow U(-max_val, max_val): worst case for the gather."""
    rng = np.random.default_rng(seed + 1000003)
    return rng.uniform(-max_val, max_val, (2,) + tuple(image_size)).astype(dtype)


def smooth_flow(image_size: Tuple[int, int], seed: int = 0, max_val: float = 3.0, dtype=np.float32) -> np.ndarray:
    """Sum of 4 low-frequency sinusoids per channel, |flow| <= max_val (solve configs)."""
    H, W = image_size
    rng = np.random.default_rng(seed + 7919)
    r = np.arange(H)[:, None] / H
    c = np.arange(W)[None, :] / W
    flow = np.zeros((2, H, W))
    for ch in range(2):
        for _ in range(4):
            fr, fc = rng.uniform(0.5, 2.5, 2)
            ph = rng.uniform(0, 2 * np.pi, 2)
            flow[ch] += np.sin(2 * np.pi * fr * r + ph[0]) * np.cos(2 * np.pi * fc * c + ph[1])
    flow *= max_val / max(1e-12, np.abs(flow).max())
    return flow.astype(dtype)


def calculate_flow_error_numpy(flow_gt: np.ndarray, flow_pred: np.ndarray,
                               event_mask: Optional[np.ndarray] = None) -> dict:
    """Flow-error statistics between [b,2,H,W] flows, optionally restricted to pixels with events.

    Same metrics and keys as src/utils/flow_utils.py:769-821: EPE, the 1/2/3/5/10/20-pixel outlier
    ratios and AE (radians, angle between the (u,v,1) vectors), each summed over valid pixels (ground
    truth finite and non-zero in both channels, inside the event mask), divided by their count + 1e-5,
    and averaged over the batch."""
    assert len(flow_gt.shape) == len(flow_pred.shape) == 4
    finite = np.logical_and(~np.isinf(flow_gt[:, [0], ...]), ~np.isinf(flow_gt[:, [1], ...]))
    nonzero = np.logical_and(np.abs(flow_gt[:, [0], ...]) > 0, np.abs(flow_gt[:, [1], ...]) > 0)
    total_mask = np.logical_and(finite, nonzero)
    if event_mask is not None:
        total_mask = np.logical_and(event_mask, total_mask)
    gt_masked = flow_gt * total_mask
    pred_masked = flow_pred * total_mask
    n_points = np.sum(total_mask, axis=(1, 2, 3)) + 1e-5
    errors = {}
    endpoint_error = np.linalg.norm(gt_masked - pred_masked, axis=1)
    errors["EPE"] = np.mean(np.sum(endpoint_error, axis=(1, 2)) / n_points)
    for thr in (1, 2, 3, 5, 10, 20):
        errors[f"{thr}PE"] = np.mean(np.sum(endpoint_error > thr, axis=(1, 2)) / n_points)
    u, v = pred_masked[:, 0, ...], pred_masked[:, 1, ...]
    u_gt, v_gt = gt_masked[:, 0, ...], gt_masked[:, 1, ...]
    cosine = (1.0 + u * u_gt + v * v_gt) / (np.sqrt(1 + u * u + v * v) * np.sqrt(1 + u_gt * u_gt + v_gt * v_gt))
    errors["AE"] = np.mean(np.sum(np.arccos(cosine), axis=(1, 2)) / n_points)
    return errors
