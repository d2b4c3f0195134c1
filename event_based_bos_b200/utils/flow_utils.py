"""Flow helpers: synthetic flows (host side, numpy) and the flow-error metrics (GPU reduction kernel)."""
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


ERROR_KEYS = ("EPE", "1PE", "2PE", "3PE", "5PE", "10PE", "20PE", "AE")


def _flow_error_device(flow_gt, flow_pred, event_mask, time_scale, tensor_variant: bool):
    """The eight metrics as a float64 [8] tensor on the device: one reduction kernel (ebos_flow_error)."""
    import torch

    from .. import _capi
    from ..ops import dtype_code
    from ..types import to_device_tensor

    _capi.require_device()
    if len(flow_gt.shape) != 4 or len(flow_pred.shape) != 4:
        raise AssertionError("flow_gt and flow_pred must be [B,2,H,W]")   # upstream asserts the same
    gt = to_device_tensor(flow_gt)
    if gt.dtype not in (torch.float32, torch.float64):
        gt = gt.to(torch.float64)
    pred = to_device_tensor(flow_pred).to(gt.dtype)
    gt, pred = torch.broadcast_tensors(gt, pred)
    gt, pred = gt.contiguous(), pred.contiguous()
    B, _, H, W = gt.shape
    mask = None
    if event_mask is not None:
        m = to_device_tensor(event_mask)
        mask = torch.broadcast_to(m.reshape((-1, 1) + tuple(m.shape[-2:])) if m.dim() != 4 else m, (B, 1, H, W))
        mask = (mask != 0).to(torch.uint8).contiguous()
    ts = None
    if time_scale is not None:
        ts = to_device_tensor(time_scale).to(gt.dtype).reshape(-1).contiguous()
        if ts.numel() != B:
            raise ValueError(f"time_scale must hold one value per batch row ({B}), got {ts.numel()}")
    lib = _capi.load()
    ws = torch.empty(lib.ebos_flow_error_workspace_doubles(B), dtype=torch.float64, device=gt.device)
    out = torch.empty(8, dtype=torch.float64, device=gt.device)
    _capi.check(lib.ebos_flow_error(_capi.ptr(gt), _capi.ptr(pred), _capi.ptr(mask), 1, _capi.ptr(ts), int(tensor_variant), B, H, W, dtype_code(gt),
                                    _capi.ptr(ws), _capi.ptr(out), _capi.current_stream()), "ebos_flow_error")
    return out, gt.dtype


def calculate_flow_error_tensor(flow_gt, flow_pred, event_mask=None, time_scale=None) -> dict:
    """src/utils/flow_utils.py:705-766 on the GPU: EPE, N-pixel outlier ratios and AE of [B,2,H,W] flows over the pixels
    whose ground truth is finite and non-zero in both channels (and inside `event_mask` [B,1,H,W]); `time_scale` [B,1]
    multiplies both flows.  Returns 0-dim tensors on the device (no host synchronisation), keys as upstream."""
    out, dtype = _flow_error_device(flow_gt, flow_pred, event_mask, time_scale, True)
    vals = out.to(dtype)
    return {k: vals[i] for i, k in enumerate(ERROR_KEYS)}


def calculate_flow_error_numpy(flow_gt: np.ndarray, flow_pred: np.ndarray,
                               event_mask: Optional[np.ndarray] = None) -> dict:
    """src/utils/flow_utils.py:769-821 with numpy in and floats out (what SolverBase.calculate_flow_error logs,
    src/solver/base.py:289-317); the reduction itself runs on the GPU (one kernel, one 64-byte read-back)."""
    out, _ = _flow_error_device(flow_gt, flow_pred, event_mask, None, False)
    vals = out.cpu().numpy()
    return {k: vals[i] for i, k in enumerate(ERROR_KEYS)}
