"""The EKLT level objective written with the SAME torch ops the reference uses, differentiated by autograd
(TEST INFRASTRUCTURE; CPU).

`oracle/spec_eklt.py` is the explicit numpy restatement with an analytic backward that the CUDA kernels are checked
against term by term.  This module is the other half: the reference's own op sequence
(`torch.nn.functional.pad` + bilinear `interpolate`, `grid_sample`, `conv2d`, `torch.gradient`, `torch.linalg.norm`,
autograd) re-typed from the cited lines, so that

  * the analytic oracle has an independent autograd cross-check that travels to the GPU box (the reference tree does not),
  * `bench.py --impl reference --workload eklt` / `cpu_baseline` time what the reference executes per iteration --
    float64 torch on all host threads -- instead of a single-threaded numpy loop.

Pinned against the reference's outputs by tests/test_eklt_host_math.py (tests/golden/reference_eklt_v1.npz).
This module must never be imported from event_based_bos_b200/ (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def sobel_over_8(p: torch.Tensor) -> torch.Tensor:
    """`poisson_to_flow`: SobelTorch(ksize=3, replicate padding) / 8 on [1,ph,pw] -> [2,ph,pw].
    src/solver/patch_eklt_dependent.py:259-281, src/utils/stat_utils.py:62-139."""
    kx = torch.tensor([[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]], dtype=p.dtype)
    x = F.pad(p[None], (1, 1, 1, 1), mode="replicate")
    dx = F.conv2d(x, kx[None, None])
    dy = F.conv2d(x, kx.t()[None, None])
    return torch.cat([dx, dy], dim=1)[0] / 8.0


def upsample_patch(flow_array: torch.Tensor, patch: int, image_size: Tuple[int, int]) -> torch.Tensor:
    """`interpolate_dense_flow_from_patch_tensor` (src/solver/patch_eklt.py:173-204): replicate pad, bilinear resize
    (torchvision `resize` = `interpolate(mode="bilinear", align_corners=False)`), centre crop."""
    c, ph, pw = flow_array.shape
    pad = int(patch / 2 // patch) + 1
    x = F.pad(flow_array[None], (pad, pad, pad, pad), mode="replicate")
    size = [x.shape[2] * patch, x.shape[3] * patch]
    dense = F.interpolate(x, size=size, mode="bilinear", align_corners=False)[0]
    cx, cy = dense.shape[1] // 2, dense.shape[2] // 2
    h1, w1 = cx - image_size[0] // 2, cy - image_size[1] // 2
    return dense[..., h1:h1 + image_size[0], w1:w1 + image_size[1]]


def warp_image_forward(image: torch.Tensor, translation: torch.Tensor) -> torch.Tensor:
    """grid_sample of `image` [H,W] displaced by `translation` [2,H,W] with the arithmetic of
    src/utils/frame_utils.py:75-86: the normalised base grid is an int64 `arange` divided by a Python float, i.e. a
    FLOAT32 tensor, and only becomes float64 when the (float64) translation is subtracted."""
    h, w = image.shape
    half_h, half_w = (h - 1) / 2.0, (w - 1) / 2.0
    base_rows = torch.arange(h) / half_h - 1          # float32 (int64 / float)
    base_cols = torch.arange(w) / half_w - 1
    grid_y = base_rows[:, None] - translation[0] / half_h
    grid_x = base_cols[None, :] - translation[1] / half_w
    grid = torch.stack([grid_x, grid_y], dim=-1)[None]
    return F.grid_sample(image[None, None], grid, mode="bilinear", align_corners=True)[0, 0]


def objective(theta: torch.Tensor, grad_x: torch.Tensor, grad_y: torch.Tensor, meas: torch.Tensor, winv: torch.Tensor,
              roi: Tuple[int, int, int, int], patch: int, w_data: float = 1.0, w_tv: float = 0.5, w_pxy: float = 0.1,
              poisson: bool = True, warp: bool = True, no_polarity: bool = False,
              weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`PatchEkltPyramid2._objective_scipy` (src/solver/patch_eklt_pyramid2.py:345-392) with the three cost terms of
    configs/hot_plate1.yaml: DifferenceNorm (costs/diff_norm.py:52), ImageGradient (costs/image_gradient.py:60-75),
    FlowNormPxy (costs/flow_norm.py:52).  theta [(1|2) + (0|2), ph, pw] float64, differentiable."""
    H, W = grad_x.shape
    mask = torch.zeros((H, W), dtype=theta.dtype)
    mask[roi[0]:roi[1], roi[2]:roi[3]] = 1
    nf = 1 if poisson else 2
    patch_flow = sobel_over_8(theta[[0]]) if poisson else theta[[0, 1]]
    dense_flow = upsample_patch(patch_flow, patch, (H, W))
    gradient_x, gradient_y = grad_x.clone(), grad_y.clone()
    loss_pxy = None
    if warp:
        translation = upsample_patch(theta[[nf, nf + 1]], patch, (H, W))
        gradient_x = warp_image_forward(gradient_x, translation)
        gradient_y = warp_image_forward(gradient_y, translation)
        loss_pxy = torch.linalg.norm(translation * mask, dim=0).mean()
    pred = dense_flow[0] * gradient_x + dense_flow[1] * gradient_y
    if no_polarity:
        pred = torch.abs(pred)
    if weights is not None:
        pred = pred * weights
    pred = pred / (torch.linalg.norm(pred.clone()) + 0.0001)
    pred = pred * mask
    loss = w_data * torch.linalg.norm(pred - meas, ord=1)
    roi_flow = dense_flow * mask
    gx = torch.gradient(roi_flow, dim=1)[0] * winv
    gy = torch.gradient(roi_flow, dim=2)[0] * winv
    loss = loss + w_tv * torch.mean(torch.abs(gx) + torch.abs(gy))
    if loss_pxy is not None:
        loss = loss + w_pxy * loss_pxy
    return loss


def value_and_grad(theta, *args, **kw):
    x = theta.detach().clone().requires_grad_()
    loss = objective(x, *args, **kw)
    loss.backward()
    return float(loss.detach()), x.grad
