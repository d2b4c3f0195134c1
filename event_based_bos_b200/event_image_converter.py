"""Drop-in `EventImageConverter` (src/event_image_converter.py:20-620) on the CUDA kernels.

Only the operators on the contrast-maximisation path are provided: `create_iwe` /
`create_image_from_events_*` / `bilinear_vote_*` / `create_eventmask` (SURVEY.md section 8, rows
a8-a12).  The weighted-image variants (IWA/IWD/IWT, event-rate) have no caller upstream and are out
of scope.
"""
from __future__ import annotations

import logging
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import ops
from .types import FLOAT_TORCH, NUMPY_TORCH, is_numpy, is_torch, like_input, to_device_tensor

logger = logging.getLogger(__name__)

# floor() bias of the two upstream branches (src/event_image_converter.py:586 and :528)
TENSOR_FLOOR_BIAS = 1e-6
NUMPY_FLOOR_BIAS = 1e-8


class EventImageConverter(object):
    """Converter of events into images.

    Args:
        image_size (tuple) ... (H, W)
        outer_padding (int, or tuple) ... padding added on every side of the image.
        deterministic (bool) ... extension: use the sorted-splat mode whose result is bit-identical
            to the reference's sequential scatter_add_ (slower).  Default: atomic mode.
    """

    def __init__(self, image_size: tuple, outer_padding: Union[int, Tuple[int, int]] = 0, deterministic: bool = False):
        if isinstance(outer_padding, (int, float)):
            self.outer_padding = (int(outer_padding), int(outer_padding))
        else:
            self.outer_padding = outer_padding
        self.image_size = tuple(int(i + p * 2) for i, p in zip(image_size, self.outer_padding))
        self.deterministic = deterministic

    def update_property(self, image_size: Optional[tuple] = None,
                        outer_padding: Optional[Union[int, Tuple[int, int]]] = None):
        # Upstream adds `p` (not 2p) here (src/event_image_converter.py:36-48); kept as is.
        if image_size is not None:
            self.image_size = image_size
        if outer_padding is not None:
            if isinstance(outer_padding, int):
                self.outer_padding = (outer_padding, outer_padding)
            else:
                self.outer_padding = outer_padding
        self.image_size = tuple(i + p for i, p in zip(self.image_size, self.outer_padding))

    # -- higher layer ---------------------------------------------------------------------------
    def create_iwe(self, events: NUMPY_TORCH, method: str = "bilinear_vote", sigma: int = 1) -> NUMPY_TORCH:
        """Image of warped events, [(b,) H, W] (src/event_image_converter.py:51-73)."""
        if is_numpy(events):
            return self.create_image_from_events_numpy(events, method, sigma=sigma)
        elif is_torch(events):
            return self.create_image_from_events_tensor(events, method, sigma=sigma)
        e = f"Non-supported type of events. {type(events)}"
        logger.error(e)
        raise RuntimeError(e)

    def create_eventmask(self, events: NUMPY_TORCH) -> NUMPY_TORCH:
        """[(b,) 1, H, W] bool: pixels touched by at least one event (src/event_image_converter.py:288-301)."""
        if is_numpy(events):
            return (0 != self.create_image_from_events_numpy(events, sigma=0))[..., None, :, :]
        elif is_torch(events):
            return (0 != self.create_image_from_events_tensor(events, sigma=0))[..., None, :, :]
        raise RuntimeError

    # -- lower layer ----------------------------------------------------------------------------
    def create_image_from_events_numpy(self, events: np.ndarray, method: str = "bilinear_vote",
                                       weight: Union[float, np.ndarray] = 1.0, sigma: int = 1) -> np.ndarray:
        """numpy branch (src/event_image_converter.py:332-370): float64 image, floor bias 1e-8, optional
        `scipy.ndimage.gaussian_filter(image, sigma)` -- evaluated on the GPU (`gaussian_filter_numpy`)."""
        if method == "count":
            raise NotImplementedError("method='count' is outside the contrast-maximisation path of this package")
        elif method == "bilinear_vote":
            image = self.bilinear_vote_numpy(events, weight=weight)
        elif method == "polarity":
            pos_flag = events[..., 3] > 0
            if is_numpy(weight):
                pos_image = self.bilinear_vote_numpy(events[pos_flag], weight=weight[pos_flag])
                neg_image = self.bilinear_vote_numpy(events[~pos_flag], weight=weight[~pos_flag])
            else:
                pos_image = self.bilinear_vote_numpy(events[pos_flag], weight=weight)
                neg_image = self.bilinear_vote_numpy(events[~pos_flag], weight=weight)
            image = np.stack([pos_image, neg_image], axis=-3)
        else:
            e = f"{method = } is not supported."
            logger.error(e)
            raise NotImplementedError(e)
        if sigma > 0:
            image = gaussian_filter_numpy(image, sigma)
        return image

    def create_image_from_events_tensor(self, events: torch.Tensor, method: str = "bilinear_vote",
                                        weight: FLOAT_TORCH = 1.0, sigma: int = 0) -> torch.Tensor:
        """tensor branch (src/event_image_converter.py:372-405).  `count` / `polarity` raise like
        upstream."""
        if method == "count":
            raise RuntimeError("Expected self.dtype to be equal to src.dtype (upstream's tensor 'count' is broken: "
                               "src/event_image_converter.py:497-500)")
        elif method == "bilinear_vote":
            image = self.bilinear_vote_tensor(events, weight=weight)
        else:
            e = f"{method = } is not implemented"
            logger.error(e)
            raise NotImplementedError(e)
        if sigma > 0:
            if len(image.shape) == 2:
                image = image[None, None, ...]
            elif len(image.shape) == 3:
                image = image[:, None, ...]
            image = gaussian_blur3(image, sigma)
        return torch.squeeze(image)

    def bilinear_vote_numpy(self, events: np.ndarray, weight: Union[float, np.ndarray] = 1.0) -> np.ndarray:
        """src/event_image_converter.py:503-560: float64 accumulation, bias 1e-8."""
        if type(weight) == np.ndarray:
            assert weight.shape == events.shape[:-1]
        ev = to_device_tensor(events).to(torch.float64)
        w = to_device_tensor(weight).to(torch.float64) if is_numpy(weight) else weight
        img = ops.iwe_splat(ev, self.image_size, self.outer_padding, w, self.deterministic, NUMPY_FLOOR_BIAS)
        return img.cpu().numpy().squeeze()

    def bilinear_vote_tensor(self, events: torch.Tensor, weight: FLOAT_TORCH = 1.0) -> torch.Tensor:
        """src/event_image_converter.py:562-620.  Differentiable w.r.t. event coordinates and weight."""
        if type(weight) == torch.Tensor:
            assert weight.shape == events.shape[:-1]
        ev = to_device_tensor(events)
        w = to_device_tensor(weight) if is_torch(weight) else weight
        img = ops.iwe_splat(ev, self.image_size, self.outer_padding, w, self.deterministic, TENSOR_FLOOR_BIAS)
        return like_input(img, events).squeeze()


def gaussian_blur3(image: torch.Tensor, sigma: float) -> torch.Tensor:
    """3x3 Gaussian with reflect padding on [b,1,H,W] -- the arithmetic of torchvision's
    `gaussian_blur(kernel_size=3, sigma)` that upstream calls (src/event_image_converter.py:399-404), as one stencil
    kernel (`ebos_blur3`); differentiable (the backward is the exact adjoint kernel)."""
    return ops.blur3(image, float(sigma))


def gaussian_filter_numpy(image: np.ndarray, sigma: float) -> np.ndarray:
    """`scipy.ndimage.gaussian_filter(image, sigma)` as the numpy branch applies it (src/event_image_converter.py:
    368-369): taps exp(-x^2 / 2 sigma^2) / sum truncated at 4 sigma, 'reflect' borders, along EVERY axis of the array --
    for a [2,H,W] polarity stack or a [b,H,W] batch upstream therefore also blurs across the leading axis, which is
    kept.  The [H,W] passes are the separable correlation kernel `ebos_sepconv2d`; the pass over a leading axis is a
    (tiny) mixing of planes on the device."""
    from . import eklt

    taps = eklt.gaussian_taps_scipy(float(sigma))
    if len(taps) > 127:
        raise ValueError(f"sigma = {sigma}: the separable kernel holds at most 127 taps (sigma <= 15.7)")
    img = to_device_tensor(image).to(torch.float64 if image.dtype != np.float32 else torch.float32)
    if img.dim() < 2:
        raise ValueError(f"expected an image [(...,) H, W], got shape {tuple(image.shape)}")
    lead = img.shape[:-2]
    flat = img.reshape((-1,) + tuple(img.shape[-2:]))
    radius = len(taps) // 2
    for axis in range(len(lead)):            # leading axes first, like scipy's axis order
        n = lead[axis]
        mix = np.zeros((n, n))
        for i in range(n):
            for u, t in enumerate(taps):
                j = i + u - radius
                while j < 0 or j >= n:       # scipy 'reflect': (d c b a | a b c d | d c b a)
                    j = -j - 1 if j < 0 else 2 * n - 1 - j
                mix[i, j] += t
        m = torch.from_numpy(mix).to(img)
        full = flat.reshape(tuple(lead) + tuple(img.shape[-2:]))
        full = torch.movedim(torch.tensordot(m, torch.movedim(full, axis, 0), dims=([1], [0])), 0, axis)
        flat = full.reshape((-1,) + tuple(img.shape[-2:]))
    out = torch.stack([eklt.sepconv2d(plane.contiguous(), taps, taps, "reflect") for plane in flat])
    return out.reshape(tuple(img.shape)).cpu().numpy()
