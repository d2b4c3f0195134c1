"""Data objectives on the IWE that the reference's configs name but its `src/costs` does not
contain: IWE variance and Sobel gradient magnitude.  Both follow the
`CostBase` contract so that `HybridCost` can mix them with `image_gradient`."""
import logging
from typing import Union

import torch

from .. import ops
from ..types import to_device_tensor
from .base import CostBase

logger = logging.getLogger(__name__)


class _IweCost(CostBase):
    required_keys = ["iwe", "omit_boundary"]
    kernel_name = ""

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> Union[float, torch.Tensor]:
        iwe = arg["iwe"]
        omit_boundary = arg["omit_boundary"]
        if not isinstance(iwe, torch.Tensor):
            e = f"Unsupported input type. {type(iwe)}."
            logger.error(e)
            raise NotImplementedError(e)
        dev = to_device_tensor(iwe)
        if dev.dtype not in (torch.float32, torch.float64):
            dev = dev.to(torch.float32)
        loss = ops.iwe_cost(dev, self.kernel_name, omit_boundary)   # minimise-direction value: -contrast
        if loss.device != iwe.device:
            loss = loss.to(iwe.device)
        return loss if self.direction == "minimize" else -loss


class ImageVariance(_IweCost):
    """minimize: -var(IWE) (unbiased); maximize/natural: +var(IWE)."""

    name = "image_variance"
    kernel_name = "image_variance"


class GradientMagnitude(_IweCost):
    """minimize: -mean(gx^2 + gy^2) with (gx, gy) = SobelTorch(ksize=3)(IWE)/8 (kernels and replicate
    padding of src/utils/stat_utils.py:62-139); maximize/natural: the positive value."""

    name = "gradient_magnitude"
    kernel_name = "gradient_magnitude"
