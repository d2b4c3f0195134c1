"""`image_gradient`: total variation (L1) of the flow -- the smoothness regulariser
(src/costs/image_gradient.py:15-75), evaluated by the `ebos_flow_tv` kernel."""
import logging
from typing import Union

import numpy as np
import torch

from .. import ops
from ..types import to_device_tensor
from .base import CostBase

logger = logging.getLogger(__name__)


class ImageGradient(CostBase):
    """mean(|d flow/d row * w| + |d flow/d col * w|) with `torch.gradient` differences.

    Required keys: `flow` [2,H,W], `omit_boundary` (accepted and ignored, as upstream) and --
    although upstream does not list it -- `weights` (src/costs/image_gradient.py:50)."""

    name = "image_gradient"
    required_keys = ["flow", "omit_boundary"]

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32",
                 visualize_intermediate=False, *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> Union[float, torch.Tensor]:
        flow = arg["flow"]
        omit_boundary = arg["omit_boundary"]
        weights = arg["weights"]
        if isinstance(flow, torch.Tensor):
            return self.calculate_torch(flow, weights, omit_boundary)
        elif isinstance(flow, np.ndarray):
            # upstream has no numpy implementation: `self.calculate_numpy` does not exist
            raise AttributeError("'ImageGradient' object has no attribute 'calculate_numpy'")
        e = f"Unsupported input type. {type(flow)}."
        logger.error(e)
        raise NotImplementedError(e)

    def calculate_torch(self, flow: torch.Tensor, weights, omit_boundary: bool) -> torch.Tensor:
        dev_flow = to_device_tensor(flow)
        w = to_device_tensor(weights) if isinstance(weights, (torch.Tensor, np.ndarray)) else weights
        loss = ops.flow_total_variation(dev_flow, w)
        if loss.device != flow.device:
            loss = loss.to(flow.device)
        return self.oriented(loss)
