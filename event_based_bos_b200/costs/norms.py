"""The three remaining objectives of the reference's registry -- the terms `configs/hot_plate1.yaml` names next to
`image_gradient` (src/costs/diff_norm.py, flow_norm.py, flow_norm_pxy.py), as operator-level `CostBase` classes.

The EKLT solver (`solver.collections["patch_eklt_pyramid2"]`) evaluates these terms INSIDE its fused kernels
(csrc/ebos_eklt.cu); the classes here exist so that `HybridCost("minimize", cfg["cost_with_weight"])` built from the
shipped config works and can be applied to tensors directly.  They are thin: inputs are moved to the CUDA device and the
value is formed there with torch's norm ops (differentiable); numpy inputs return a float like upstream.
"""
import logging
from typing import Union

import numpy as np
import torch

from ..types import is_numpy, to_device_tensor
from .base import CostBase

logger = logging.getLogger(__name__)


def _on_device(x) -> torch.Tensor:
    if not isinstance(x, (torch.Tensor, np.ndarray)):
        e = f"Unsupported input type. {type(x)}."
        logger.error(e)
        raise NotImplementedError(e)
    return to_device_tensor(x)


def _back(value: torch.Tensor, proto):
    if is_numpy(proto):
        return float(value)
    return value if value.device == proto.device else value.to(proto.device)


class DifferenceNorm(CostBase):
    """`torch.linalg.norm(prediction - measurement, ord=1)` -- for the 2-D increment images the solver passes this is
    the MATRIX 1-norm, max_j sum_i |.|_ij (src/costs/diff_norm.py:52); `weights` must be
    present in the argument dict and is ignored, as upstream (:42)."""

    name = "diff_norm"
    required_keys = ["prediction", "measurement"]

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> Union[float, torch.Tensor]:
        prediction, measurement = arg["prediction"], arg["measurement"]
        arg["weights"]
        diff = _on_device(prediction) - _on_device(measurement)
        value = torch.linalg.norm(diff, ord=1)
        if is_numpy(prediction):
            return float(value)          # upstream's numpy branch returns +loss for every direction (:62-65)
        return self.oriented(_back(value, prediction))


class FlowNorm(CostBase):
    """mean over pixels of the per-pixel 2-norm of `flow` [2,H,W] (src/costs/flow_norm.py:52)."""

    name = "flow_norm"
    required_keys = ["flow"]
    key = "flow"

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> Union[float, torch.Tensor]:
        field = arg[self.key]
        value = torch.linalg.norm(_on_device(field), dim=0).mean()
        if is_numpy(field):
            return float(value)
        return self.oriented(_back(value, field))


class FlowNormPxy(FlowNorm):
    """The same norm on the translation field `pxy` [2,H,W] (src/costs/flow_norm_pxy.py)."""

    name = "flow_norm_pxy"
    required_keys = ["pxy"]
    key = "pxy"
