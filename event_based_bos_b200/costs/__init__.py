"""Objective registry with the reference's layout (src/costs/__init__.py:1-24): `functions[name]` -> cost class (every
cost of the reference's registry -- diff_norm, flow_norm, flow_norm_pxy, image_gradient -- plus the two IWE contrast
objectives the configs name but upstream lacks), and `HybridCost`."""
from .base import CostBase
from .image_gradient import ImageGradient
from .iwe_costs import GradientMagnitude, ImageVariance
from .norms import DifferenceNorm, FlowNorm, FlowNormPxy

functions = {k.name: k for k in (DifferenceNorm, FlowNorm, FlowNormPxy, ImageGradient, ImageVariance, GradientMagnitude)}

# For hybrid loss (kept outside `functions`, as upstream)
from .hybrid import HybridCost  # noqa: E402
