"""Objective registry with the reference's layout (src/costs/__init__.py:1-24):
`functions[name]` -> cost class, plus `HybridCost`."""
from .base import CostBase
from .image_gradient import ImageGradient
from .iwe_costs import GradientMagnitude, ImageVariance

functions = {k.name: k for k in (ImageGradient, ImageVariance, GradientMagnitude)}

# For hybrid loss (kept outside `functions`, as upstream)
from .hybrid import HybridCost  # noqa: E402
