"""Solver registry with the reference's layout (src/solver/__init__.py:4-16)."""
from .base import SolverBase
from .contrast_maximization import ContrastMaximizationDense
from .patch_eklt_pyramid2 import PatchEkltPyramid2

# List of supported solvers
collections = {
    "contrast_maximization": ContrastMaximizationDense,
    "contrast_maximization_dense": ContrastMaximizationDense,
    "patch_eklt_pyramid2": PatchEkltPyramid2,      # src/solver/__init__.py:15
}
