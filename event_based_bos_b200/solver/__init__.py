"""Solver registry with the reference's layout (src/solver/__init__.py:4-16)."""
from .base import SolverBase
from .contrast_maximization import ContrastMaximizationDense

# List of supported solvers
collections = {
    "contrast_maximization": ContrastMaximizationDense,
    "contrast_maximization_dense": ContrastMaximizationDense,
}
