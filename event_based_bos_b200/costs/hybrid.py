"""`HybridCost`: a weighted sum of registered objectives (contract of src/costs/hybrid.py:12-79).

    HybridCost("minimize", {"diff_norm": 1.0, "image_gradient": 0.5, "flow_norm_pxy": 0.1})

Every term receives the same argument dict; a weight given as the string "inv" contributes 1 / term instead of
weight * term.  The per-term histories are reported next to the total under the term's registry name.
"""
import logging
from typing import Dict, List, Union

import torch

from .base import CostBase

logger = logging.getLogger(__name__)

Weight = Union[float, int, str]


class HybridCost(CostBase):
    name = "hybrid"

    def __init__(self, direction: str, cost_with_weight: Dict[str, Weight], store_history: bool = False, *args, **kwargs):
        from . import functions

        unknown = [k for k in cost_with_weight if k not in functions]
        if unknown:
            raise KeyError(f"unknown cost(s) {unknown}; registered: {sorted(functions)}")
        logger.info(f"Log functions are mix of {cost_with_weight}")
        self.terms: Dict[str, CostBase] = {
            key: functions[key](direction=direction, store_history=store_history, *args, **kwargs) for key in cost_with_weight}
        self.weights: Dict[str, Weight] = dict(cost_with_weight)
        super().__init__(direction=direction, store_history=store_history)
        self.required_keys: List[str] = [k for term in self.terms.values() for k in term.required_keys]

    @property
    def cost_func(self) -> Dict[str, dict]:
        """{name: {"func": cost object, "weight": weight}} -- the view upstream stores as an attribute."""
        return {k: {"func": self.terms[k], "weight": self.weights[k]} for k in self.terms}

    def update_weight(self, cost_with_weight: Dict[str, Weight]) -> None:
        assert set(cost_with_weight) == set(self.terms)
        self.weights.update(cost_with_weight)

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> Union[float, torch.Tensor]:
        total = 0.0
        for key, term in self.terms.items():
            value = term.calculate(arg)
            total = total + (1.0 / value if self.weights[key] == "inv" else self.weights[key] * value)
        return total

    # the history switches reach every term
    def clear_history(self) -> None:
        self.history = {"loss": []}
        for term in getattr(self, "terms", {}).values():
            term.clear_history()

    def get_history(self) -> dict:
        merged = dict(self.history)
        merged.update({key: term.get_history()["loss"] for key, term in self.terms.items()})
        return merged

    def enable_history_register(self) -> None:
        self.store_history = True
        for term in self.terms.values():
            term.store_history = True

    def disable_history_register(self) -> None:
        self.store_history = False
        for term in self.terms.values():
            term.store_history = False
