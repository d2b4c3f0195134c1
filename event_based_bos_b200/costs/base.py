"""Objective base class.  Contract taken from the reference (src/costs/base.py:11-77), implementation ours:

* class attributes `name` (registry key) and `required_keys` (keys `calculate` reads from its argument dict);
* `direction` in {"minimize", "maximize", "natural"}: "minimize" returns the value an optimiser should decrease,
  the other two its negation;
* `calculate(arg: dict) -> scalar`, wrapped by the two decorators below, which subclasses apply as
  `@CostBase.register_history` / `@CostBase.catch_key_error` exactly like upstream's subclasses do;
* an optional per-call loss history (`store_history`, `get_history`, `clear_history`, `enable_/disable_history_register`).
"""
import functools
import logging
from typing import Callable, Dict, List

import torch

from ..types import FLOAT_TORCH

logger = logging.getLogger(__name__)

DIRECTIONS = ("minimize", "maximize", "natural")


def _guard_missing_keys(calculate: Callable) -> Callable:
    """A KeyError raised while reading `arg` is reported together with the keys the cost needs, then re-raised."""

    @functools.wraps(calculate)
    def guarded(self, arg: dict):
        try:
            return calculate(self, arg)
        except KeyError:
            logger.error("Input for the cost needs keys of:")
            logger.error(self.required_keys)
            raise

    return guarded


def _record_loss(calculate: Callable) -> Callable:
    """Append the returned value to `history["loss"]` while `store_history` is set (costs one `.item()` per call)."""

    @functools.wraps(calculate)
    def recorded(self, arg: dict):
        value = calculate(self, arg)
        if self.store_history:
            self.history["loss"].append(self.get_item(value))
        return value

    return recorded


class CostBase(object):
    name = "base"
    required_keys: List[str] = []

    # decorator names are part of the contract (subclasses of the reference's CostBase use them)
    catch_key_error = staticmethod(_guard_missing_keys)
    register_history = staticmethod(_record_loss)

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        if direction not in DIRECTIONS:
            e = f"direction should be minimize, maximize, and natural. Got {direction}."
            logger.error(e)
            raise ValueError(e)
        self.direction = direction
        self.store_history = store_history
        self.clear_history()

    # -- history ------------------------------------------------------------------------------------
    def clear_history(self) -> None:
        self.history: Dict[str, list] = {"loss": []}

    def get_history(self) -> dict:
        return dict(self.history)

    def enable_history_register(self) -> None:
        self.store_history = True

    def disable_history_register(self) -> None:
        self.store_history = False

    def get_item(self, loss: FLOAT_TORCH) -> float:
        return loss.item() if isinstance(loss, torch.Tensor) else loss

    # -- objective ----------------------------------------------------------------------------------
    def oriented(self, value):
        """`value` is the quantity to MINIMISE; other directions flip the sign (with upstream's warning)."""
        if self.direction == "minimize":
            return value
        logger.warning("The loss is specified as maximize direction")
        return -value

    @_record_loss
    @_guard_missing_keys
    def calculate(self, arg: dict) -> FLOAT_TORCH:
        raise NotImplementedError
