"""`CostBase` with the reference's contract (src/costs/base.py:11-77): a `direction`, an optional
loss history, and `calculate(arg: dict) -> scalar`."""
import logging
from typing import Dict, List

import torch

from ..types import FLOAT_TORCH

logger = logging.getLogger(__name__)


class CostBase(object):
    """Base of the cost classes.

    Args:
        direction (str) ... 'minimize', 'maximize' or 'natural'.
        store_history (bool) ... append `loss.item()` to `history["loss"]` on every call (one device
            synchronisation per call, as upstream).
    """

    name = "base"
    required_keys: List[str] = []

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        if direction not in ["minimize", "maximize", "natural"]:
            e = f"direction should be minimize, maximize, and natural. Got {direction}."
            logger.error(e)
            raise ValueError(e)
        self.direction = direction
        self.store_history = store_history
        self.clear_history()

    def catch_key_error(func):
        """Log the required keys when the argument dict misses one, then re-raise."""

        def wrapper(self, arg: dict):
            try:
                return func(self, arg)
            except KeyError as e:
                logger.error("Input for the cost needs keys of:")
                logger.error(self.required_keys)
                raise e

        return wrapper

    def register_history(func):
        """Record the loss value when `store_history` is on."""

        def wrapper(self, arg: dict):
            loss = func(self, arg)
            if self.store_history:
                self.history["loss"].append(self.get_item(loss))
            return loss

        return wrapper

    def get_item(self, loss: FLOAT_TORCH) -> float:
        if isinstance(loss, torch.Tensor):
            return loss.item()
        return loss

    def clear_history(self) -> None:
        self.history: Dict[str, list] = {"loss": []}

    def get_history(self) -> dict:
        return self.history.copy()

    def enable_history_register(self) -> None:
        self.store_history = True

    def disable_history_register(self) -> None:
        self.store_history = False

    @register_history
    @catch_key_error
    def calculate(self, arg: dict) -> FLOAT_TORCH:
        raise NotImplementedError

    catch_key_error = staticmethod(catch_key_error)
    register_history = staticmethod(register_history)
