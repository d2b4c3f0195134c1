"""Drop-in `Warp` with the reference's constructor and method signatures (src/warp.py:55-383),
backed by the CUDA kernels in libebos.so.

numpy arrays and CPU tensors are accepted exactly like upstream: they are moved to the GPU, run
through the same kernels, and the result comes back in the caller's container type.  There is no
CPU implementation -- without a CUDA device every call raises.
"""
from __future__ import annotations

import logging
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import ops
from .types import FLOAT_TORCH, NUMPY_TORCH, is_numpy, is_torch, like_input, nt_max, nt_min, to_device_tensor

logger = logging.getLogger(__name__)


class MotionModelKeyError(Exception):
    """Raised for an unsupported `motion_model` (src/warp.py:18-22)."""

    def __init__(self, message):
        logger.error(message)
        super().__init__(message)


def _feature_stub() -> dict:
    """The reference's disabled feature calculator always returns this dict of `None`s
    (FeatureCalculatorMock.skip, src/warp.py:30-38)."""
    return {
        "determinant": {"per_event": True, "value": None},
        "trace": {"per_event": True, "value": None},
        "divergence": {"per_event": True, "value": None},
        "straint": {"per_event": True, "value": None},
        "absement": {"per_event": False, "value": None},
    }


class Warp(object):
    """Warp functions class (same API as src/warp.py:55).

    Args:
        image_size (tuple[int, int]) ... (H, W).
        calculate_feature (bool) ... kept for signature parity; features are disabled upstream.
        normalize_t (bool) ... divide dt by the window's own time span.  Defaults to False.
        calib_param ... stored only.
        validate (bool) ... extension: check (with one device sync) that every event's integer
            pixel lies inside the flow grid and raise like upstream's torch.gather would.
    """

    def __init__(self, image_size: tuple, calculate_feature: bool = False, normalize_t: bool = False,
                 calib_param: Optional[np.ndarray] = None, validate: bool = True):
        self.update_property(image_size, calculate_feature, normalize_t, calib_param)
        self.validate = validate

    def update_property(self, image_size: Optional[tuple] = None, calculate_feature: Optional[bool] = None,
                        normalize_t: Optional[bool] = None, calib_param: Optional[np.ndarray] = None):
        if image_size is not None:
            self.image_size = image_size
        if calculate_feature is not None:
            self.calculate_feature = calculate_feature
        if normalize_t is not None:
            self.normalize_t = normalize_t
        if calib_param is not None:
            logger.info("Set camera matrix K.")
            self.calib_param = calib_param

    # -- motion-model bookkeeping (src/warp.py:95-165) ---------------------------------------
    def get_key_names(self, motion_model: str) -> list:
        if motion_model in ["dense-flow", "2d-translation", "rigid-optical-flow"]:
            return ["trans_x", "trans_y"]
        elif motion_model in ["scaler"]:
            return ["scaler"]
        raise MotionModelKeyError(f"{motion_model = } not supported")

    def get_motion_vector_size(self, motion_model: str) -> int:
        params = {k: 0.0 for k in self.get_key_names(motion_model)}
        return len(self.motion_model_to_motion(motion_model, params))

    def motion_model_to_motion(self, motion_model: str, params: dict) -> np.ndarray:
        if motion_model == "dense-flow":
            motion = np.array([params["trans_x"], params["trans_y"]])
            return self.get_flow_from_motion(motion, "2d-translation")
        elif motion_model in ["2d-translation", "rigid-optical-flow"]:
            return np.array([params["trans_x"], params["trans_y"]])
        elif motion_model in ["scaler"]:
            return np.array([params["scaler"]])
        raise MotionModelKeyError(f"{motion_model = } not supported")

    def motion_model_from_motion(self, motion: np.ndarray, motion_model: str) -> dict:
        if motion_model in ["dense-flow", "2d-translation", "rigid-optical-flow"]:
            return {"trans_x": motion[0], "trans_y": motion[1]}
        elif motion_model in ["scaler"]:
            return {"scaler": motion[0]}
        raise MotionModelKeyError(f"{motion_model = } not supported")

    def get_flow_from_motion(self, motion: NUMPY_TORCH, motion_model: str) -> NUMPY_TORCH:
        """Dense flow [2,H,W] that a parametric motion induces, obtained by warping one probe event
        per pixel at t=1 against a t=0 anchor (src/warp.py:167-190)."""
        H, W = self.image_size
        xs, ys = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        probes = np.stack([xs.ravel(), ys.ravel(), np.ones(H * W), np.ones(H * W)], axis=1).astype(np.float64)
        probes = np.concatenate([np.zeros((1, 4)), probes])
        ev: NUMPY_TORCH = torch.from_numpy(probes) if is_torch(motion) else probes
        warped, _ = self.warp_event(ev, motion, motion_model)
        u = -(warped[1:, 0] - ev[1:, 0]).reshape(self.image_size)[None, ...]
        v = -(warped[1:, 1] - ev[1:, 1]).reshape(self.image_size)[None, ...]
        if is_torch(motion):
            return torch.cat([u, v], dim=0)
        return np.concatenate([u, v], axis=0)

    # -- the operator ------------------------------------------------------------------------
    def warp_event(self, events: NUMPY_TORCH, motion: NUMPY_TORCH, motion_model: str,
                   direction: Union[str, float] = "first", flow_propagate_bin: Optional[int] = None
                   ) -> Tuple[NUMPY_TORCH, dict]:
        """Warp events to the reference time (src/warp.py:193-228).

        Inputs:
            events ... [(b,) n_events, 4] rows (x=row, y=col, t, p).
            motion ... [(b,) 2, H, W] for "dense-flow"; [2] for "2d-translation"/"rigid-optical-flow".
            direction ... 'first', 'middle', 'last', 'random', 'before', 'after' or a float.
        Returns:
            warped ... [(b,) n_events, 4] (x', y', dt, p), squeezed like upstream; differentiable
                w.r.t. a dense flow when given CUDA/CPU tensors.
            feature (dict) ... the reference's disabled-feature dict.
        """
        if motion_model == "dense-flow":
            return self.warp_event_from_optical_flow(events, motion, direction)
        elif motion_model in ["2d-translation", "rigid-optical-flow"]:
            assert motion.shape[-1] == 2
            return self.warp_event_2dof_xy(events, motion, direction)
        raise MotionModelKeyError(f"{motion_model = } not supported")

    def calculate_reftime(self, events: NUMPY_TORCH, direction: Union[str, float] = "first") -> FLOAT_TORCH:
        """Reference time (src/warp.py:230-262); plain min/max arithmetic on the caller's container."""
        if type(direction) is float:
            per = nt_max(events[..., 2], -1) - nt_min(events[..., 2], -1)
            return nt_min(events[..., 2], -1) + per * direction
        elif direction == "first":
            return nt_min(events[..., 2], -1)
        elif direction == "middle":
            return self.calculate_reftime(events, 0.5)
        elif direction == "last":
            return nt_max(events[..., 2], -1)
        elif direction == "random":
            return self.calculate_reftime(events, np.random.uniform(low=0.0, high=1.0))
        elif direction == "before":
            return self.calculate_reftime(events, -1.0)
        elif direction == "after":
            return self.calculate_reftime(events, 2.0)
        e = f"direction argument should be first, middle, last. Or float. {direction}"
        logger.error(e)
        raise ValueError(e)

    def calculate_dt(self, event: NUMPY_TORCH, reference_time: FLOAT_TORCH,
                     time_period: Optional[FLOAT_TORCH] = None) -> NUMPY_TORCH:
        """dt = t - reference_time, normalised by the span when `normalize_t` (src/warp.py:264-288)."""
        dt = event[..., 2] - reference_time
        if self.normalize_t:
            if time_period is None:
                time_period = nt_max(dt, -1) - nt_min(dt, -1)
            dt /= time_period[..., None]
        return dt

    def warp_event_from_optical_flow(self, event: NUMPY_TORCH, flow: NUMPY_TORCH,
                                     direction: Union[str, float] = "first") -> Tuple[NUMPY_TORCH, dict]:
        """x' = x - dt*flow[0, x, y], y' = y - dt*flow[1, x, y] (src/warp.py:292-342).

        NOTE: upstream passes the already computed reference time as third argument; here the
        kernel derives it from `direction` on the device, so the third argument is the direction."""
        if is_numpy(event):
            assert is_numpy(flow)
        elif is_torch(event):
            assert is_torch(flow)
        else:
            raise TypeError(f"Non-supported type of events. {type(event)}")
        ev = to_device_tensor(event)
        fl = to_device_tensor(flow).to(ev.dtype)
        assert ev.dim() + 1 == fl.dim() and ev.dim() in (2, 3)
        warped = ops.warp_dense_flow(ev, fl, self.image_size, direction, self.normalize_t, self.validate)
        return like_input(warped, event).squeeze(), _feature_stub()

    def warp_event_2dof_xy(self, event: NUMPY_TORCH, translation: NUMPY_TORCH,
                           direction: Union[str, float] = "first") -> Tuple[NUMPY_TORCH, dict]:
        """x' = x + dt*theta0, y' = y + dt*theta1 (src/warp.py:344-383; sign opposite to dense flow)."""
        if len(event.shape) == 1:
            event = event[None, :]
        ev = to_device_tensor(event)
        th = to_device_tensor(translation).to(ev.dtype)
        warped = ops.warp_2dof(ev, th, direction, self.normalize_t)
        return like_input(warped, event), _feature_stub()
