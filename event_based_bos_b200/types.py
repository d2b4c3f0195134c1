"""numpy/torch dispatch helpers with the reference's names (src/types/__init__.py:8-47)."""
from typing import Union

import numpy as np
import torch

NUMPY_TORCH = Union[np.ndarray, torch.Tensor]
FLOAT_TORCH = Union[float, torch.Tensor]


def is_torch(arg) -> bool:
    return isinstance(arg, torch.Tensor)


def is_numpy(arg) -> bool:
    return isinstance(arg, np.ndarray)


def nt_max(array: NUMPY_TORCH, dim: int) -> NUMPY_TORCH:
    if is_numpy(array):
        return array.max(axis=dim)
    return torch.max(array, dim).values


def nt_min(array: NUMPY_TORCH, dim: int) -> NUMPY_TORCH:
    if is_numpy(array):
        return array.min(axis=dim)
    return torch.min(array, dim).values


def to_device_tensor(array: NUMPY_TORCH, device=None) -> torch.Tensor:
    """Move a numpy array / CPU tensor onto the CUDA device the kernels run on."""
    from . import _capi

    _capi.require_device()  # no CPU implementation exists: fail loudly
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda"
    if is_numpy(array):
        return torch.from_numpy(np.ascontiguousarray(array)).to(device)
    return array.to(device)


def like_input(result: torch.Tensor, proto: NUMPY_TORCH) -> NUMPY_TORCH:
    """Return `result` in the container type / device of `proto` (numpy in -> numpy out)."""
    if is_numpy(proto):
        return result.detach().cpu().numpy()
    return result if result.device == proto.device else result.to(proto.device)
