"""event_based_bos_b200: B200-native (sm_100a CUDA) implementation of the contrast-maximisation inner
loop of tub-rip/event_based_bos, behind the reference's Python API.

    from event_based_bos_b200 import Warp, EventImageConverter, costs, solver

Layout mirrors the reference's `src/` package: `warp`, `event_image_converter`, `costs`, `solver`,
`types`, `utils`; `ops` is the functional layer over the C-ABI library `libebos.so` (include/ebos.h).
"""
from . import costs, event_image_converter, ops, solver, types, utils, warp  # noqa: F401
from .event_image_converter import EventImageConverter  # noqa: F401
from .warp import Warp  # noqa: F401

__version__ = "0.1.0"
