"""Event ingestion on the GPU: the raw sensor stream stays in HBM in its compact form and
windows are cut out, CROP-filtered and widened to the reference's event rows by one kernel pass.

Mirrors the part of `CcsDataLoader` (src/data_loader/ccs.py) the solve loop uses:
`len()`, `time_to_index(time)` (:345-357), `load_event(start, end)` (:288-296: rows = (y, x, t / 1e6, p), float64)
and the solver's CROP preprocessing (src/solver/base.py:123-139 -> src/utils/event_utils.py:109-129).
HDF5 decoding itself (h5py, hdf5plugin) is host I/O and stays outside this package: the constructor takes the
four `raw_events` arrays.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import check, current_stream, ptr


class RawEventStream:
    """x:int16 (sensor column), y:int16 (sensor row), t:int32 [us], p:bool -- 9 bytes per event on the device."""

    def __init__(self, x, y, t, p, device="cuda"):
        _capi.require_device()
        dev = torch.device(device)
        self.x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.int16)).to(dev)
        self.y = torch.as_tensor(np.ascontiguousarray(y, dtype=np.int16)).to(dev)
        self.t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.int32)).to(dev)
        self.p = torch.as_tensor(np.ascontiguousarray(p, dtype=np.bool_).view(np.uint8)).to(dev)
        if not (len(self.x) == len(self.y) == len(self.t) == len(self.p)):
            raise ValueError("x, y, t, p must have the same length")
        self._index = torch.zeros(1, dtype=torch.int64, device=dev)
        self._count = torch.zeros(1, dtype=torch.int64, device=dev)

    def __len__(self) -> int:
        return int(self.x.shape[0])

    def time_to_index(self, time: float) -> int:
        """searchsorted(t / 1e6, time) - 1  (src/data_loader/ccs.py:345-357)."""
        check(_capi.load().ebos_time_to_index(ptr(self.t), len(self), float(time), ptr(self._index), current_stream()),
              "ebos_time_to_index")
        return int(self._index.item())

    def load_event(self, start_index: int, end_index: int, crop: Optional[Tuple[int, int, int, int]] = None,
                   dtype: torch.dtype = torch.float64, rebase: bool = False) -> torch.Tensor:
        """Events [start_index, end_index) as a device tensor [n,4] = (row, col, t [s], p).

        crop = (xmin, xmax, ymin, ymax) in the event convention (x = row): the solver's CROP filter, order preserved.
        dtype float64 and rebase=False reproduce CcsDataLoader.load_event bit for bit; rebase=True measures time from
        the window's first event (integer microseconds subtracted before the division), needed for float32 rows."""
        start_index, end_index = int(start_index), int(end_index)
        if end_index > len(self):   # src/data_loader/ccs.py:251-254
            raise IndexError(f"Specified {start_index} to {end_index} index, but there are only {len(self)} events.")
        if len(self) <= start_index or end_index <= start_index:   # :283-285 and :263-266 (no events)
            raise IndexError(f"Specified {start_index} to {end_index} index, but no events.")
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"dtype must be float32 or float64, got {dtype}")
        n = end_index - start_index
        out = torch.empty((n, 4), dtype=dtype, device=self.x.device)
        lib = _capi.load()
        ws = None
        if crop is not None and n:
            ws = torch.empty(lib.ebos_ingest_workspace_bytes(n), dtype=torch.uint8, device=self.x.device)
        r0, r1, c0, c1 = (int(v) for v in crop) if crop is not None else (0, 0, 0, 0)
        t_origin = int(self.t[start_index].item()) if (rebase and n) else 0
        code = _capi.EBOS_F64 if dtype == torch.float64 else _capi.EBOS_F32
        check(lib.ebos_ingest_raw(ptr(self.x[start_index:]), ptr(self.y[start_index:]), ptr(self.t[start_index:]),
                                  ptr(self.p[start_index:]), n, int(crop is not None), r0, r1, c0, c1, t_origin,
                                  int(rebase), code, ptr(out), ptr(self._count), ptr(ws), 0 if ws is None else ws.numel(),
                                  current_stream()), "ebos_ingest_raw")
        if crop is None:
            return out
        return out[: int(self._count.item())]
