"""The C-ABI shared library: builds, loads without a GPU, exports exactly what include/ebos.h declares,
and the product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest
import torch

from event_based_bos_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ebos.h")).read()
    return sorted(set(re.findall(r"EBOS_API\s+[\w\s\*]+?\b(ebos_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 20
    for must in ("ebos_warp_dense_flow", "ebos_iwe_splat", "ebos_window_prepare", "ebos_window_splat",
                 "ebos_window_backward", "ebos_iwe_cost", "ebos_flow_tv", "ebos_adam_step", "ebos_cmax_value_and_grad"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = _build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ebos.h but not exported by libebos.so"
    # the ctypes prototypes cover the same set
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_version_and_error_string():
    lib = _capi.load()
    assert lib.ebos_version() == 200
    assert isinstance(_capi.last_error(), str)
    # argument validation happens before any CUDA call: usable without a device
    assert lib.ebos_time_stats(None, -1, 1, 0, None, None) == -1
    assert "bad argument" in _capi.last_error()
    assert lib.ebos_window_bytes(1000, 64, 96, 0) >= 5 * 4000 + 256
    assert lib.ebos_window_bytes(1000, 64, 96, 1) >= 4 * 8000 + 4000 + 256
    assert lib.ebos_window_bytes(1000, 64, 96, 0) % 256 == 0


def test_replay_slot_entries_validate_their_arguments_without_a_device():
    """The executable-graph slot API (ebos_capture_end_exec / ebos_exec_launch / ebos_exec_destroy): argument validation
    happens before any CUDA call, an empty slot cannot be launched, destroying an empty slot is a no-op."""
    lib = _capi.load()
    assert lib.ebos_exec_launch(None, None) == -1
    assert "no executable" in _capi.last_error()
    assert lib.ebos_exec_destroy(None) == 0
    from event_based_bos_b200 import ops

    slot = ops.ReplaySlot()
    assert slot.updates == 0 and slot.instantiations == 0
    slot.close()                                   # empty slot: nothing to free, no CUDA call
    with pytest.raises(RuntimeError, match="no executable"):
        slot.launch() if torch.cuda.is_available() else _capi.check(lib.ebos_exec_launch(None, None), "ebos_exec_launch")
    # the fused-TV iteration refuses what its kernel does not cover (nothing is enqueued): fp64, odd width
    for dtype, W in ((1, 96), (0, 97)):
        rc = lib.ebos_cmax_adam_iteration_fused_tv(1, 10, 0, 16, 32, 64, W, 0, 0, 2, 0, 1.0, 0.5, dtype, 48, 64, 80, 96, 112, 128,
                                                   144, 0.05, 0.9, 0.999, 1e-8, 160, 0.0, None, None)
        assert rc == -4 and "ebos_cmax_adam_iteration" in _capi.last_error()      # EBOS_ERR_UNSUPPORTED


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    import numpy as np

    import event_based_bos_b200 as ebos

    assert _capi.load().ebos_device_available() == 0
    ev = np.zeros((4, 4))
    with pytest.raises(RuntimeError, match="CUDA device"):
        ebos.Warp((4, 4), normalize_t=True).warp_event(ev, np.zeros((2, 4, 4)), "dense-flow")
    with pytest.raises(RuntimeError, match="CUDA device"):
        ebos.EventImageConverter((4, 4)).create_iwe(torch.zeros(4, 4), "bilinear_vote", sigma=0)
    with pytest.raises(RuntimeError, match="CUDA device"):
        ebos.ops.PreparedWindow(torch.zeros(4, 4), (4, 4))
    with pytest.raises(RuntimeError, match="CUDA device"):
        ebos.costs.functions["image_gradient"]().calculate({"flow": torch.zeros(2, 4, 4), "omit_boundary": False, "weights": 1.0})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "event_based_bos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src
