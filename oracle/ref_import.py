"""Import shim for the upstream reference (container-only; TEST INFRASTRUCTURE).

The reference tree (``/root/reference``, override with ``EBOS_REFERENCE_ROOT``) is pure
Python and importable here; it does not exist on the GPU box.  This module is used ONLY by
``oracle/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that pin the
restated spec (``oracle/spec.py``) to the live reference when the tree is mounted.

Nothing under ``event_based_bos_b200/`` may import this file.

``src.costs`` / ``src.solver`` / ``src.utils`` import packages that are absent in this image
and irrelevant to the numeric path (plotting, HDF5, PIV, optuna); they are stubbed with
``MagicMock`` so that the import succeeds.
"""
import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("EBOS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "openpiv", "openpiv.tools", "openpiv.pyprocess", "openpiv.validation", "openpiv.filters",
    "openpiv.scaling", "openpiv.windef", "openpiv.preprocess", "openpiv.smoothn",
    "optuna", "optuna.storages", "optuna.trial", "optuna.study",
    "ffmpeg", "plotly", "plotly.graph_objects", "matplotlib", "matplotlib.pyplot",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "skimage", "skimage.util", "skimage.metrics",
    "h5py", "hdf5plugin", "pivpy",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _install_stubs() -> None:
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        m = mock.MagicMock(name=name)
        m.__path__ = []  # behave like a package
        m.__name__ = name
        sys.modules[name] = m
    # src/utils/misc.py subclasses optuna.storages.InMemoryStorage: needs a real class.
    st = sys.modules.get("optuna.storages")
    if isinstance(st, mock.MagicMock):
        st.InMemoryStorage = type("InMemoryStorage", (object,), {})
        sys.modules["optuna"].storages = st


_ref_pkg = None


def load():
    """Return the reference ``src`` package (imported under the name ``ebos_reference_src``)."""
    global _ref_pkg
    if _ref_pkg is not None:
        return _ref_pkg
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    _install_stubs()
    # Import as a uniquely named top-level package so it cannot shadow anything called ``src``.
    spec = importlib.util.spec_from_file_location(
        "ebos_reference_src", os.path.join(REFERENCE_ROOT, "src", "__init__.py"),
        submodule_search_locations=[os.path.join(REFERENCE_ROOT, "src")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["ebos_reference_src"] = pkg
    spec.loader.exec_module(pkg)
    for sub in ("types", "warp", "event_image_converter", "utils", "costs", "solver"):
        setattr(pkg, sub, importlib.import_module(f"ebos_reference_src.{sub}"))
    _ref_pkg = pkg
    return pkg


def load_light():
    """Only ``warp``, ``event_image_converter`` and ``types`` (no stubs needed)."""
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    name = "ebos_reference_light"
    if name in sys.modules:
        return sys.modules[name]
    pkg = types.ModuleType(name)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
    sys.modules[name] = pkg
    for sub in ("types", "warp", "event_image_converter"):
        importlib.import_module(f"{name}.{sub}")
    return pkg
