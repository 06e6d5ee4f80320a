"""The C-ABI library loads and exports every symbol include/hx_b200.h declares (no compute)."""
import ctypes
import os
import re

from helmholtz_x_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    lib = ctypes.CDLL(path)
    with open(os.path.join(ROOT, "include", "hx_b200.h")) as fh:
        header = fh.read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hx_[A-Za-z0-9_]+)\s*\(", header))
    assert len(declared) > 30
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/hx_b200.h but not exported: {missing}"
    # and the ctypes table covers exactly the header
    assert set(_lib.SIGNATURES) == declared


def test_version_and_error_text_without_gpu():
    lib = _lib.load()
    assert lib.hx_version() == 100
    assert isinstance(lib.hx_last_error(), bytes)
    assert _lib.call("hx_reduce_scratch_bytes", 8) > 0


def test_product_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from helmholtz_x_b200.backend import CudaBackend
    with pytest.raises(_lib.HxLibraryError):
        CudaBackend()


def test_host_colouring_is_valid():
    import numpy as np
    from tests import cases
    m = cases.mesh("rijke3d")
    cells = np.ascontiguousarray(m.cells, dtype=np.int32)
    color = np.empty(len(cells), np.int32)
    nc = _lib.call("hx_color_cells_h", len(cells), 4, cells.ctypes.data_as(ctypes.c_void_p), m.n_nodes,
                   color.ctypes.data_as(ctypes.c_void_p))
    assert 0 < nc <= 256
    for c in range(nc):
        nodes = cells[color == c].ravel()
        assert len(np.unique(nodes)) == len(nodes), "two cells of one colour share a vertex"


def test_every_entry_point_is_documented_in_integration_md():
    """INTEGRATION.md maps each C entry point to the reference call site it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "hx_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = set(re.findall(r"\b(hx_[A-Za-z0-9_]+)\s*\(", hdr))
    prefixes = [p[:-1] for p in re.findall(r"`(hx_[a-z0-9_]+\*)`", doc)]
    missing = [s for s in sorted(syms) if s not in doc and not any(s.startswith(p) for p in prefixes)]
    assert not missing, missing
