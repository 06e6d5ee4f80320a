"""Host-side containers on the CPU (HostBackend stands in for the device): meshes built from device tensors keep
their host copies lazy, device-backed coefficient fields materialise the host array only when read, the phase
timers attribute time to the innermost phase, and the prolongator filter keeps row sums."""
import numpy as np
import torch

from helmholtz_x_b200 import fem, phases
from helmholtz_x_b200.backend import CsrMatrix
from oracle.host_backend import HostBackend
from tests import cases


def test_mesh_from_tensors_keeps_host_arrays_lazy():
    be = HostBackend()
    m = cases.mesh("rijke3d")
    a = fem.Mesh(m.x, m.cells, m.cell_tags, m.facets, m.facet_tags, backend=be)
    assert set(a._host) == {"x", "cells", "cell_tags", "facets", "facet_tags"}           # numpy in: host kept
    b = fem.Mesh(a.xd, a.cellsd, a.cell_tagsd, a.facetsd, a.facet_tagsd, backend=be)
    assert b._host == {} and b.n_nodes == a.n_nodes and b.n_cells == a.n_cells
    assert np.array_equal(b.cells, a.cells) and "cells" in b._host and "x" not in b._host   # materialised on first read
    assert np.array_equal(b.geometry.x, a.x)
    assert b.cellsd.dtype == torch.int32 and b.xd.dtype == torch.float64
    try:
        b.no_such_attribute
    except AttributeError:
        pass
    else:
        raise AssertionError("unknown attributes must raise")


def test_device_backed_function_is_lazy_and_consistent():
    be = HostBackend()
    m = cases.mesh("rijke3d")
    mesh = fem.Mesh(m.x, m.cells, m.cell_tags, m.facets, m.facet_tags, backend=be)
    V = fem.DG0Space(mesh)
    vals = torch.arange(V.n, dtype=torch.float64)
    f = fem.Function.from_device(V, vals.clone(), name="soundspeed")
    assert f._x is None and f.real_device() is f._dev                  # nothing materialised
    g = f.copy().fill(1.4)
    assert g._x is None and float(g.real_device()[5]) == 1.4 and float(f.real_device()[5]) == 5.0
    arr = f.x.array                                                     # first host read
    assert arr.dtype == np.complex128 and arr[7] == 7.0 and f._dev is None
    f.x.array[7] = 3.0                                                  # the host copy is the truth from now on
    assert float(f.real_device()[7]) == 3.0
    h = f.copy()
    assert h.x.array[7] == 3.0 and h.x.array is not f.x.array
    h.fill(2.0)
    assert np.all(h.x.array == 2.0)


def test_phase_timers_attribute_time_to_the_innermost_phase():
    import time
    phases.enable(True)
    try:
        with phases.phase("outer"):
            time.sleep(0.02)
            with phases.phase("inner"):
                time.sleep(0.03)
            time.sleep(0.01)
        rep = phases.report()
    finally:
        phases.enable(False)
    assert set(rep) == {"outer", "inner"}
    assert 0.025 < rep["inner"] < 0.06 and 0.025 < rep["outer"] < 0.06      # exclusive times
    with phases.phase("off"):                                             # disabled: a no-op
        pass
    assert phases.report() == {}


def test_prolongator_filter_keeps_row_sums_and_drops_small_entries():
    from helmholtz_x_b200.amg import filter_prolongator
    rng = np.random.default_rng(0)
    n, nc = 200, 40
    rows = np.repeat(np.arange(n), 5)
    cols = rng.integers(0, nc, size=n * 5)
    key = np.unique(rows * nc + cols)
    rows, cols = key // nc, key % nc
    vals = rng.uniform(0.01, 1.0, len(key)) * rng.choice([1.0, 1.0, 0.02], len(key))
    ptr = np.zeros(n + 1, np.int64)
    ptr[1:] = np.cumsum(np.bincount(rows, minlength=n))
    P = CsrMatrix(n, nc, torch.from_numpy(ptr.astype(np.int32)), torch.from_numpy(cols.astype(np.int32)), torch.from_numpy(vals))
    F = filter_prolongator(P, 0.1)
    assert F.nnz < P.nnz and filter_prolongator(P, 0.0) is P
    ps, fs = P.to_scipy(), F.to_scipy()
    assert np.allclose(np.asarray(ps.sum(axis=1)).ravel(), np.asarray(fs.sum(axis=1)).ravel(), rtol=1e-13)
    rmax = np.asarray(abs(ps).max(axis=1).todense()).ravel()
    kept = fs.tocoo()
    assert np.all(np.abs(ps[kept.row, kept.col]).A1 >= 0.1 * rmax[kept.row] - 1e-15)
    dropped = (abs(ps) - abs(ps.multiply(fs != 0))).tocoo()
    assert np.all(dropped.data < 0.1 * rmax[dropped.row] + 1e-15)
