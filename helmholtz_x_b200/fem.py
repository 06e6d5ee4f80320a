"""Mesh / function-space / Function containers and the assembly front-ends.

Stands where DOLFINx stands for the reference (mesh, functionspace, Function,
assemble_matrix, assemble_vector): integer preprocessing (dof maps, colouring,
pattern sizes) is orchestrated here with torch tensors on the device, the numerics
are libhx_b200 kernels (include/hx_b200.h, K1-K6).

DOF numbering: P1 dof i == mesh node i in the input (XDMF/meshio) order; P2 appends
one dof per edge, edges numbered by ascending (min vertex, max vertex).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .backend import CsrMatrix, CudaBackend
from .phases import phase

f64 = torch.float64
i32 = torch.int32
c128 = torch.complex128

TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
TRI_EDGES = ((0, 1), (0, 2), (1, 2))

_default_backend = None


def default_backend():
    global _default_backend
    if _default_backend is None:
        _default_backend = CudaBackend()
    return _default_backend


def set_default_backend(be):
    global _default_backend
    _default_backend = be


class _Topology:
    def __init__(self, dim):
        self.dim = dim


class _Geometry:
    def __init__(self, mesh):
        self._mesh = mesh
        self.dim = 3

    @property
    def x(self):
        return self._mesh.x


class MeshTags:
    """Subset of dolfinx.mesh.MeshTags: .indices, .values, .find(tag)."""

    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.int32)
        self.indices = np.arange(len(self.values), dtype=np.int32)

    def find(self, tag):
        return self.indices[self.values == tag]


class Mesh:
    """Tetrahedral mesh.  Every array exists on the device (.xd, .cellsd, .cell_tagsd, .facetsd) and on
    the host (.x, .cells, .cell_tags, .facets, .facet_tags as numpy).  The constructor takes either: numpy
    arrays are uploaded, device tensors (the per-rank sub-meshes of a multi-GPU run are cut on the device)
    are kept and their host copies appear only when somebody reads them."""

    _LAZY = {"x": "xd", "cells": "cellsd", "cell_tags": "cell_tagsd", "facets": "facetsd", "facet_tags": "facet_tagsd"}

    def __init__(self, x, cells, cell_tags=None, facets=None, facet_tags=None, backend=None):
        self.be = be = backend or default_backend()
        self._host = {}

        def both(name, a, dtype_np, dtype_t, shape):
            if torch.is_tensor(a):
                return a.to(be.device).to(dtype_t).reshape(shape).contiguous()
            h = np.ascontiguousarray(a, dtype=dtype_np).reshape(shape)
            self._host[name] = h
            return be.asarray(h, dtype=dtype_t)
        self.xd = both("x", x, np.float64, f64, (-1, 3))
        self.cellsd = both("cells", cells, np.int32, i32, (-1, 4))
        self.n_nodes, self.n_cells = int(self.xd.shape[0]), int(self.cellsd.shape[0])
        self.cell_tagsd = both("cell_tags", np.zeros(self.n_cells, np.int32) if cell_tags is None else cell_tags, np.int32, i32, (-1,))
        self.facetsd = both("facets", np.zeros((0, 3), np.int32) if facets is None else facets, np.int32, i32, (-1, 3))
        self.facet_tagsd = both("facet_tags", np.zeros(int(self.facetsd.shape[0]), np.int32) if facet_tags is None else facet_tags,
                                np.int32, i32, (-1,))
        self.geometry = _Geometry(self)
        self.topology = _Topology(3)
        self._spaces = {}
        self._cell_colors = None
        self._facet_colors = {}
        self._volumes = None
        self._facet_cell = None
        self._part = None
        self._dof_parts = {}

    def __getattr__(self, name):
        lazy = Mesh._LAZY.get(name)
        if lazy is None:
            raise AttributeError(name)
        host = self.__dict__["_host"]
        if name not in host:
            host[name] = self.__dict__[lazy].cpu().numpy()
        return host[name]

    def partition(self):
        """Row partition of this mesh over the process group (dist.Partition) with the
        rank's sub-mesh attached as .local_mesh; None when running on one GPU."""
        from . import dist as hxdist
        world, rank = hxdist.world_info()
        if world == 1:
            return None
        if self._part is None:
            import os
            with phase("partition"):
                part = hxdist.Partition(self.xd, self.cellsd, world, rank, os.environ.get("HX_PARTITION", "morton"), self.facetsd)
                part.local_mesh = Mesh(self.xd[part.dev("l2g")], part.dev("local_cells"), self.cell_tagsd[part.dev("cell_ids")],
                                       part.dev("local_facets"), self.facet_tagsd[part.dev("facet_ids")], backend=self.be)
            self._part = part
        return self._part

    def dof_partition(self, degree):
        """Row partition of the degree-`degree` space over the process group: the node partition for degree 1,
        dist.DofPartition (vertex + edge dofs) for degree 2; None on one GPU."""
        part = self.partition()
        if part is None or degree == 1:
            return part
        if degree not in self._dof_parts:
            from . import dist as hxdist
            with phase("partition"):
                self._dof_parts[degree] = hxdist.DofPartition(part, functionspace(self, ("Lagrange", degree)),
                                                              functionspace(part.local_mesh, ("Lagrange", degree)))
        return self._dof_parts[degree]

    # -- colouring (device: Jones-Plassmann rounds with fixed priorities, hx_color_cells) ------------
    @staticmethod
    def _color(entities_d, n_nodes, be):
        """entities_d: (n, nv) int32 device tensor.  Returns (#colours, host int64 colour pointer,
        device int32 entity order grouped by colour, ascending entity index inside a colour)."""
        ent = entities_d.contiguous()
        n, nv = int(ent.shape[0]), int(ent.shape[1])
        if n == 0:
            return 0, np.zeros(1, np.int64), be.zeros(0, dtype=i32), be.zeros(0, dtype=i32)
        st = be.stream
        count = be.zeros(n_nodes, dtype=i32)
        _lib.call("hx_dof_cell_count", n, nv, ent.data_ptr(), n_nodes, count.data_ptr(), st)
        adj_ptr = be.zeros(n_nodes + 1, dtype=i32)
        adj_ptr[1:] = torch.cumsum(count, 0)
        cursor = be.zeros(n_nodes, dtype=i32)
        adj = be.empty(n * nv, dtype=i32)
        _lib.call("hx_dof_cell_fill", n, nv, ent.data_ptr(), n_nodes, adj_ptr.data_ptr(), cursor.data_ptr(), adj.data_ptr(), st)
        ca = torch.full((n,), -1, dtype=i32, device=be.device)
        cb = torch.empty_like(ca)
        remaining = torch.zeros(1, dtype=torch.int64, device=be.device)
        for _ in range(512):
            _lib.call("hx_color_cells", n, nv, ent.data_ptr(), adj_ptr.data_ptr(), adj.data_ptr(), ca.data_ptr(),
                      cb.data_ptr(), 8, remaining.data_ptr(), st)
            if int(remaining) == 0:
                break
        else:
            raise _lib.HxError("hx_color_cells did not finish")
        if int(ca.min()) < 0:
            raise _lib.HxError("hx_color_cells: more than 128 colours needed")
        nc = int(ca.max()) + 1
        order = torch.sort(ca, stable=True).indices.to(i32)
        ptr = np.zeros(nc + 1, np.int64)
        ptr[1:] = torch.cumsum(torch.bincount(ca, minlength=nc), 0).cpu().numpy()
        return nc, ptr, order.contiguous(), ca

    def cell_colors(self):
        """(#colours, host colour pointer, device cell order grouped by colour)."""
        if self._cell_colors is None:
            with phase("colouring"):
                self._cell_colors = self._color(self.cellsd, self.n_nodes, self.be)
        return self._cell_colors[:3]

    def tagged_cell_colors(self, tag):
        """The same triple restricted to the cells carrying `tag` (flame zones are a small part of
        the mesh: the flame-vector kernels then visit those cells only)."""
        self.cell_colors()
        nc, _, _, ca = self._cell_colors
        sel = torch.nonzero(self.cell_tagsd == int(tag)).reshape(-1)
        col = ca[sel]
        order = torch.sort(col, stable=True).indices
        ptr = np.zeros(nc + 1, np.int64)
        ptr[1:] = torch.cumsum(torch.bincount(col, minlength=nc), 0).cpu().numpy()
        return nc, ptr, sel[order].to(i32).contiguous()

    def facet_colors(self, tag):
        if tag not in self._facet_colors:
            sel = self.be.asarray(np.flatnonzero(self.facet_tags == tag), dtype=torch.int64)
            nc, ptr, order, _ = self._color(self.facetsd[sel], self.n_nodes, self.be)
            self._facet_colors[tag] = (nc, ptr, sel[order.long()].to(i32).contiguous())
        return self._facet_colors[tag]

    def volumes(self):
        if self._volumes is None:
            v = self.be.empty(self.n_cells, dtype=f64)
            _lib.call("hx_cell_volumes", self.n_cells, self.xd.data_ptr(), self.cellsd.data_ptr(), v.data_ptr(), self.be.stream)
            self._volumes = v
        return self._volumes

    def facet_cell(self):
        """Owning cell of every tagged boundary facet (device int32)."""
        if self._facet_cell is None:
            n = self.n_nodes
            c = self.cellsd.long()
            faces = torch.cat([c[:, [1, 2, 3]], c[:, [0, 2, 3]], c[:, [0, 1, 3]], c[:, [0, 1, 2]]])
            owner = torch.arange(self.n_cells, device=c.device).repeat(4)
            fs = torch.sort(faces, dim=1).values
            kab, kc = fs[:, 0] * n + fs[:, 1], fs[:, 2]
            o1 = torch.sort(kc, stable=True).indices
            o2 = torch.sort(kab[o1], stable=True).indices
            order = o1[o2]
            kab_s, kc_s, own_s = kab[order], kc[order], owner[order]
            q = torch.sort(self.facetsd.long(), dim=1).values
            qab, qc = q[:, 0] * n + q[:, 1], q[:, 2]
            lo = torch.searchsorted(kab_s, qab)
            res = torch.full((q.shape[0],), -1, dtype=torch.int64, device=c.device)
            for off in range(32):
                p = (lo + off).clamp_max(kab_s.numel() - 1)
                hit = (kab_s[p] == qab) & (kc_s[p] == qc) & (res < 0)
                res = torch.where(hit, own_s[p], res)
            if q.shape[0] and int(res.min()) < 0:
                raise ValueError("a tagged facet has no owning cell")
            self._facet_cell = res.to(i32).contiguous()
        return self._facet_cell


class FunctionSpace:
    def __init__(self, mesh: Mesh, degree: int):
        if degree not in (1, 2):
            raise ValueError("only Lagrange degree 1 or 2 on tetrahedra")
        self.mesh, self.degree, self.be = mesh, degree, mesh.be
        nn = mesh.n_nodes
        if degree == 1:
            self.n = nn
            self.cell_dofs = mesh.cellsd
            self.facet_dofs = mesh.facetsd
            self.dof_coords = mesh.xd
            self.nd, self.nfd = 4, 3
        else:
            c = mesh.cellsd.long()
            ek = torch.stack([torch.minimum(c[:, a], c[:, b]) * nn + torch.maximum(c[:, a], c[:, b]) for a, b in TET_EDGES], 1)
            uniq, inv = torch.unique(ek.reshape(-1), return_inverse=True)
            self.n = nn + int(uniq.numel())
            self.cell_dofs = torch.cat([c, nn + inv.reshape(ek.shape)], 1).to(i32).contiguous()
            f = mesh.facetsd.long()
            if f.shape[0]:
                fk = torch.stack([torch.minimum(f[:, a], f[:, b]) * nn + torch.maximum(f[:, a], f[:, b]) for a, b in TRI_EDGES], 1)
                pos = torch.searchsorted(uniq, fk.reshape(-1))
                self.facet_dofs = torch.cat([f, nn + pos.reshape(fk.shape)], 1).to(i32).contiguous()
            else:
                self.facet_dofs = torch.zeros(0, 6, dtype=i32, device=c.device)
            e0, e1 = uniq // nn, uniq % nn
            self.edges = torch.stack([e0, e1], 1)
            self.dof_coords = torch.cat([mesh.xd, 0.5 * (mesh.xd[e0] + mesh.xd[e1])]).contiguous()
            self.nd, self.nfd = 10, 6
        self._pattern = None

    # -- K4: CSR pattern ---------------------------------------------------------------------
    def pattern(self):
        if self._pattern is None:
            with phase("pattern"):
                self._build_pattern()
        return self._pattern

    def _build_pattern(self):
        be, m = self.be, self.mesh
        st = be.stream
        n, nd, ncell = self.n, self.nd, m.n_cells
        count = be.zeros(n, dtype=i32)
        _lib.call("hx_dof_cell_count", ncell, nd, self.cell_dofs.data_ptr(), n, count.data_ptr(), st)
        adj_ptr = be.zeros(n + 1, dtype=i32)
        adj_ptr[1:] = torch.cumsum(count, 0)
        cursor = be.zeros(n, dtype=i32)
        adj = be.empty(ncell * nd, dtype=i32)
        _lib.call("hx_dof_cell_fill", ncell, nd, self.cell_dofs.data_ptr(), n, adj_ptr.data_ptr(), cursor.data_ptr(),
                  adj.data_ptr(), st)
        row_nnz = be.zeros(n, dtype=i32)
        _lib.call("hx_pattern_rows", n, nd, self.cell_dofs.data_ptr(), adj_ptr.data_ptr(), adj.data_ptr(),
                  row_nnz.data_ptr(), None, None, 0, st)
        if int(row_nnz.min()) < 0:
            raise _lib.HxError("pattern build: a dof touches too many cells for the row buffer")
        indptr = be.zeros(n + 1, dtype=torch.int64)
        indptr[1:] = torch.cumsum(row_nnz.long(), 0)
        nnz = int(indptr[-1])
        if nnz >= 2 ** 31:
            raise _lib.HxError("pattern build: nnz exceeds int32 indexing")
        indptr = indptr.to(i32).contiguous()
        indices = be.empty(nnz, dtype=i32)
        _lib.call("hx_pattern_rows", n, nd, self.cell_dofs.data_ptr(), adj_ptr.data_ptr(), adj.data_ptr(),
                  row_nnz.data_ptr(), indptr.data_ptr(), indices.data_ptr(), 1, st)
        self._pattern = (indptr, indices)

    def matrix(self, values):
        indptr, indices = self.pattern()
        return CsrMatrix(self.n, self.n, indptr, indices, values)

    @property
    def dofmap(self):
        return self

    def tabulate_dof_coordinates(self):
        return self.dof_coords.cpu().numpy()


def functionspace(mesh, element):
    """dolfinx.fem.functionspace(mesh, ("Lagrange"|"CG"|"DG", degree)) subset."""
    family, degree = element[0], int(element[1])
    if family in ("DG",) and degree == 0:
        return DG0Space(mesh)
    key = degree
    if key not in mesh._spaces:
        mesh._spaces[key] = FunctionSpace(mesh, degree)
    return mesh._spaces[key]


class DG0Space:
    def __init__(self, mesh):
        self.mesh, self.degree, self.n, self.be = mesh, 0, mesh.n_cells, mesh.be

    def tabulate_dof_coordinates(self):
        return self.mesh.x[self.mesh.cells].mean(axis=1)


class _Vec:
    """The few petsc4py.Vec methods the reference touches (petsc4py_utils.py)."""

    def __init__(self, array):
        self._a = array

    @property
    def array(self):
        return self._a

    def getArray(self):
        return self._a

    def setArray(self, v):
        self._a[:] = v

    def setValueLocal(self, i, v):
        self._a[i] = v

    def copy(self):
        return _Vec(self._a.copy())

    def scale(self, z):
        self._a *= z

    def conjugate(self):
        np.conjugate(self._a, out=self._a)

    def dot(self, other):
        """petsc4py Vec.dot conjugates its ARGUMENT: x.dot(y) = sum x_i conj(y_i)."""
        return complex(np.vdot(other._a, self._a))

    def __getitem__(self, i):
        return self._a[i]


class _X:
    def __init__(self, array):
        self.array = array
        self.petsc_vec = _Vec(array)

    def scatter_forward(self):
        pass


class Function:
    """dolfinx.fem.Function subset: .x.array (host numpy, writable), .name, .function_space."""

    def __init__(self, V, values=None, dtype=np.complex128, name="f"):
        self.function_space = V
        self.name = name
        self._dev = None
        arr = np.zeros(V.n, dtype=dtype) if values is None else np.array(values, dtype=dtype).reshape(V.n)
        self._x = _X(arr)

    @classmethod
    def from_device(cls, V, values_f64, name="f"):
        """A real-valued field computed on the device (a coefficient builder's output): the host
        array .x.array is only materialised when somebody asks for it."""
        f = cls.__new__(cls)
        f.function_space, f.name, f._dev, f._x = V, name, values_f64, None
        return f

    @property
    def x(self):
        if self._x is None:
            self._x = _X(self._dev.cpu().numpy().astype(np.complex128))
            self._dev = None               # the host copy is writable: it is the truth from now on
        return self._x

    def copy(self):
        if self._x is None:
            return Function.from_device(self.function_space, self._dev.clone(), self.name)
        return Function(self.function_space, self.x.array.copy(), self.x.array.dtype, self.name)

    def fill(self, value):
        """Set every entry (without materialising the host array of a device-backed field)."""
        if self._x is None:
            self._dev.fill_(float(np.real(value)))
        else:
            self._x.array[:] = value
        return self

    def real_device(self):
        if self._x is None:
            return self._dev
        return self.function_space.be.asarray(np.ascontiguousarray(self.x.array.real), dtype=f64)

    def interpolate(self, fn):
        pts = self.function_space.tabulate_dof_coordinates().T
        self.x.array[:] = fn(pts)


# ---------------------------------------------------------------------------------------------
# assembly front-ends
# ---------------------------------------------------------------------------------------------
def _field(V, f):
    """(device float64 tensor, is_dg0) of a coefficient given as Function / ndarray / tensor."""
    if isinstance(f, Function):
        is_dg0 = isinstance(f.function_space, DG0Space)
        return f.real_device(), is_dg0
    t = V.be.asarray(np.ascontiguousarray(np.real(f)) if not torch.is_tensor(f) else f, dtype=f64)
    return t, t.numel() == V.mesh.n_cells and t.numel() != V.mesh.n_nodes


def _p1_nodal(V, t):
    """Coefficient kernels take P1 nodal data; a P2 coefficient is restricted to the vertices."""
    return t[:V.mesh.n_nodes].contiguous()


def assemble_AC(V: FunctionSpace, c):
    """A = -int c^2 grad.grad, C = int phi phi (acoustic_matrices.py:101-103,121-123)."""
    be, m = V.be, V.mesh
    cd, dg0 = _field(V, c)
    if not dg0:
        cd = _p1_nodal(V, cd)
    indptr, indices = V.pattern()
    nnz = int(indices.numel())
    a = be.zeros(nnz, dtype=f64)
    cv = be.zeros(nnz, dtype=f64)
    ncol, ptr, order = m.cell_colors()
    with phase("assembly"):
        _lib.call("hx_assemble_AC", V.degree, m.n_cells, m.xd.data_ptr(), m.cellsd.data_ptr(), V.cell_dofs.data_ptr(),
                  cd.data_ptr(), int(dg0), ncol, ptr.ctypes.data_as(C.c_void_p), order.data_ptr(), indptr.data_ptr(),
                  indices.data_ptr(), a.data_ptr(), cv.data_ptr(), be.stream)
    return a, cv


def assemble_B(V: FunctionSpace, c, terms):
    """B = sum over (tag, coef) of coef * int_tag c phi phi ds (acoustic_matrices.py:71-109);
    complex values on the full cell pattern."""
    be, m = V.be, V.mesh
    cd, dg0 = _field(V, c)
    if not dg0:
        cd = _p1_nodal(V, cd)
    indptr, indices = V.pattern()
    b = be.zeros(int(indices.numel()), dtype=c128)
    fc = m.facet_cell() if dg0 else None
    for tag, coef in terms:
        ncol, ptr, order = m.facet_colors(tag)
        if ncol == 0:
            continue
        coef = complex(coef)
        cz = (C.c_double * 2)(coef.real, coef.imag)
        _lib.call("hx_assemble_B", V.degree, len(m.facets), m.xd.data_ptr(), m.facetsd.data_ptr(), V.facet_dofs.data_ptr(),
                  fc.data_ptr() if fc is not None else None, cd.data_ptr(), int(dg0), cz, ncol,
                  ptr.ctypes.data_as(C.c_void_p), order.data_ptr(), indptr.data_ptr(), indices.data_ptr(), b.data_ptr(),
                  be.stream)
    return b


def apply_dirichlet(V, values, dofs):
    be = V.be
    mask = be.zeros(V.n, dtype=torch.uint8)
    mask[be.asarray(dofs, dtype=torch.int64)] = 1
    indptr, indices = V.pattern()
    _lib.call("hx_apply_dirichlet", V.n, indptr.data_ptr(), indices.data_ptr(), mask.data_ptr(), values.data_ptr(),
              int(values.dtype == c128), be.stream)
    return values


def facet_integrals(mesh: Mesh, tag, f_nodal=None):
    """(area, int_tag f ds) for a P1 nodal field (acoustic_matrices.py:76-78,88-90)."""
    be = mesh.be
    sel = be.asarray(np.flatnonzero(mesh.facet_tags == tag), dtype=torch.int64)
    fac = mesh.facetsd[sel].contiguous()
    out = be.zeros(2, dtype=f64)
    fd = be.asarray(f_nodal, dtype=f64) if f_nodal is not None else None
    _lib.call("hx_facet_integrals", int(fac.shape[0]), mesh.xd.data_ptr(), fac.data_ptr(),
              fd.data_ptr() if fd is not None else None, out.data_ptr(), None, be.stream)
    o = out.cpu().numpy()
    return float(o[0]), float(o[1])


def flame_left(V, h, scale, gm1_nodal=None, gm1_const=0.0, tag=None):
    be, m = V.be, V.mesh
    hd, h_dg0 = _field(V, h)
    if not h_dg0:
        hd = _p1_nodal(V, hd)
    out = be.zeros(V.n, dtype=f64)
    # tag given: only the cells carrying it are visited (colour-ordered like the full list)
    ncol, ptr, order = m.cell_colors() if tag is None else m.tagged_cell_colors(tag)
    gd = None
    if gm1_nodal is not None:
        gd = _p1_nodal(V, be.asarray(gm1_nodal, dtype=f64))
    _lib.call("hx_flame_left", V.degree, m.n_cells, m.xd.data_ptr(), m.cellsd.data_ptr(), V.cell_dofs.data_ptr(),
              gd.data_ptr() if gd is not None else None, float(gm1_const), hd.data_ptr(), int(h_dg0), float(scale),
              None, 0, ncol, ptr.ctypes.data_as(C.c_void_p), order.data_ptr(), out.data_ptr(), be.stream)
    return out


def flame_right(V, w, rho):
    be, m = V.be, V.mesh
    wd = _p1_nodal(V, _field(V, w)[0])
    rd = _p1_nodal(V, _field(V, rho)[0])
    out = be.zeros(V.n, dtype=f64)
    ncol, ptr, order = m.cell_colors()
    _lib.call("hx_flame_right", V.degree, m.n_cells, m.xd.data_ptr(), m.cellsd.data_ptr(), V.cell_dofs.data_ptr(),
              wd.data_ptr(), rd.data_ptr(), ncol, ptr.ctypes.data_as(C.c_void_p), order.data_ptr(), out.data_ptr(), be.stream)
    return out


def locate_points(mesh: Mesh, points, tol=1e-10):
    be = mesh.be
    pts = be.asarray(np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3), dtype=f64)
    owner = be.zeros(pts.shape[0], dtype=i32)
    _lib.call("hx_locate_points", mesh.n_cells, mesh.xd.data_ptr(), mesh.cellsd.data_ptr(), int(pts.shape[0]),
              pts.data_ptr(), float(tol), owner.data_ptr(), be.stream)
    return owner, pts


def point_dphidz(V, pts, owner):
    be, m = V.be, V.mesh
    out = be.zeros(pts.shape[0], V.nd, dtype=f64)
    _lib.call("hx_point_dphidz", V.degree, m.xd.data_ptr(), m.cellsd.data_ptr(), int(pts.shape[0]), pts.data_ptr(),
              owner.data_ptr(), out.data_ptr(), be.stream)
    return out


def threshold(be, v, tol):
    _lib.call("hx_threshold", v.numel(), v.data_ptr(), float(tol), be.stream)
    return v
