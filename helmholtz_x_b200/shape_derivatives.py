"""Adjoint sensitivity tail (helmholtz_x/shape_derivatives.py): boundary shape-derivative
integral  int_{ds(tag)} (V.n) div(conj(p_adj) c^2 grad p) ds  evaluated on the device.

The reference builds the displacement field V of every free-form-deformation control
point from a live gmsh model (shape_derivatives.py:39-77); gmsh is not a dependency here,
so V is an input (P1 nodal, n_nodes x 3)."""
import numpy as np
import torch

from . import _lib, fem
from .eigenvectors import normalize_adjoint
from .petsc4py_utils import conjugate_function  # noqa: F401  (re-export, shape_derivatives.py:1)


def boundary_shape_integral(mesh, degree, physical_facet_tag, V_nodal, p_dir, p_adj_normalized, c_nodal):
    """One control point: hx_shape_derivative over the facets tagged `physical_facet_tag`."""
    be = mesh.be
    Vs = fem.functionspace(mesh, ("Lagrange", degree))
    sel = be.asarray(np.flatnonzero(mesh.facet_tags == physical_facet_tag).astype(np.int32), dtype=torch.int32)
    Vd = be.asarray(np.ascontiguousarray(np.asarray(V_nodal, float).reshape(mesh.n_nodes, 3)), dtype=torch.float64)
    pd = be.asarray(np.asarray(p_dir, complex), dtype=torch.complex128)
    pa = be.asarray(np.asarray(p_adj_normalized, complex), dtype=torch.complex128)
    cd = be.asarray(np.ascontiguousarray(np.real(c_nodal))[:mesh.n_nodes], dtype=torch.float64)
    out = be.zeros(1)
    _lib.call("hx_shape_derivative", degree, int(sel.numel()), sel.data_ptr(), mesh.xd.data_ptr(), mesh.cellsd.data_ptr(),
              Vs.cell_dofs.data_ptr(), mesh.facetsd.data_ptr(), mesh.facet_cell().data_ptr(), Vd.data_ptr(), pd.data_ptr(),
              pa.data_ptr(), cd.data_ptr(), out.data_ptr(), be.stream)
    return complex(out.cpu().numpy()[0])


def shapeDerivativesFFD(geometry, lattice, physical_facet_tag, omega_dir, p_dir, p_adj, c, acousticMatrices, FlameMatrix,
                        displacement_fields=None):
    """shape_derivatives.py:12-37.  `displacement_fields[(zeta, phi)]` (or a callable
    (phi, zeta) -> array) supplies the FFD displacement vectors the reference takes from gmsh."""
    if displacement_fields is None:
        raise NotImplementedError("the FFD displacement fields need a live gmsh model in the reference; "
                                  "pass them as displacement_fields")
    mesh = geometry.mesh if hasattr(geometry, "mesh") else geometry
    p_adj_norm = normalize_adjoint(omega_dir, p_dir, p_adj, acousticMatrices, FlameMatrix)
    degree = acousticMatrices.degree
    cvals = c.x.array if isinstance(c, fem.Function) else np.asarray(c)
    derivatives = {}
    for zeta in range(0, lattice.n):
        derivatives[zeta] = {}
        for phi in range(0, lattice.m):
            V = displacement_fields(phi, zeta) if callable(displacement_fields) else displacement_fields[(zeta, phi)]
            derivatives[zeta][phi] = boundary_shape_integral(mesh, degree, physical_facet_tag, V, p_dir.x.array,
                                                             p_adj_norm.x.array, cvals)
    return derivatives
