"""helmholtz_x/eigenvectors.py: normalize_eigenvector (a11), normalize_adjoint (a13)."""
import numpy as np
import torch

from .eigensolvers import EPS, PEP
from .fem import Function, assemble_AC, functionspace
from .petsc4py_utils import FixSign, matrix_vector, multiply, vector_matrix_vector
from .solver_utils import info, rank0


def _mass_form(V, matrices, vec):
    """p^T M p with the UNCONJUGATED mass form p*p*dx (eigenvectors.py:47): SpMV + dot on the device."""
    if matrices is not None and getattr(matrices, "ops", None) is not None and matrices.ops.part is not None:
        # multi-GPU: mass matrix rows of this rank (no Dirichlet rows), dot all-reduced
        ops = matrices.ops
        be = ops.be
        x = be.asarray(ops.to_local(np.asarray(vec, complex)), dtype=torch.complex128)
        y = be.zeros(ops.n)
        be.spmv(ops.space.matrix(matrices.C_nobc_values), x, y)
        out = be.zeros(2)
        be.multi_dot(x.view(1, -1), 1, y, out, conj=False)
        return complex(out[0].cpu().numpy())
    be = V.be
    if not hasattr(V, "_mass"):
        ones = np.ones(V.mesh.n_nodes)
        _, cvals = assemble_AC(V, ones)
        V._mass = V.matrix(cvals)
    x = be.asarray(np.asarray(vec, complex), dtype=torch.complex128)
    y = be.zeros(V.n)
    be.spmv(V._mass, x, y)
    out = be.zeros(2)
    be.multi_dot(x.view(1, -1), 1, y, out, conj=False)
    return complex(out[0].cpu().numpy())


def normalize_eigenvector(mesh, obj, i, absolute=False, degree=1, which='right', BlochRemapper=None, matrices=None,
                          print_eigs=True):
    """Extract eigenvector i, FixSign, scale so that int p p dx = 1 (eigenvectors.py:11-64)."""
    A = obj.getOperators()[0]
    vr, vi = A.createVecs()
    if isinstance(obj, EPS):
        eig = obj.getEigenvalue(i)
        omega = np.sqrt(eig)
        if which == 'right':
            obj.getEigenvector(i, vr, vi)
        elif which == 'left':
            obj.getLeftEigenvector(i, vr, vi)
    elif isinstance(obj, PEP):
        eig = obj.getEigenpair(i, vr, vi)
        omega = eig
    if BlochRemapper:
        vr = matrix_vector(BlochRemapper, vr)
    if matrices and getattr(matrices.ops, "part", None) is None:
        V = matrices.V
    else:
        V = functionspace(mesh, ("CG", degree))        # global container (multi-GPU: replicated host vector)
    p = Function(V)
    FixSign(vr)
    p.x.petsc_vec.setArray(vr.array)
    meas = np.sqrt(_mass_form(V, matrices, p.x.array))
    temp = vr.array / meas
    if absolute:
        abs_temp = abs(temp)
        temp = abs_temp / np.amax(abs_temp)
    p_normalized = Function(V)
    p_normalized.x.petsc_vec.setArray(temp)
    if rank0() and print_eigs:
        print(f"Eigenvalue-> {omega:.6f} | Eigenfrequency-> {omega/(2*np.pi):.6f}\n ")
    return omega, p_normalized


class _VectorSpace:
    """Blocked (x, y, z interleaved) P1 vector space: what velocity_eigenvector's result lives in."""

    def __init__(self, mesh):
        self.mesh, self.degree, self.bs = mesh, 1, 3
        self.n = 3 * mesh.n_nodes
        self.be = mesh.be


def velocity_eigenvector(mesh, p, omega, rho, degree=1, normalize=True, absolute=False):
    """u = grad(p) / (i omega rho) interpolated into the P1 vector space (eigenvectors.py:65-123).
    Post-processing next to the hot path, done on the host: the gradient of a P1 field is constant per
    cell and DOLFINx's interpolation leaves every vertex with the value of the last cell that holds it;
    normalised so that int u . conj(u) dx = 1.  P1 pressure fields only."""
    if degree != 1:
        raise NotImplementedError("velocity_eigenvector: degree 1 only")
    x, cells = mesh.x, mesh.cells.astype(np.int64)
    pv = np.asarray(p.x.array)[:mesh.n_nodes]
    e = x[cells[:, 1:]] - x[cells[:, :1]]                              # (nc, 3, 3) edge vectors
    dp = pv[cells[:, 1:]] - pv[cells[:, :1]]                           # (nc, 3)
    grad = np.linalg.solve(e.astype(complex), dp[:, :, None])[:, :, 0]  # e @ grad = dp
    last = np.zeros(mesh.n_nodes, np.int64)
    np.maximum.at(last, cells.ravel(), np.repeat(np.arange(len(cells)), 4))
    u = grad[last]
    if isinstance(rho, Function):
        u = u / np.asarray(rho.x.array)[:mesh.n_nodes, None]
        u = u / (1j * omega)
    else:
        u = u / (1j * omega * rho)
    if normalize:
        vol = np.abs(np.linalg.det(e)) / 6.0
        uc = u[cells]                                                   # (nc, 4, 3)
        integrand = (np.abs(uc) ** 2).sum(axis=(1, 2)) + (np.abs(uc.sum(axis=1)) ** 2).sum(axis=1)
        u = u / np.sqrt(float((vol / 20.0 * integrand).sum()))
    if absolute:
        mag = np.sqrt((u ** 2).sum(axis=1))
        u = np.abs(u) / np.abs(mag).max()
    out = Function(_VectorSpace(mesh), u.reshape(-1), name="U")
    return out


def normalize_adjoint(omega_dir, p_dir, p_adj, matrices, D=None):
    """p_adj <- p_adj / (p_adj . dL/domega p_dir) with the reference's dot convention
    (eigenvectors.py:125-177)."""
    info("- Normalizing the adjoint eigenvector to calculate shape derivatives..")
    B = matrices.B
    p_dir_vec = p_dir.x.petsc_vec
    p_adj_vec = p_adj.x.petsc_vec
    if not B and not D:
        dL_domega = matrices.C * (2 * omega_dir)
    elif B and not D:
        dL_domega = (B + matrices.C * (2 * omega_dir))
    elif D and not B:
        dL_domega = (matrices.C * (2 * omega_dir) - D.get_derivative(omega_dir))
    else:
        dL_domega = (B + matrices.C * (2 * omega_dir) - D.get_derivative(omega_dir))
    meas = vector_matrix_vector(p_adj_vec, dL_domega, p_dir_vec)
    p_adj_vec = multiply(p_adj_vec, 1 / meas)
    p_adj1 = p_adj
    p_adj1.name = "p_adj"
    p_adj1.x.petsc_vec.setArray(p_adj_vec.getArray())
    integral = vector_matrix_vector(p_adj1.x.petsc_vec, dL_domega, p_dir_vec)
    if rank0():
        print("! Normalization Check: ", integral)
    return p_adj1
