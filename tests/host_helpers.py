"""Host-logic test rig: the product's OperatorSet / Mat / eigensolvers running on the CPU
test double (oracle/host_backend.py) with oracle-assembled values."""
import numpy as np
import torch

from oracle import hx_oracle as ox
from oracle.host_backend import HostBackend
from helmholtz_x_b200.backend import CsrMatrix
from helmholtz_x_b200.operators import LowRankMat, Mat, OperatorSet, build_lowrank
from tests import cases


class HostSpace:
    """Quacks like fem.FunctionSpace for OperatorSet/AMG (pattern, matrix, dof_coords)."""

    def __init__(self, space, be):
        self.be, self.n, self.degree = be, space.n, space.degree
        ip, ix = ox.csr_pattern(space)
        self._pattern = (torch.from_numpy(ip.copy()), torch.from_numpy(ix.copy()))
        self.dof_coords = torch.from_numpy(np.ascontiguousarray(space.dof_x))
        self.oracle_space = space
        self.mesh = space.mesh
        self.facet_dofs = torch.from_numpy(np.ascontiguousarray(space.facet_dofs))

    def pattern(self):
        return self._pattern

    def matrix(self, values):
        return CsrMatrix(self.n, self.n, self._pattern[0], self._pattern[1], values)


class HostOperators:
    """Stand-in for AcousticMatrices built from oracle matrices."""

    def __init__(self, case, passive=False):
        self.oracle = cases.oracle_operators(case, passive)
        be = HostBackend()
        sp_ = self.oracle.space
        self.V = HostSpace(sp_, be)
        a = torch.from_numpy(ox._on_pattern(sp_, self.oracle.A).real.copy())
        c = torch.from_numpy(ox._on_pattern(sp_, self.oracle.C).real.copy())
        b = torch.from_numpy(ox._on_pattern(sp_, self.oracle.B)) if self.oracle.B is not None else None
        self.ops = OperatorSet(self.V, a, c, b)
        self.mesh = case.mesh
        self.A = Mat(self.ops, {"A": 1.0})
        self.C = Mat(self.ops, {"C": 1.0})
        self.B = Mat(self.ops, {"B": 1.0}) if b is not None else None
        self.B_adj = Mat(self.ops, {"Bh": 1.0}) if b is not None else None
        self.C_nobc_values = c
        self.V._mass = self.V.matrix(c.to(torch.complex128))      # eigenvectors._mass_form (no Dirichlet rows in these cases)


class HostFlame:
    """FlameMatrix stand-in from the oracle's thresholded vectors."""

    def __init__(self, case, hops):
        self.flame = cases.oracle_flame(case)
        self.FTF = self.flame.FTF
        be, n = hops.ops.be, hops.ops.n

        def lists(Mx):
            return [(np.flatnonzero(Mx[:, f]).astype(np.int32), Mx[np.flatnonzero(Mx[:, f]), f]) for f in range(Mx.shape[1])]
        L, R = lists(self.flame.left), lists(self.flame.right)
        lr, lrT = build_lowrank(be, n, L, R), build_lowrank(be, n, R, L)
        self._D_ij = LowRankMat(n, lr, lrT, 1.0, (L, R))
        self._D_ij_adj = LowRankMat(n, lrT, lr, 1.0, (R, L))
        self.matrix = self.adjoint_matrix = None

    def blochify(self, bloch_object):
        self._D_ij = bloch_object.blochify(self._D_ij)

    def assemble_matrix(self, omega, problem_type='direct'):
        if problem_type == 'direct':
            self.matrix = self._D_ij * self.FTF(omega)
        else:
            self.adjoint_matrix = self._D_ij_adj * np.conj(self.FTF(np.conj(omega)))

    def get_derivative(self, omega):
        return self.FTF.derivative(omega) * self._D_ij
