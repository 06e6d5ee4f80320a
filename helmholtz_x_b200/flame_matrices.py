"""Drop-in for helmholtz_x/flame_matrices.py.  The flame operator is kept as the
FTF(omega)-scaled sum of sparse left/right vector pairs and applied matrix-free; the
dense block the reference builds (flame_matrices.py:91,233) is never formed."""
import numpy as np
import torch

from . import fem
from .operators import LowRankMat, build_lowrank
from .parameters_utils import gamma_function
from .solver_utils import info


class FlameMatrix:
    def __init__(self, mesh, h, q_0, u_b, FTF, degree, bloch_object=None, tol=1e-5):
        self.mesh = mesh
        self.h = h
        self.q_0 = q_0
        self.u_b = u_b
        self.FTF = FTF
        self.degree = degree
        self.bloch_object = bloch_object
        self.tol = tol
        self.V = fem.functionspace(mesh, ("Lagrange", degree))
        self.gdim = 3
        self.global_size = self.V.n
        self.local_size = self.V.n
        self._D_ij = None
        self._D_ij_adj = None
        self._D = None
        self._D_adj = None

    @property
    def matrix(self):
        return self._D

    @property
    def submatrices(self):
        return self._D_ij

    @property
    def adjoint_matrix(self):
        return self._D_adj

    @property
    def adjoint_submatrices(self):
        return self._D_ij_adj

    def indices_and_values(self, dense):
        """Threshold |v| < tol -> 0 and compact to (dof, value) (flame_matrices.py:61-73)."""
        fem.threshold(self.V.be, dense, self.tol)
        idx = torch.nonzero(dense).reshape(-1)
        return idx.cpu().numpy().astype(np.int32), dense[idx].cpu().numpy()

    def _set(self, lefts, rights, problem_type):
        be, n = self.V.be, self.V.n
        lr = build_lowrank(be, n, lefts, rights)
        lr_T = build_lowrank(be, n, rights, lefts)
        if problem_type == 'direct':
            self._D_ij = LowRankMat(n, lr, lr_T, 1.0, (lefts, rights))
        elif problem_type == 'adjoint':
            self._D_ij_adj = LowRankMat(n, lr_T, lr, 1.0, (rights, lefts))
        else:
            ValueError("The problem type should be specified as 'direct' or 'adjoint'.")

    def assemble_matrix(self, omega, problem_type='direct'):
        """D = FTF(omega) D_ij ; D_adj = conj(FTF(conj omega)) D_ij_adj (flame_matrices.py:96-108)."""
        if problem_type == 'direct':
            self._D = self._D_ij * self.FTF(omega)
            info("- Direct matrix D is assembling...")
        elif problem_type == 'adjoint':
            self._D_adj = self._D_ij_adj * np.conj(self.FTF(np.conj(omega)))
            info("- Adjoint matrix D is assembling...")
        else:
            ValueError("The problem type should be specified as 'direct' or 'adjoint'.")
        info("- Matrix D is assembled.")

    def get_derivative(self, omega):
        dD_domega = self.FTF.derivative(omega) * self._D_ij
        info("- Derivative of matrix D is assembled.")
        return dD_domega


class PointwiseFlameMatrix(FlameMatrix):

    def __init__(self, mesh, subdomains, x_r, h, rho_u, q_0, u_b, FTF, degree=1, bloch_object=None, gamma=1.4, tol=1e-10):
        super().__init__(mesh, h, q_0, u_b, FTF, degree, bloch_object, tol)
        self.x_r = x_r
        self.rho_u = rho_u
        self.gamma = gamma
        self.subdomains = subdomains

    def _assemble_vectors(self, flame, point=None):
        left = fem.flame_left(self.V, self.h, self.q_0 / self.u_b, gm1_const=self.gamma - 1, tag=flame)
        return self.indices_and_values(left)

    def assemble_submatrices(self, problem_type='direct'):
        info("- Generating matrix D..")
        V = self.V
        pts = np.asarray(self.x_r, float).reshape(-1, 3)
        owner, ptsd = fem.locate_points(self.mesh, pts, 1e-10)
        dz = fem.point_dphidz(V, ptsd, owner).cpu().numpy()
        owner_h = owner.cpu().numpy()
        cell_dofs = V.cell_dofs[owner.clamp_max(self.mesh.n_cells - 1).long()].cpu().numpy()
        lefts, rights = [], []
        for flame in range(len(pts)):
            lefts.append(self._assemble_vectors(flame))
            if owner_h[flame] >= self.mesh.n_cells:
                rights.append((np.zeros(0, np.int32), np.zeros(0)))
            else:
                vals = dz[flame] / self.rho_u
                vals = np.where(np.abs(vals) < self.tol, 0.0, vals)
                keep = vals != 0.0
                rights.append((cell_dofs[flame][keep].astype(np.int32), vals[keep]))
            info("- Matrix contribution of flame " + str(flame) + " is computed.")
        self._set(lefts, rights, problem_type)
        info("- Submatrix D is Assembled.")


class DistributedFlameMatrix(FlameMatrix):

    def __init__(self, mesh, w, h, rho, T, q_0, u_b, FTF, degree=1, bloch_object=None, gamma=None, tol=1e-5):
        super().__init__(mesh, h, q_0, u_b, FTF, degree, bloch_object, tol)
        if gamma is None:
            gamma = gamma_function(T)
        self.gamma = gamma
        self.w, self.rho = w, rho

    def _assemble_vectors(self, problem_type='direct'):
        if np.ndim(self.gamma) == 0 and not isinstance(self.gamma, fem.Function):
            left = fem.flame_left(self.V, self.h, self.q_0 / self.u_b, gm1_const=float(self.gamma) - 1.0)
        else:
            g = self.gamma.x.array.real if isinstance(self.gamma, fem.Function) else np.asarray(self.gamma)
            left = fem.flame_left(self.V, self.h, self.q_0 / self.u_b, gm1_nodal=g - 1.0)
        right = fem.flame_right(self.V, self.w, self.rho)
        return self.indices_and_values(left), self.indices_and_values(right)

    def assemble_submatrices(self, problem_type='direct'):
        left, right = self._assemble_vectors(problem_type)
        info("- Generating matrix D..")
        self._set([left], [right], problem_type)
        info("- Submatrix D is Assembled.")
