"""Drop-in for helmholtz_x/flame_matrices.py.  The flame operator is kept as the
FTF(omega)-scaled sum of sparse left/right vector pairs and applied matrix-free; the
dense block the reference builds (flame_matrices.py:91,233) is never formed."""
import numpy as np
import torch

from . import fem
from .operators import LowRankMat, build_lowrank
from .parameters_utils import gamma_function
from .phases import phase
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
        # multi-GPU: this rank works on its sub-mesh and keeps the entries of the rows it owns
        self.part = mesh.partition()
        self.dpart = mesh.dof_partition(degree)           # rows of the degree-`degree` space over the ranks
        self.amesh = mesh if self.part is None else self.part.local_mesh
        self.V = fem.functionspace(self.amesh, ("Lagrange", degree))
        self.gdim = 3
        self.global_size = fem.functionspace(mesh, ("Lagrange", degree)).n if self.part is None else self.dpart.n_global
        self.local_size = self.V.n if self.part is None else self.dpart.n_own
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
        if self.part is not None:                        # ghost rows are incomplete and belong to other ranks
            dense = self._own(dense)
        idx = torch.nonzero(dense).reshape(-1)
        return idx.cpu().numpy().astype(np.int32), dense[idx].cpu().numpy()

    def _own(self, v):
        """Entries of the rows this rank owns, in the distributed layout, from a vector of the local space."""
        n_own = self.dpart.n_own
        if self.degree == 1:
            return v[:n_own]
        return v[self.dpart.dev("dof_perm").to(v.device)[:n_own]].contiguous()

    def _local(self, f):
        """Coefficient restricted to this rank's sub-mesh (identity on one GPU)."""
        if self.part is None or not isinstance(f, fem.Function):
            return f
        vals = f.real_device()                            # restricted on the device, no host round trip
        if isinstance(f.function_space, fem.DG0Space):
            return fem.Function.from_device(fem.DG0Space(self.amesh), self.part.restrict_cell(vals))
        return fem.Function.from_device(fem.functionspace(self.amesh, ("Lagrange", 1)),
                                        self.part.restrict_nodal(vals[:self.mesh.n_nodes]))

    def _set(self, lefts, rights, problem_type):
        be, n = self.V.be, self.local_size
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

    def blochify(self, problem_type='direct'):
        """Reduce the vector pairs with the Blochifier handed to the constructor (flame_matrices.py:117-127)."""
        if problem_type == 'direct':
            self._D_ij = self.bloch_object.blochify(self.submatrices)
        elif problem_type == 'adjoint':
            self._D_ij_adj = self.bloch_object.blochify(self.adjoint_submatrices)
        else:
            ValueError("The problem type should be specified as 'direct' or 'adjoint'.")

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
        if getattr(self, "_h_dev", None) is None:
            # one upload of the heat-release field for all flames (a DG0 field has one value per cell)
            self._h_dev = fem._field(self.V, self._local(self.h))[0]
        left = fem.flame_left(self.V, self._h_dev, self.q_0 / self.u_b, gm1_const=self.gamma - 1, tag=flame)
        return self.indices_and_values(left)

    def assemble_submatrices(self, problem_type='direct'):
        with phase("flame_vectors"):
            self._assemble_submatrices(problem_type)

    def _assemble_submatrices(self, problem_type):
        info("- Generating matrix D..")
        V = self.V
        pts = np.asarray(self.x_r, float).reshape(-1, 3)
        amesh = self.amesh
        owner, ptsd = fem.locate_points(amesh, pts, 1e-10)
        owner_h = owner.cpu().numpy().astype(np.int64)
        if self.part is not None:
            # lowest GLOBAL cell index wins (same tie-break as one GPU); every rank that holds
            # that cell evaluates it and keeps the entries of the rows it owns
            import torch.distributed as tdist
            none = np.iinfo(np.int64).max
            gid = np.where(owner_h < amesh.n_cells, self.part.cell_ids[np.minimum(owner_h, amesh.n_cells - 1)], none)
            gmin = torch.as_tensor(gid, device=owner.device)
            tdist.all_reduce(gmin, op=tdist.ReduceOp.MIN)
            gmin = gmin.cpu().numpy()
            pos = np.searchsorted(self.part.cell_ids, gmin)
            have = (pos < amesh.n_cells) & (self.part.cell_ids[np.minimum(pos, amesh.n_cells - 1)] == gmin)
            owner_h = np.where(have, pos, amesh.n_cells)
            owner = V.be.asarray(owner_h.astype(np.int32), dtype=torch.int32)
        dz = fem.point_dphidz(V, ptsd, owner.clamp_max(amesh.n_cells - 1)).cpu().numpy()
        cell_dofs = V.cell_dofs[owner.clamp_max(amesh.n_cells - 1).long()].cpu().numpy()
        lefts, rights = [], []
        for flame in range(len(pts)):
            lefts.append(self._assemble_vectors(flame))
            if owner_h[flame] >= amesh.n_cells:
                rights.append((np.zeros(0, np.int32), np.zeros(0)))
            else:
                vals = dz[flame] / self.rho_u
                vals = np.where(np.abs(vals) < self.tol, 0.0, vals)
                keep = vals != 0.0
                dofs = cell_dofs[flame]
                if self.part is not None:
                    if self.degree != 1:
                        dofs = self.dpart.dof_inv_perm[dofs]         # local space numbering -> [owned | ghosts]
                    keep &= dofs < self.dpart.n_own
                rights.append((dofs[keep].astype(np.int32), vals[keep]))
            info("- Matrix contribution of flame " + str(flame) + " is computed.")
        self._h_dev = None
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
            left = fem.flame_left(self.V, self._local(self.h), self.q_0 / self.u_b, gm1_const=float(self.gamma) - 1.0)
        else:
            g = self.gamma.x.array.real if isinstance(self.gamma, fem.Function) else np.asarray(self.gamma)
            if self.part is not None:
                g = self.part.restrict_nodal(g[:self.mesh.n_nodes])
            left = fem.flame_left(self.V, self._local(self.h), self.q_0 / self.u_b, gm1_nodal=g - 1.0)
        right = fem.flame_right(self.V, self._local(self.w), self._local(self.rho))
        return self.indices_and_values(left), self.indices_and_values(right)

    def assemble_submatrices(self, problem_type='direct'):
        with phase("flame_vectors"):
            left, right = self._assemble_vectors(problem_type)
        info("- Generating matrix D..")
        self._set([left], [right], problem_type)
        info("- Submatrix D is Assembled.")
