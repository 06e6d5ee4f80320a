"""Drop-in for helmholtz_x/bloch_operator.py: Bloch reduction of one sector of an N-periodic
geometry (BASELINE config 4).

The reference builds two SciPy maps BN (n x n_red) and NB = BN^H and forms NB*M*BN on the
host for every matrix (bloch_operator.py:42-78,104-111).  Here the maps are never formed: the
reduced operator is a relabelling of the assembled CSR entries,

    M_b[red(i), red(j)] += conj(phase_i) * M[i, j] * phase_j ,

red(i) = reduced index of dof i (a master dof takes its slave's), phase_i = f_b =
exp(2 pi i b / N) on master dofs and 1 elsewhere.  A, B and C share one reduced pattern, so
the shifted operator P(sigma) = A_b + sigma B_b + sigma^2 C_b is still one `combine_abc` and
the multigrid hierarchy is built once.  A_b and C_b are Hermitian, not real.

`pairing`: the reference pairs the k-th master dof with the k-th slave dof in ascending dof
index (bloch_operator.py:33-40), which is only a valid periodic identification when DOLFINx
happens to number both faces alike.  pairing="geometric" (default) pairs a
master dof with the slave dof it maps onto under the rotation by 2 pi / N about z;
pairing="sorted" restates the reference and takes `numbering` (reference dof index of every
dof here) to reproduce its golden logs.
"""
import numpy as np
import torch

from .backend import CsrMatrix
from .fem import _Vec
from .operators import LowRankMat, Mat, OperatorSet, build_lowrank

c128 = torch.complex128
i32 = torch.int32


class ReducedSpace:
    """What OperatorSet needs from a function space, on the reduced (master-free) dof set."""

    def __init__(self, be, n, indptr, indices, dof_coords):
        self.be, self.n = be, n
        self._pattern = (indptr, indices)
        self.dof_coords = dof_coords

    def pattern(self):
        return self._pattern

    def matrix(self, values):
        return CsrMatrix(self.n, self.n, self._pattern[0], self._pattern[1], values)


class BlochRemapper:
    """BN: reduced vector -> full-sector vector, full[i] = phase_i * reduced[red(i)]
    (what normalize_eigenvector(..., BlochRemapper=bloch.remapper) multiplies with,
    eigenvectors.py:35-36)."""

    def __init__(self, be, red, phase, n_red):
        self.be, self.red, self.phase, self.n_red = be, red, phase, n_red

    def getSize(self):
        return (int(self.red.numel()), self.n_red)

    def createVecs(self):
        return _Vec(np.zeros(self.n_red, complex)), _Vec(np.zeros(int(self.red.numel()), complex))

    def apply(self, x, y):
        torch.mul(self.phase, x[self.red], out=y)
        return y

    def mult(self, x, y):
        xd = self.be.asarray(np.asarray(x.array, complex), dtype=c128)
        y.setArray((self.phase * xd[self.red]).cpu().numpy())

    def __bool__(self):
        return True


def _face_dofs(V, tag):
    sel = np.flatnonzero(V.mesh.facet_tags == tag)
    return np.unique(V.facet_dofs.cpu().numpy()[sel].ravel()).astype(np.int64)


def _geometric_pairs(V, md, sd, N, tol):
    """slave partner of every master dof: nearest slave dof to the rotated master point."""
    dev = V.dof_coords.device
    X = V.dof_coords
    Xm = X[torch.as_tensor(md, device=dev)]
    Xs = X[torch.as_tensor(sd, device=dev)]
    scale = float(X.abs().max())
    best = None
    for sgn in (1.0, -1.0):
        a = sgn * 2 * np.pi / N
        Rm = torch.tensor([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=X.dtype, device=dev)
        Y = Xm @ Rm.T
        idx = torch.empty(len(md), dtype=torch.int64, device=dev)
        dmax = 0.0
        for s in range(0, len(md), 2048):
            d = torch.cdist(Y[s:s + 2048], Xs)
            dm, j = d.min(dim=1)
            idx[s:s + 2048] = j
            dmax = max(dmax, float(dm.max()))
        if best is None or dmax < best[0]:
            best = (dmax, idx)
    if best[0] > tol * max(scale, 1.0):
        raise ValueError(f"master and slave faces are not images under the rotation by 2 pi/{N} "
                         f"(largest mismatch {best[0]:.3e})")
    j = best[1].cpu().numpy()
    if len(np.unique(j)) != len(j):
        raise ValueError("master/slave pairing is not one-to-one")
    return sd[j]


class Blochifier:
    def __init__(self, geometry, boundary_conditions, N, passive_matrices, active_matrix=None,
                 pairing="geometric", numbering=None, tol=1e-6):
        self.passive_matrices = passive_matrices
        self.active_matrix = active_matrix
        self.b = 1.0
        self.periodicity = N
        self.f_b = np.exp(self.b * 1j * 2 * np.pi / N)
        self.mesh = getattr(geometry, "mesh", geometry)
        self.facet_tags = getattr(geometry, "facet_tags", None)
        self.V = V = passive_matrices.V
        full = passive_matrices.ops
        if full.part is not None:
            raise NotImplementedError("the Bloch reduction runs on one GPU (the sector is 1/N of the problem)")
        be = self.be = full.be
        self._A = self._B = self._B_adj = self._C = self._D = None

        vals = list(boundary_conditions.values())
        keys = list(boundary_conditions.keys())
        master_tag = keys[vals.index("Master")]
        slave_tag = keys[vals.index("Slave")]
        md = _face_dofs(V, master_tag)
        sd = _face_dofs(V, slave_tag)
        assert len(md) == len(sd)
        if pairing == "sorted":
            if numbering is not None:
                numbering = np.asarray(numbering)
                md = md[np.argsort(numbering[md], kind="stable")]
                sd = sd[np.argsort(numbering[sd], kind="stable")]
        elif pairing == "geometric":
            sd = _geometric_pairs(V, md, sd, N, tol)
        else:
            raise ValueError("pairing must be 'geometric' or 'sorted'")
        self.dofs_master, self.dofs_slave = md, sd

        n = V.n
        self.N = n                                   # the reference overwrites N with the full size (bloch_operator.py:45)
        keep = np.ones(n, bool)
        keep[md] = False
        n_red = int(keep.sum())
        red = -np.ones(n, np.int64)
        red[keep] = np.arange(n_red)
        if np.any(red[sd] < 0):
            raise ValueError("a dof lies on both the master and the slave face (the sector axis); not supported")
        red[md] = red[sd]
        phase = np.ones(n, complex)
        phase[md] = self.f_b
        self._red_h = red
        self._is_master = ~keep
        dev = V.dof_coords.device
        self.red = torch.as_tensor(red, device=dev)
        self.phase = torch.as_tensor(phase, device=dev)
        self.n_red = n_red

        # reduced pattern: unique (row, col) keys of the relabelled entries, row-major
        indptr, indices = V.pattern()
        nnz = int(indices.numel())
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (indptr[1:] - indptr[:-1]).long())
        cols = indices.long()
        key = self.red[rows] * n_red + self.red[cols]
        ukey, self._inv = torch.unique(key, return_inverse=True)
        self._w = torch.conj_physical(self.phase[rows]) * self.phase[cols]
        r_rows = torch.div(ukey, n_red, rounding_mode="floor")
        r_ptr = torch.zeros(n_red + 1, dtype=torch.int64, device=dev)
        r_ptr[1:] = torch.cumsum(torch.bincount(r_rows, minlength=n_red), 0)
        self._nnz_red = int(ukey.numel())
        coords = V.dof_coords[torch.as_tensor(np.flatnonzero(keep), device=dev)].contiguous()
        self.space = ReducedSpace(be, n_red, r_ptr.to(i32).contiguous(), (ukey - r_rows * n_red).to(i32).contiguous(), coords)
        del key, rows, cols, ukey, r_rows

        a_b = self._reduce(full.base["A"])
        c_b = self._reduce(full.base["C"])
        b_b = self._reduce(full.base["B"]) if full.base["B"] is not None else None
        self.ops = OperatorSet(self.space, a_b, c_b, b_b)
        self.ops.symmetric = False
        if b_b is not None:
            # (NB B BN)^H = NB conj(B) BN, not the entrywise conjugate
            self.ops.base["Bh"] = self._reduce(full.base["Bh"])
        self._A = Mat(self.ops, {"A": 1.0})
        self._C = Mat(self.ops, {"C": 1.0})
        if b_b is not None:
            self._B = Mat(self.ops, {"B": 1.0})
        self._BN = BlochRemapper(be, self.red, self.phase, n_red)

    def _reduce(self, values):
        out = torch.zeros(self._nnz_red, dtype=c128, device=values.device)
        out.index_add_(0, self._inv, values.to(c128) * self._w)
        return out

    @property
    def A(self):
        return self._A

    @property
    def B(self):
        return self._B

    @property
    def B_adj(self):
        return self._B_adj

    @property
    def C(self):
        return self._C

    @property
    def D(self):
        return self._D

    @property
    def remapper(self):
        return self._BN

    def blochify(self, matrix):
        """NB * matrix * BN (bloch_operator.py:104-111) for a combination of the passive operators
        or for the flame operator's vector pairs."""
        if isinstance(matrix, Mat):
            if matrix.ops is not self.passive_matrices.ops:
                raise TypeError("blochify: the matrix was not assembled on this Blochifier's AcousticMatrices")
            return Mat(self.ops, matrix.terms, [self.blochify(m) for m in matrix.lowrank])
        if isinstance(matrix, LowRankMat):
            lefts, rights = matrix.lists
            def remap(pairs):
                out = []
                for idx, val in pairs:
                    idx = np.asarray(idx, np.int64)
                    if np.any(self._is_master[idx]):
                        raise NotImplementedError("the flame / measurement support touches the master face; "
                                                  "its Bloch image would need complex-valued flame vectors")
                    out.append((self._red_h[idx].astype(np.int32), np.asarray(val)))
                return out
            l, r = remap(lefts), remap(rights)
            return LowRankMat(self.n_red, build_lowrank(self.be, self.n_red, l, r), build_lowrank(self.be, self.n_red, r, l),
                              matrix.coef, (l, r))
        raise TypeError("blochify expects a Mat or the flame operator's LowRankMat")
