"""Matrix objects handed across the helmholtz_x API and the shifted inner solve.

The reference passes PETSc ``Mat`` handles around and builds new matrices with
``A + omega*B - D`` (helmholtz_x/eigensolvers.py:173-176,240,309-315).  Here a ``Mat`` is
a *symbolic* linear combination of the assembled base operators A, B, B^H, C that
share one CSR pattern, plus optional rank-r flame terms kept as sparse vectors
(never densified): the same expressions work, nothing is re-assembled, and the
shift-invert solve sees the structure it needs (multigrid on the sparse part,
Woodbury for the low-rank part).
"""
from __future__ import annotations

import numpy as np
import torch

from . import krylov
from .amg import AMG
from .backend import LowRank
from .fem import _Vec
from .phases import phase

c128 = torch.complex128
f64 = torch.float64

BASES = ("A", "B", "Bh", "C")
# Gram-Schmidt passes per inner GMRES step.  One pass (PETSc's GMRES default, classical Gram-Schmidt
# without refinement) halves the basis traffic; the multigrid-preconditioned operator keeps
# ||w_after|| / ||w_before|| around 0.35 per step, iteration counts and true residuals are the same as
# with two passes (tests/test_host_logic.py), and every solve ends on a recomputed true residual.
# The outer Krylov-Schur basis always uses two passes.  HX_GMRES_ORTH=cgs2 restores two passes.
import os as _os
GMRES_ORTH_PASSES = 2 if _os.environ.get("HX_GMRES_ORTH", "cgs1") == "cgs2" else 1
# Normwise backward error ||b - P x|| / (||P|| ||x||) at which an inner solve is as good as the exact
# LU of the reference (PETSc LU / MUMPS behind eigensolvers.py:49-50): accepted outright at BACKWARD_ETA,
# accepted at the stagnation floor up to BACKWARD_ETA_FLOOR.  The relative residual ||r|| / ||b|| cannot
# reach rtol when the shift sits on an eigenvalue (||x|| ~ ||b|| / |lambda_min|); the backward error can.
BACKWARD_ETA = 1e-14
BACKWARD_ETA_FLOOR = 1e-12


def build_lowrank(be, n, left_list, right_list) -> LowRank:
    """left_list / right_list: per flame (idx int array, val float array), host numpy."""
    r = len(left_list)
    rptr = np.zeros(r + 1, np.int32)
    for f, (idx, _) in enumerate(right_list):
        rptr[f + 1] = rptr[f] + len(idx)
    ridx = np.concatenate([np.asarray(i, np.int32) for i, _ in right_list]) if r else np.zeros(0, np.int32)
    rval = np.concatenate([np.asarray(v, np.float64) for _, v in right_list]) if r else np.zeros(0)
    rows = np.concatenate([np.asarray(i, np.int64) for i, _ in left_list]) if r else np.zeros(0, np.int64)
    cols = np.concatenate([np.full(len(i), f, np.int32) for f, (i, _) in enumerate(left_list)]) if r else np.zeros(0, np.int32)
    vals = np.concatenate([np.asarray(v, np.float64) for _, v in left_list]) if r else np.zeros(0)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    lrow, start = np.unique(rows, return_index=True)
    lptr = np.append(start, len(rows)).astype(np.int32)
    a = be.asarray
    return LowRank(n, r, a(rptr, dtype=torch.int32), a(ridx, dtype=torch.int32), a(rval, dtype=f64),
                   a(lrow.astype(np.int32), dtype=torch.int32), a(lptr, dtype=torch.int32), a(cols, dtype=torch.int32),
                   a(vals, dtype=f64))


class LowRankMat:
    """coef * sum_f left_f right_f^T  (what FlameMatrix.matrix returns)."""

    def __init__(self, n, lr: LowRank, lr_T: LowRank, coef=1.0, lists=None):
        self.n, self.lr, self.lr_T, self.coef, self.lists = n, lr, lr_T, complex(coef), lists

    def __mul__(self, z):
        return LowRankMat(self.n, self.lr, self.lr_T, self.coef * complex(z), self.lists)

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    def transpose(self):
        return LowRankMat(self.n, self.lr_T, self.lr, self.coef, None if self.lists is None else self.lists[::-1])

    def getSize(self):
        return (self.n, self.n)

    def dense_block_nnz(self):
        l, r = self.lists
        return int(sum(len(a[0]) * len(b[0]) for a, b in zip(l, r)))

    def to_scipy(self):
        """Densified D_ij as SciPy CSR (tests only; the product never forms it)."""
        import scipy.sparse as sp
        l, r = self.lists
        M = sp.csr_matrix((self.n, self.n), dtype=complex)
        for (li, lv), (ri, rv) in zip(l, r):
            M = M + sp.csr_matrix((np.outer(lv, rv).ravel(), (np.repeat(li, len(ri)), np.tile(ri, len(li)))),
                                  shape=(self.n, self.n))
        return self.coef * M


class Mat:
    """sum_k terms[k] * base_k  +  sum (LowRankMat)   on one OperatorSet."""

    def __init__(self, ops, terms, lowrank=()):
        self.ops = ops
        self.terms = {k: complex(v) for k, v in terms.items() if v != 0}
        self.lowrank = tuple(lowrank)

    # -- algebra ---------------------------------------------------------------------------
    def __mul__(self, z):
        z = complex(z)
        return Mat(self.ops, {k: v * z for k, v in self.terms.items()}, [m * z for m in self.lowrank])

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    def __add__(self, other):
        if isinstance(other, LowRankMat):
            return Mat(self.ops, self.terms, self.lowrank + (other,))
        if other is None or (np.isscalar(other) and other == 0):
            return self
        if not isinstance(other, Mat) or other.ops is not self.ops:
            raise TypeError("Mat + Mat needs operands assembled on the same AcousticMatrices")
        t = dict(self.terms)
        for k, v in other.terms.items():
            t[k] = t.get(k, 0) + v
        return Mat(self.ops, t, self.lowrank + other.lowrank)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-other)

    def __bool__(self):
        return True

    def copy(self):
        return Mat(self.ops, self.terms, self.lowrank)

    def hermitian_transpose(self):
        """conj-transpose: A, C real symmetric; B complex symmetric => B^H = conj(B)."""
        sw = {"A": "A", "C": "C", "B": "Bh", "Bh": "B"}
        return Mat(self.ops, {sw[k]: np.conj(v) for k, v in self.terms.items()},
                   [LowRankMat(m.n, m.lr_T, m.lr, np.conj(m.coef), None if m.lists is None else m.lists[::-1])
                    for m in self.lowrank])

    # -- PETSc-like surface used by drivers / eigenvectors.py --------------------------------------
    def getSize(self):
        return (self.ops.n_global, self.ops.n_global)

    def createVecs(self):
        return _Vec(np.zeros(self.ops.n_global, complex)), _Vec(np.zeros(self.ops.n_global, complex))

    getVecs = createVecs

    def values(self):
        return self.ops.combine(self.terms)

    def csr(self):
        return self.ops.space.matrix(self.values())

    def getValuesCSR(self):
        indptr, indices = self.ops.space.pattern()
        return indptr.cpu().numpy(), indices.cpu().numpy(), self.values().cpu().numpy()

    def to_scipy(self):
        import scipy.sparse as sp
        ip, ix, v = self.getValuesCSR()
        M = sp.csr_matrix((v, ix, ip), shape=self.getSize())
        for lr in self.lowrank:
            M = M + lr.to_scipy()
        return M

    def apply(self, x, y):
        """y = self @ x on device tensors (sparse part by SpMV, flame part matrix-free)."""
        be = self.ops.be
        be.spmv(self.csr(), x, y)
        for m in self.lowrank:
            t = be.zeros(max(m.lr.r, 1))
            be.lowrank_dots(m.lr, x, t)
            be.lowrank_update(m.lr, t, m.coef, y)
        return y

    def mult(self, x, y):
        """PETSc MatMult on host Vec wrappers (petsc4py_utils.py:86,96)."""
        be = self.ops.be
        xd = be.asarray(self.ops.to_local(np.asarray(x.array, complex)), dtype=c128)
        yd = be.zeros(self.ops.n)
        self.apply(xd, yd)
        y.setArray(self.ops.to_global(yd))


class OperatorSet:
    """Base operators of one AcousticMatrices on the device + the multigrid hierarchy."""

    def __init__(self, space, a_vals, c_vals, b_vals=None):
        self.space, self.be, self.n = space, space.be, space.n
        self.part = getattr(space, "part", None)                   # multi-GPU row partition (dist.DistSpace)
        self.n_global = getattr(space, "n_global", space.n)
        self.base = {"A": a_vals, "C": c_vals, "B": b_vals, "Bh": torch.conj_physical(b_vals) if b_vals is not None else None}
        self._amg = None
        self._cache = {}
        self.amg_options = {}
        self.stats = {"inner_solves": 0, "inner_iterations": 0, "amg_setups": 0, "shifts": 0,
                      "t_amg_setup": 0.0, "t_shift": 0.0, "t_inner": 0.0}

    def combine(self, terms):
        key = tuple(sorted(terms.items()))
        if key not in self._cache:
            if len(self._cache) > 8:
                self._cache.clear()
            out = self.be.empty(self.space.pattern()[1].numel())
            t = dict(terms)
            if t.get("B", 0) != 0 and t.get("Bh", 0) != 0:
                raise ValueError("B and B^H cannot be combined in one operator")
            b, cb = (self.base["Bh"], t.get("Bh")) if t.get("Bh", 0) != 0 else (self.base["B"], t.get("B", 0))
            if cb and b is None:
                raise ValueError("operator has no B matrix")
            self.be.combine_abc(self.base["A"], b if cb else None, self.base["C"], t.get("A", 0), cb or 0, t.get("C", 0), out)
            self._cache[key] = out
        return self._cache[key]

    def to_local(self, global_host_array):
        """Owned slice of a global host vector (identity on one GPU)."""
        if self.part is None:
            return global_host_array
        return np.ascontiguousarray(global_host_array[self.part.l2g[:self.part.n_own]])

    def to_global(self, local_device_vector):
        """Global host vector from the owned device pieces (all ranks get it)."""
        if self.part is None:
            return local_device_vector.cpu().numpy()
        return self.part.gather_global(local_device_vector)

    def coarse(self):
        """Global coarse correction of the multi-GPU preconditioner (dist.CoarseCorrection)."""
        if getattr(self, "_coarse", None) is None:
            from .dist import CoarseCorrection
            import os
            # measured on the 35 k-DoF annulus, 2 ranks (GMRES iterations per solve; one rank: 46):
            # nc 1000 -> 68, 2000 -> 60; a second (post) coarse correction changes nothing.  The dense
            # inverse kernel holds nc <= 1400.
            self._coarse = CoarseCorrection(self.space, self.base, nc=min(int(os.environ.get("HX_DIST_NC", "1000")), 1400))
        return self._coarse

    def hierarchy(self):
        """Row-distributed multigrid cycle (dist.DistHierarchy): the multi-GPU default -- the cycle is the
        single-GPU cycle up to summation order, so the iteration count does not depend on the number of
        ranks.  HX_DIST_HIERARCHY=0: two-level Schwarz (global coarse correction + rank-local cycles)."""
        import os
        if self.part is None or os.environ.get("HX_DIST_HIERARCHY", "1") != "1":
            return None
        if getattr(self, "_hier", None) is None:
            import time
            from .dist import DistHierarchy
            t0 = time.perf_counter()
            with phase("amg_setup"):
                self._hier = DistHierarchy(self.space, self.base, **self.amg_options)
            self._amg = self._hier.mg
            self.stats["amg_setups"] += 1
            self.stats["t_amg_setup"] += time.perf_counter() - t0
        return self._hier

    def amg(self):
        if self._amg is None and self.hierarchy() is not None:
            return self._amg
        if self._amg is None and self.part is not None:
            # block-Jacobi across GPUs: each rank's hierarchy lives on its diagonal block
            import time
            t0 = time.perf_counter()
            dm = self.space.diag_matrix
            B = dm(self.base["B"]) if self.base["B"] is not None else None
            with phase("amg_setup"):
                self._amg = AMG(self.space.local_be, dm(self.base["A"]), dm(self.base["C"]), B, self.space.dof_coords,
                                **self.amg_options)
            self.stats["amg_setups"] += 1
            self.stats["t_amg_setup"] += time.perf_counter() - t0
        if self._amg is None:
            pat = self.space.matrix
            B = pat(self.base["B"]) if self.base["B"] is not None else None
            import time
            t0 = time.perf_counter()
            with phase("amg_setup"):
                self._amg = AMG(self.be, pat(self.base["A"]), pat(self.base["C"]), B, self.space.dof_coords, **self.amg_options)
            self.be.synchronize()
            self.stats["amg_setups"] += 1
            self.stats["t_amg_setup"] += time.perf_counter() - t0
        return self._amg


class ShiftedSolver:
    """x = (P + sum coef_k L_k R_k^T)^{-1} b,  P = sum terms*base  -- GMRES preconditioned
    by the SA-AMG V-cycle for P, Woodbury for the flame terms (hard part 2).

    The state that depends only on P (level operators, coarse inverse, and the solves
    Zbase = P^-1 L of the Woodbury update) is cached on the OperatorSet: in the fixed-point
    iteration the shift is the same for every iterate (only FTF(omega_k) changes), so the
    r extra solves are paid once per target instead of once per iterate."""

    def __init__(self, ops: OperatorSet, terms, lowrank=(), rtol=1e-11, restart=64, maxiter=512, transposed=False):
        import time
        hier = ops.hierarchy()
        if ops.part is not None and hier is None:
            restart, maxiter = 80, 800          # the two-level Schwarz preconditioner needs more iterations
        self.ops, self.be = ops, ops.be
        be = self.be
        self.rtol, self.restart, self.maxiter = rtol, restart, maxiter
        t = {k: complex(v) for k, v in terms.items() if v != 0}
        key = tuple(sorted(t.items()))
        mg = ops.amg()
        st = getattr(ops, "_shift_state", None)
        if st is None or st["key"] != key:
            with phase("shift_setup"):
                t_shift0 = time.perf_counter()
                use_bh = t.get("Bh", 0) != 0
                P_values = ops.combine(t)
                P = ops.space.matrix(P_values)
                if hier is not None:
                    fine_vals = None                 # replicated global levels are combined from their own bases
                else:
                    fine_vals = P_values if ops.part is None else P_values[ops.space.diag_sel].contiguous()
                if use_bh:
                    # coarse B^H = conj(B_c): run the hierarchy on conjugated B
                    for L in mg.levels:
                        if L.b is not None and not hasattr(L, "b_direct"):
                            L.b_direct = L.b
                            L.b_conj = torch.conj_physical(L.b)
                    for L in mg.levels:
                        if L.b is not None:
                            L.b = L.b_conj
                    mg.set_shift(t.get("A", 0), t.get("Bh", 0), t.get("C", 0), fine_values=fine_vals)
                    for L in mg.levels:
                        if L.b is not None:
                            L.b = L.b_direct
                else:
                    mg.set_shift(t.get("A", 0), t.get("B", 0), t.get("C", 0), fine_values=fine_vals)
                Pop = mg.fine_operator() if (len(mg.levels) > 1 and ops.part is None) else P
                if ops.part is not None and hasattr(ops.space, "matrix_sell"):
                    Pop = ops.space.matrix_sell(P_values) or P
                if hier is not None:
                    hier.set_fine(P_values)
                elif ops.part is not None:
                    ops.coarse().set_shift(t)
                # ||P|| estimate for the backward-error test: 2 max|p_ij| (<= 2 ||P||_2; FEM rows give
                # ||P||_inf of 2-3 max|p_ii|); the factor is immaterial against eta's decades
                pmax = P_values.abs().max() if P_values.numel() else torch.zeros((), dtype=f64)
                if ops.part is not None:
                    ops.part.all_reduce_max(pmax)
                st = {"key": key, "P_values": P_values, "P": P, "Pop": Pop, "wood": {}, "anorm": 2.0 * float(pmax),
                      "basis": krylov.ArnoldiBasis(be, ops.n, restart), "work": be.zeros(ops.n),
                      "xc": be.zeros(ops.n), "r": be.zeros(ops.n),
                      "zbasis": be.zeros(restart, ops.n) if getattr(mg, "single", False) else None}
                ops._shift_state = st
                ops.stats["shifts"] += 1
                ops.stats["t_shift"] += time.perf_counter() - t_shift0
        self.st = st
        self.mg = mg
        self.P_values, self.P, self.Pop = st["P_values"], st["P"], st["Pop"]
        self.basis, self.work = st["basis"], st["work"]
        self.lowrank = [m for m in lowrank if m.coef != 0 and m.lr.r > 0]
        if transposed and not getattr(ops, "symmetric", True):
            raise NotImplementedError("left eigenvectors need the transposed operator; the Bloch-reduced operators are "
                                      "Hermitian, not symmetric (the reference leaves Blochifier.B_adj unset as well)")
        if transposed:
            self.lowrank = [m.transpose() for m in self.lowrank]
        if self.lowrank:
            self._setup_woodbury()

    def _precond(self, v, out):
        """One GPU: the AMG V-cycle.  Multi-GPU: global coarse correction, then the rank-local
        AMG cycle on the updated residual (two-level multiplicative Schwarz)."""
        if self.ops.part is None:
            return self.mg.apply(v, out)
        if self.ops.hierarchy() is not None:
            return self.ops.hierarchy().apply(v, out)
        cc = self.ops.coarse()
        be = self.be
        xc, r = self.st["xc"], self.st["r"]
        cc.apply(v, xc)
        be.spmv(self.Pop, xc, r, alpha=-1.0, beta=1.0, y0=v)          # r = v - P xc (halo exchange inside)
        self.mg.apply(r, out)
        be.axpby(1.0, xc, 1.0, out)
        return out

    def _solve_P(self, b, x):
        with phase("inner_solve"):
            return self._solve_P_timed(b, x)

    def _solve_P_timed(self, b, x):
        import time
        t0 = time.perf_counter()
        info = {}

        def run(passes, x0=False):
            return krylov.gmres(self.be, lambda v, o: self.be.spmv(self.Pop, v, o), b, x, precond=self._precond,
                                rtol=self.rtol, restart=self.restart, maxiter=self.maxiter, basis=self.basis,
                                work=self.work, zbasis=self.st["zbasis"], orth_passes=passes,
                                anorm=self.st["anorm"], eta=BACKWARD_ETA, x0=x0, info=info)

        def acceptable():
            # converged to rtol, backward stable (next to an eigenvalue of P: see krylov.gmres), or at
            # the attainable floor with a backward error an exact factorisation would not beat by much
            return (info["status"] in ("converged", "backward_stable") or info["rel"] <= max(self.rtol * 100, 1e-8)
                    or (info["eta"] is not None and info["eta"] <= BACKWARD_ETA_FLOOR))
        its, rel = run(GMRES_ORTH_PASSES)
        stats = self.ops.stats
        if not acceptable():
            # iterative refinement: continue from the current x with re-orthogonalised Arnoldi steps
            stats["refinements"] = stats.get("refinements", 0) + 1
            its2, rel = run(2, x0=True)
            its += its2
        if info["status"] != "converged":
            stats["floor_accepts"] = stats.get("floor_accepts", 0) + 1
        stats["inner_solves"] += 1
        stats["inner_iterations"] += its
        stats["t_inner"] += time.perf_counter() - t0
        if not acceptable():
            raise RuntimeError(f"inner GMRES {info['status']}: rel. residual {rel:.2e}, backward error "
                               f"{info['eta']} after {its} iterations")
        return x

    def _setup_woodbury(self):
        """(P - U W^T)^-1 with U = -L diag(coef) (n x r), W = R:
        Zbase = P^-1 L (cached per shift), S = I - W^T Z = I + (W^T Zbase) diag(coef)."""
        be, n = self.be, self.ops.n
        wkey = tuple(id(m.lr) for m in self.lowrank)
        r_tot = sum(m.lr.r for m in self.lowrank)
        cache = self.st["wood"].get(wkey)
        if cache is None:
            with phase("woodbury"):
                Zbase = be.zeros(r_tot, n)
                u = be.zeros(n)
                col = 0
                for m in self.lowrank:
                    e = be.zeros(m.lr.r)
                    for f in range(m.lr.r):
                        u.zero_()
                        e.zero_()
                        e[f] = 1.0
                        be.lowrank_update(m.lr, e, 1.0, u)              # u = left_f
                        self._solve_P(u, Zbase[col])
                        col += 1
                Sbase = np.zeros((r_tot, r_tot), complex)
                row = 0
                t = be.zeros(max(max(m.lr.r for m in self.lowrank), 1))
                for m in self.lowrank:
                    for cz in range(r_tot):
                        be.lowrank_dots(m.lr, Zbase[cz], t)
                        Sbase[row:row + m.lr.r, cz] = t[:m.lr.r].cpu().numpy()
                    row += m.lr.r
                cache = (Zbase, Sbase, t)
                self.st["wood"] = {wkey: cache}                          # keep one entry: bounded memory
        self.Zbase, Sbase, self._t = cache
        self.coefs = np.concatenate([np.full(m.lr.r, m.coef) for m in self.lowrank])
        self.S_inv = np.linalg.inv(np.eye(r_tot, dtype=complex) + Sbase * self.coefs[None, :])
        self.r_tot = r_tot

    def solve(self, b, x):
        self._solve_P(b, x)
        if self.lowrank:
            be = self.be
            wt = np.zeros(self.r_tot, complex)
            row = 0
            for m in self.lowrank:
                be.lowrank_dots(m.lr, x, self._t)
                wt[row:row + m.lr.r] = self._t[:m.lr.r].cpu().numpy()
                row += m.lr.r
            cvec = self.S_inv @ wt
            # x += Z cvec with Z = -Zbase diag(coef)  <=>  x -= sum_j (coef_j cvec_j) Zbase_j
            be.multi_axpy(self.Zbase, self.r_tot, be.asarray(self.coefs * cvec, dtype=c128), x)
        return x
