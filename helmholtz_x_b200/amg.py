"""Smoothed-aggregation multigrid preconditioner for the shifted operator
P(sigma) = A + sigma B + sigma^2 C  -- the iterative stand-in for the exact LU
(PETSc LU / MUMPS) behind SLEPc's ST 'sinvert' (helmholtz_x/eigensolvers.py:49-50,
102-103).

Set-up (once per mesh; sigma-independent):
  * aggregates = runs of `agg_size` dofs along a Morton (Z-order) curve through the
    dof coordinates (embarrassingly parallel: one sort);
  * tentative prolongator T (piecewise constant, normalised), smoothed with one damped
    Jacobi step on the SPD surrogate K = -A + tau C:  P = (I - 4/(3 rho) D^-1 K) T;
  * Galerkin coarse operators A_c = P^T A P, B_c, C_c on a common coarse pattern.
  The sparse products of the set-up run through the library's own SpGEMM kernels (spgemm.py) on
  levels of >= 20 k rows; small / dense coarse levels and the CPU test double use torch.sparse.
Per shift (every outer fixed-point / Newton step): one `combine_abc` kernel per level
forms P_l(sigma) on that level's pattern, one Gauss-Jordan kernel inverts the coarsest.
Per application: V(nu,nu) cycle of damped-Jacobi sweeps (SELL-32 on levels >= 250 k rows, CSR-vector
below), CSR SpMVs for restriction / prolongation and a dense GEMV on the coarsest level -- all
libhx_b200 kernels, in complex64 by default, captured once per shift in a CUDA graph and replayed.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .backend import CsrMatrix
from .phases import phase

c128 = torch.complex128
f64 = torch.float64


# ---- torch sparse helpers (set-up only) -------------------------------------------------
def _rows_of(indptr, nnz):
    n = indptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n, device=indptr.device), (indptr[1:] - indptr[:-1]).long(),
                                   output_size=nnz)


def _to_coo(M: CsrMatrix, values=None):
    vals = M.values if values is None else values
    rows = _rows_of(M.indptr, M.nnz)
    return torch.sparse_coo_tensor(torch.stack([rows, M.indices.long()]), vals, size=M.shape).coalesce()


def _from_coo(T) -> CsrMatrix:
    T = T.coalesce()
    idx = T.indices()
    n_rows, n_cols = T.shape
    counts = torch.bincount(idx[0], minlength=n_rows)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=idx.device)
    indptr[1:] = torch.cumsum(counts, 0)
    return CsrMatrix(n_rows, n_cols, indptr.to(torch.int32).contiguous(), idx[1].to(torch.int32).contiguous(),
                     T.values().contiguous())


#: cuSPARSE SpGEMM (behind torch.sparse.mm) fails with "insufficient resources" once the number
#: of intermediate products gets large (seen at 10 M DoF); the left operand is therefore
#: processed in row blocks bounded by this many estimated products.
_SPMM_MAX_PRODUCTS = 3.0e8


def _spmm(A, B):
    A = A.coalesce()
    B = B.coalesce()
    nA, nB = A._nnz(), B._nnz()
    est = float(nA) * (float(nB) / max(B.shape[0], 1))
    if est <= _SPMM_MAX_PRODUCTS or A.shape[0] < 2:
        return torch.sparse.mm(A, B).coalesce()
    nblk = int(min(A.shape[0], np.ceil(est / _SPMM_MAX_PRODUCTS)))
    rows = A.indices()[0]
    bounds = torch.linspace(0, A.shape[0], nblk + 1, device=rows.device).long()
    cut = torch.searchsorted(rows, bounds)            # coalesced => rows sorted
    idx_out, val_out = [], []
    for b in range(nblk):
        lo, hi = int(cut[b]), int(cut[b + 1])
        if hi == lo:
            continue
        r0, r1 = int(bounds[b]), int(bounds[b + 1])
        sub_idx = A.indices()[:, lo:hi].clone()
        sub_idx[0] -= r0
        sub = torch.sparse_coo_tensor(sub_idx, A.values()[lo:hi], size=(r1 - r0, A.shape[1])).coalesce()
        Cb = torch.sparse.mm(sub, B).coalesce()
        ci = Cb.indices().clone()
        ci[0] += r0
        idx_out.append(ci)
        val_out.append(Cb.values())
        del Cb, sub
    return torch.sparse_coo_tensor(torch.cat(idx_out, 1), torch.cat(val_out), size=(A.shape[0], B.shape[1])).coalesce()


def _diag(M: CsrMatrix):
    rows = _rows_of(M.indptr, M.nnz)
    mask = M.indices.long() == rows
    d = torch.zeros(M.n_rows, dtype=M.values.dtype, device=M.values.device)
    d[rows[mask]] = M.values[mask]
    return d


def morton_order(coords, bits=16):
    """Permutation sorting points along a 3-D Z-order curve."""
    lo = coords.min(dim=0).values
    ext = (coords.max(dim=0).values - lo).max().clamp_min(1e-300)
    q = ((coords - lo) / ext * (2 ** bits - 1)).long().clamp_(0, 2 ** bits - 1)
    key = torch.zeros(coords.shape[0], dtype=torch.int64, device=coords.device)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> b) & 1) << (3 * b + d)
    return torch.sort(key, stable=True).indices


def _values_on(pattern_keys, T, n_cols):
    """Values of coalesced COO tensor T laid out on the sorted key list pattern_keys."""
    idx = T.indices()
    keys = idx[0] * n_cols + idx[1]
    pos = torch.searchsorted(pattern_keys, keys)
    out = torch.zeros(pattern_keys.numel(), dtype=T.values().dtype, device=keys.device)
    out[pos] = T.values()
    return out


def _real(v):
    return v.real.contiguous() if v.is_complex() else v


class Level:
    pass


import ctypes as _C


class TailDesc(_C.Structure):
    """hx_tail_desc (include/hx_b200.h)."""
    _fields_ = [("n", _C.c_int32), ("nc", _C.c_int32), ("nu", _C.c_int32), ("pad_", _C.c_int32),
                ("a_ptr", _C.c_void_p), ("a_idx", _C.c_void_p), ("a_val", _C.c_void_p), ("dinv", _C.c_void_p),
                ("r_ptr", _C.c_void_p), ("r_idx", _C.c_void_p), ("r_val", _C.c_void_p),
                ("p_ptr", _C.c_void_p), ("p_idx", _C.c_void_p), ("p_val", _C.c_void_p),
                ("coarse_inv", _C.c_void_p), ("omega", _C.c_float * 4),
                ("buf0", _C.c_void_p), ("buf1", _C.c_void_p), ("r", _C.c_void_p), ("bc", _C.c_void_p), ("xc", _C.c_void_p),
                ("barrier", _C.c_void_p)]


def filter_prolongator(P: CsrMatrix, theta):
    """Drop the entries of a smoothed prolongator below theta * (largest entry of the row) and rescale the
    kept ones so that every row keeps its sum (constants stay in the range of P).  The Galerkin operators
    R A P inherit the square of the saving: without it the level operators of the 10 M-DoF annulus carry
    14.8 / 70 / 473 / 1464 nonzeros per row and the coarse levels cost as much as the fine one."""
    if not theta:
        return P
    n, dev = P.n_rows, P.values.device
    rows = _rows_of(P.indptr, P.nnz)
    a = P.values.abs()
    rmax = torch.zeros(n, dtype=a.dtype, device=dev).scatter_reduce_(0, rows, a, reduce="amax", include_self=True)
    keep = a >= theta * rmax[rows]
    rsum = torch.zeros(n, dtype=P.values.dtype, device=dev).index_add_(0, rows, P.values)
    ksum = torch.zeros(n, dtype=P.values.dtype, device=dev).index_add_(0, rows[keep], P.values[keep])
    scale = torch.where(ksum.abs() > 1e-300, rsum / ksum, torch.ones_like(ksum))
    kr = rows[keep]
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(kr, minlength=n), 0)
    return CsrMatrix(n, P.n_cols, indptr.to(torch.int32).contiguous(), P.indices[keep].contiguous(),
                     (P.values[keep] * scale[kr]).contiguous())


#: how the spectral radius of D^-1 K is estimated for the prolongator smoothing:
#: "smoothpower" (default) = 12 power iterations from a smooth start vector: underestimates rho (1.45-1.6 vs ~2),
#: i.e. over-relaxes the prolongator smoothing, which measured BEST (25 vs 30 GMRES iterations at 60k DoF,
#: 36 vs 45 at 250k); "gershgorin" = max_i sum_j |k_ij| / k_ii (safe upper bound); "power" = random start
RHO_MODE = "smoothpower"
#: prolongator filter threshold (filter_prolongator) and aggregate size from level 1 down; HX_AMG_PFILTER /
#: HX_AMG_AGG_COARSE override.  Timed on B200 (profiles/r2_hier_*.json, 5 M DoF, inner-solve seconds of one step):
#: 16 / no filter 13.67; 32 / no filter 12.29; 16 / 0.1 12.93; 32 / 0.1 11.81 (with the W-cycle from level 1: 10.85);
#: 32 / 0.15 11.97; 32 / 0.2 12.22; 24 / 0.1 12.61; 48 / 0.1 13.23; 64 / no filter 13.17.  Level nonzeros
#: 73.8 M / 21.4 M / 8.5 M / 1.1 M become 73.8 M / 17.9 M / 1.8 M / 37 k: one visit of level 2 and below 185 -> 84 us.
P_FILTER = 0.1
AGG_COARSE = 32
#: fine-level rows from which the W-cycle pays on B200 (see AMG.__init__).  With the coarse levels above
#: (profiles/r2_shape2_*.json, whole step): 250 k DoF V 1.00 s / W from level 1 1.12 s; 1 M 3.22 / 2.98 s
#: (level 2 alone twice: 2.94 s); 10 M: W from level 1 27.1 s, level 2 alone twice 31.0 s.
W_AUTO_MIN_ROWS = 600_000


class AMG:
    def __init__(self, be, A: CsrMatrix, C: CsrMatrix, B: CsrMatrix | None, coords, tau=None, agg_size=16,
                 coarse_max=900, max_levels=12, nu=2, omega=2.0 / 3.0, sell_min_rows=250000, precision="single",
                 w_from=None, smoother=None, w_to=None, p_filter=None, nu_coarse=None, agg_coarse=None):
        """A, C real-valued, B complex or None -- all on ONE shared fine pattern.
        precision="single": the V-cycle (smoother, residual, transfers) runs in complex64 --
        it is only a preconditioner; GMRES and everything outside stay complex128."""
        self.be, self.nu, self.omega = be, nu, omega
        # W-cycle: levels >= w_from are visited twice per visit of their parent (None: V-cycle).  The
        # V-cycle count grows with the number of levels, the W-cycle count hardly does; the extra coarse
        # visits are latency-bound.  Default (HX_AMG_WCYCLE unset or "auto"): W from level 1 on meshes of at
        # least W_AUTO_MIN_ROWS rows, V below; HX_AMG_WCYCLE=<from>[:<to>] forces a shape, "off" the V-cycle.
        # damping of sweep s (pre- and post-smoothing alike); a list makes the nu sweeps a polynomial
        # smoother with those roots (e.g. the Chebyshev pair) at no extra cost
        self.omegas = list(omega) if isinstance(omega, (list, tuple)) else [omega] * nu
        if len(self.omegas) != nu:
            raise ValueError("omega: one damping factor, or one per sweep")
        # smoother="chebyshev": the nu sweeps of a level use the roots of the degree-nu Chebyshev
        # polynomial on [rho_l/8, rho_l], rho_l = 1.1 x a 20-step power-iteration estimate of rho(D^-1 P_l) made at
        # every shift -- same kernels and cost as damped Jacobi, different damping per sweep and level.
        # The default since it was timed on B200 (profiles/r2_ab_switches.md): whole step at 1 M DoF 5.35 ->
        # 4.31 s (inner iterations 7410 -> 5666); HX_AMG_SMOOTHER=jacobi restores constant damping.
        self.smoother = smoother or os.environ.get("HX_AMG_SMOOTHER", "chebyshev")
        if self.smoother not in ("jacobi", "chebyshev"):
            raise ValueError("smoother must be 'jacobi' or 'chebyshev'")
        if w_from is None:
            env = os.environ.get("HX_AMG_WCYCLE", "auto")
            if env == "auto":
                w_from = 1 if A.n_rows >= W_AUTO_MIN_ROWS else None
            elif env not in ("off", ""):
                # "<from>" or "<from>:<to>": levels from..to (inclusive) are visited twice per visit of their parent
                lo, _, hi = env.partition(":")
                w_from, w_to = int(lo), (int(hi) if hi else w_to)
        elif w_from == "off":
            w_from = None
        self.w_from = w_from
        self.w_to = w_to if w_to is not None else 10 ** 6
        self.native_min_rows = 20000
        # sweeps per smoothing on levels >= 2 and aggregate size from level 1 down (the coarse operators of
        # smoothed aggregation are dense -- 70 / 470 nonzeros per row on levels 1 / 2 -- so what is spent there
        # is tuned separately from the fine level)
        self.nu_coarse = int(os.environ.get("HX_AMG_NU_COARSE", nu)) if nu_coarse is None else int(nu_coarse)
        self.agg_coarse = int(os.environ.get("HX_AMG_AGG_COARSE", max(AGG_COARSE, agg_size))) if agg_coarse is None else int(agg_coarse)
        self.p_filter = float(os.environ.get("HX_AMG_PFILTER", P_FILTER)) if p_filter is None else float(p_filter)
        # the V-cycle is a fixed sequence of ~25 small launches on fixed buffers: captured once per
        # shift in a CUDA graph and replayed (HX_AMG_GRAPH=0 launches it kernel by kernel)
        self.use_graph = getattr(be, "supports_graphs", False) and os.environ.get("HX_AMG_GRAPH", "1") != "0"
        self._graph = None
        self.single = precision == "single" and getattr(be, "supports_mixed", False)
        self.wdtype = torch.complex64 if self.single else c128
        self.sell_min_rows = sell_min_rows if getattr(be, "supports_sell", False) else None
        dev = A.values.device
        self.levels = []
        a_re = A.values
        c_re = C.values
        b_cx = B.values if B is not None else None
        if tau is None:
            # shift making K = -A + tau C safely SPD and of the size of the operator's low spectrum
            dA = _diag(A).abs().max()
            dC = _diag(C).abs().max()
            tau = float(1e-3 * dA / dC)
        pat = A
        coords = coords.to(dev)
        while True:
            L = Level()
            L.n = pat.n_rows
            L.pattern = pat
            L.a, L.c, L.b = a_re, c_re, b_cx
            self.levels.append(L)
            if L.n <= coarse_max or len(self.levels) >= max_levels:
                break
            n = L.n
            order = morton_order(coords)
            agg = torch.empty(n, dtype=torch.int64, device=dev)
            agg[order] = torch.arange(n, device=dev) // (agg_size if len(self.levels) == 1 else self.agg_coarse)
            nc = int(agg.max().item()) + 1
            L.agg = agg                           # dof -> aggregate (ownership of the coarse dofs on several GPUs)
            cnt = torch.bincount(agg, minlength=nc).to(f64)
            tval = 1.0 / torch.sqrt(cnt[agg])
            if getattr(be, "supports_spgemm", False) and n >= self.native_min_rows:
                from .spgemm import Overflow
                try:
                    with phase("amg_setup_native_spgemm"):
                        L.P, L.R, pat, a_re, c_re, b_cx = self._coarsen_native(be, pat, a_re, c_re, b_cx, agg, nc, tval, tau)
                    csum = torch.zeros(nc, 3, dtype=f64, device=dev)
                    csum.index_add_(0, agg, coords)
                    coords = csum / cnt.view(-1, 1)
                    continue
                except Overflow:
                    pass          # a product row exceeds the kernel's shared-memory budget: library path below
            T = torch.sparse_coo_tensor(torch.stack([torch.arange(n, device=dev), agg]), tval, size=(n, nc)).coalesce()
            kval = -_real(a_re) + tau * _real(c_re)
            K = _to_coo(pat, kval)
            d = _diag(pat.with_values(kval))
            rows = K.indices()[0]
            S = torch.sparse_coo_tensor(K.indices(), K.values() / d[rows], size=K.shape).coalesce()
            # spectral radius of D^-1 K: Gershgorin bound (a power iteration started from a smooth
            # vector underestimates it badly on large meshes, which over-smooths the prolongator)
            v = None
            if RHO_MODE == "power":
                gen = torch.Generator().manual_seed(0)
                v = torch.randn(n, 1, dtype=f64, generator=gen).to(dev)
                rho = 1.0
                for _ in range(20):
                    v = torch.sparse.mm(S, v)
                    rho = float(torch.linalg.norm(v))
                    v = v / rho
            elif RHO_MODE == "smoothpower":
                v = torch.ones(n, 1, dtype=f64, device=dev) + 0.1 * torch.sin(torch.arange(n, device=dev, dtype=f64)).view(-1, 1)
                rho = 1.0
                for _ in range(12):
                    v = torch.sparse.mm(S, v)
                    rho = float(torch.linalg.norm(v))
                    v = v / rho
            else:
                rowsum = torch.zeros(n, dtype=f64, device=dev)
                rowsum.index_add_(0, rows, S.values().abs())
                rho = float(rowsum.max())
            self.rhos = getattr(self, "rhos", []) + [round(rho, 4)]
            ST = _spmm(S, T)
            Pm = (T - (4.0 / (3.0 * rho)) * ST).coalesce()
            del K, S, ST, T, v
            if self.p_filter:
                Pf = filter_prolongator(_from_coo(Pm), self.p_filter)
                Pm = _to_coo(Pf)
            Rm = Pm.t().coalesce()

            def galerkin(vals):
                X = _to_coo(pat, vals)
                Y = _spmm(X, Pm)
                del X
                Z = _spmm(Rm, Y)
                del Y
                if vals.is_cuda and n > 3_000_000:
                    torch.cuda.empty_cache()      # the SpGEMM temporaries are GBs at 10 M DoF
                return Z
            Ac = galerkin(_real(a_re))
            Cc = galerkin(_real(c_re))
            mats = [Ac, Cc]
            Ai = galerkin(a_re.imag.contiguous()) if a_re.is_complex() else None
            Ci = galerkin(c_re.imag.contiguous()) if c_re.is_complex() else None
            mats += [m_ for m_ in (Ai, Ci) if m_ is not None]
            if b_cx is not None:
                Br = galerkin(b_cx.real.contiguous())
                Bi = galerkin(b_cx.imag.contiguous())
                mats += [Br, Bi]
            keys = torch.unique(torch.cat([m.indices()[0] * nc + m.indices()[1] for m in mats]))
            crow, ccol = keys // nc, keys % nc
            counts = torch.bincount(crow, minlength=nc)
            indptr = torch.zeros(nc + 1, dtype=torch.int64, device=dev)
            indptr[1:] = torch.cumsum(counts, 0)
            L.P = _from_coo(Pm)
            L.R = _from_coo(Rm)
            pat = CsrMatrix(nc, nc, indptr.to(torch.int32).contiguous(), ccol.to(torch.int32).contiguous(),
                            torch.zeros(keys.numel(), dtype=f64, device=dev))
            a_re = _values_on(keys, Ac, nc)
            c_re = _values_on(keys, Cc, nc)
            if Ai is not None:
                a_re = torch.complex(a_re, _values_on(keys, Ai, nc))
            if Ci is not None:
                c_re = torch.complex(c_re, _values_on(keys, Ci, nc))
            if b_cx is not None:
                b_cx = torch.complex(_values_on(keys, Br, nc), _values_on(keys, Bi, nc))
            csum = torch.zeros(nc, 3, dtype=f64, device=dev)
            csum.index_add_(0, agg, coords)
            coords = csum / cnt.view(-1, 1)
        # work vectors
        wd = self.wdtype
        for li, L in enumerate(self.levels):
            L.nu = self.nu if li < 2 else self.nu_coarse
            L.x = be.zeros(L.n, dtype=wd); L.b_ = be.zeros(L.n, dtype=wd)
            L.r = be.zeros(L.n, dtype=wd); L.t = be.zeros(L.n, dtype=wd)
            if self.w_from is not None:
                L.xs = be.zeros(L.n, dtype=wd); L.bs = be.zeros(L.n, dtype=wd)
            L.M = None
            L.Mop = None
            L.sellp = None
            L.dinv = be.zeros(L.n)
            L.dinv_w = L.dinv
            if self.single and hasattr(L, "P"):
                L.P = L.P.with_values(L.P.values.float())
                L.R = L.R.with_values(L.R.values.float())
            if hasattr(L, "P"):
                # transfer operators of the cycle: SELL-32 (one thread per row, coalesced) where the CSR-vector kernel
                # is latency-bound on the short rows -- prolongation 336 -> us at 10 M DoF (5.6 nonzeros per row)
                L.P_op, L.R_op = self._transfer_op(L.P), self._transfer_op(L.R)
        last = self.levels[-1]
        last.b64 = be.zeros(last.n)
        last.x64 = be.zeros(last.n)
        self.coarse_inv = None

    def _transfer_op(self, M):
        """SELL-32 copy of a float32 transfer operator on large levels (the CSR matrix otherwise)."""
        if (self.sell_min_rows is None or M.n_rows < self.sell_min_rows or M.values.dtype != torch.float32
                or os.environ.get("HX_AMG_SELL_TRANSFER", "1") == "0"):
            return M
        from .sell import SellMatrix, SellPattern
        p = SellPattern(self.be, M.indptr, M.indices, M.n_rows, M.n_cols)
        return SellMatrix(p, p.values_from_csr(M.values))

    def _coarsen_native(self, be, pat, a_re, c_re, b_cx, agg, nc, tval, tau):
        """One coarsening step with the library's own SpGEMM kernels (hx_spgemm_*): P = (I - w D^-1 K) T,
        R = P^T, and the Galerkin products on ONE symbolic pattern shared by A, C, Re B, Im B."""
        from . import spgemm
        n, dev = pat.n_rows, a_re.device
        self.native_levels = getattr(self, "native_levels", 0) + 1
        kval = -_real(a_re) + tau * _real(c_re)
        rows = _rows_of(pat.indptr, pat.nnz)
        d = _diag(pat.with_values(kval))
        S = pat.with_values((kval / d[rows]).contiguous())
        # same (deliberately low) spectral-radius estimate as the library path: 12 power steps from a smooth vector
        v = (torch.ones(n, dtype=f64, device=dev) + 0.1 * torch.sin(torch.arange(n, device=dev, dtype=f64))).to(c128)
        w = torch.zeros_like(v)
        rho = 1.0
        for _ in range(12):
            be.spmv(S, v, w)
            rho = float(torch.linalg.norm(w))
            v, w = w / rho, v
        self.rhos = getattr(self, "rhos", []) + [round(rho, 4)]
        T = CsrMatrix(n, nc, torch.arange(n + 1, device=dev, dtype=torch.int32), agg.to(torch.int32).contiguous(), tval.contiguous())
        ST = spgemm.multiply(be, S, T)
        pvals = (-(4.0 / (3.0 * rho))) * ST.values
        keys = _rows_of(ST.indptr, ST.nnz) * nc + ST.indices.long()                 # sorted (CSR order)
        pos = torch.searchsorted(keys, torch.arange(n, device=dev) * nc + agg)
        pvals[pos] += tval                                                          # + T (its entry lies in the pattern of S*T)
        P = filter_prolongator(ST.with_values(pvals.contiguous()), self.p_filter)
        R = spgemm.transpose(P)
        yp = spgemm.symbolic(be, pat, P)
        Ypat = CsrMatrix(n, nc, yp[0], yp[1], None)
        zp = spgemm.symbolic(be, R, Ypat)
        ybuf = be.empty(max(int(yp[1].numel()), 1), dtype=f64)

        def galerkin(vals):
            yv = spgemm.numeric(be, pat.with_values(vals.contiguous()), P, yp[0], yp[1], out=ybuf)
            return spgemm.numeric(be, R, CsrMatrix(n, nc, yp[0], yp[1], yv), zp[0], zp[1]).clone()
        def galerkin_any(vals):
            return torch.complex(galerkin(vals.real), galerkin(vals.imag)) if vals.is_complex() else galerkin(vals)
        a_c = galerkin_any(a_re)
        c_c = galerkin_any(c_re)
        b_c = None
        if b_cx is not None:
            nzb = torch.nonzero(b_cx).reshape(-1)
            if nzb.numel() * 8 < b_cx.numel():
                # B lives on the impedance boundary only: its Galerkin product runs on the compacted entries
                # (own small symbolic pattern), scattered afterwards into the pattern shared with A and C
                brow = rows[nzb]
                bptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
                bptr[1:] = torch.cumsum(torch.bincount(brow, minlength=n), 0)
                Bc = CsrMatrix(n, n, bptr.to(torch.int32).contiguous(), pat.indices[nzb].contiguous(), None)
                ypb = spgemm.symbolic(be, Bc, P)
                zpb = spgemm.symbolic(be, R, CsrMatrix(n, nc, ypb[0], ypb[1], None))
                parts = []
                for comp in (b_cx.real[nzb].contiguous(), b_cx.imag[nzb].contiguous()):
                    yv = spgemm.numeric(be, Bc.with_values(comp), P, ypb[0], ypb[1])
                    parts.append(spgemm.numeric(be, R, CsrMatrix(n, nc, ypb[0], ypb[1], yv), zpb[0], zpb[1]))
                nzc = int(zpb[1].numel())
                kb = _rows_of(zpb[0], nzc) * nc + zpb[1].long()
                kall = _rows_of(zp[0], int(zp[1].numel())) * nc + zp[1].long()
                b_c = torch.zeros(int(zp[1].numel()), dtype=c128, device=dev)
                b_c[torch.searchsorted(kall, kb)] = torch.complex(parts[0][:nzc], parts[1][:nzc])
            else:
                b_c = torch.complex(galerkin(b_cx.real), galerkin(b_cx.imag))
        pat_c = CsrMatrix(nc, nc, zp[0], zp[1], torch.zeros(int(zp[1].numel()), dtype=f64, device=dev))
        return P, R, pat_c, a_c, c_c, b_c

    @property
    def sizes(self):
        return [L.n for L in self.levels]

    @property
    def operator_complexity(self):
        return sum(L.pattern.nnz for L in self.levels) / self.levels[0].pattern.nnz

    def set_shift(self, ca, cb, cc, fine_values=None):
        """Form P_l = ca*A_l + cb*B_l + cc*C_l on every level; invert the coarsest."""
        be = self.be
        self._graph = None                    # the captured cycle points at the previous shift's operators
        for i, L in enumerate(self.levels):
            if i == 0 and fine_values is not None:
                vals = fine_values
            else:
                vals = be.empty(L.pattern.nnz)
                be.combine_abc(L.a, L.b, L.c, ca, cb, cc, vals)
            L.M = L.pattern.with_values(vals)
            L.Mop = L.M
            L.M64 = L.M                      # complex128 operator (outer GMRES apply on the fine level)
            if i < len(self.levels) - 1:
                be.diag_inv(L.M, L.dinv)
                L.dinv_w = L.dinv.to(self.wdtype) if self.single else L.dinv
                L.omegas = self._chebyshev_dampings(L) if self.smoother == "chebyshev" else (self.omegas * 2)[:L.nu]
                if self.sell_min_rows is not None and L.n >= self.sell_min_rows:
                    from .sell import SellMatrix, SellPattern
                    if L.sellp is None:
                        L.sellp = SellPattern(be, L.pattern.indptr, L.pattern.indices, L.n, L.n)
                        L.sell_vals = be.empty(max(L.sellp.total, 1), dtype=self.wdtype)
                        L.sell_vals64 = be.empty(max(L.sellp.total, 1)) if (self.single and i == 0) else None
                    # kernel variant (unroll, min CTAs/SM) measured on B200 (profiles/r1_profile_parts_*.json):
                    # complex64 sweeps on a million-row level 37.0 us with (4,6) against 43.5 us with (4,4)
                    L.Mop = SellMatrix(L.sellp, L.sellp.values_from_csr(vals, out=L.sell_vals),
                                       variant=2 if (self.single and i == 0) else 0)
                    if i == 0:
                        L.M64 = L.Mop if not self.single else SellMatrix(L.sellp, L.sellp.values_from_csr(vals, out=L.sell_vals64))
                elif self.single:
                    L.Mop = L.pattern.with_values(vals.to(self.wdtype))
        L = self.levels[-1]
        dense = torch.zeros(L.n, L.n, dtype=c128, device=L.M.values.device)
        rows = _rows_of(L.M.indptr, L.M.nnz)
        # column-major storage of the coarse matrix == row-major storage of its transpose
        dense[L.M.indices.long(), rows] = L.M.values
        info = be.dense_inverse(dense)
        if int(info[0]) != 0:          # one read per shift: an exactly singular coarsest operator would fill the cycle with NaNs
            raise RuntimeError(f"multigrid: the coarsest operator is singular (zero pivot in column {int(info[0]) - 1})")
        self.coarse_inv = dense
        self._coarse_info = info
        self._tail = self._tail_descriptor()
        return self

    def _tail_descriptor(self):
        """Descriptor of the fused cycle tail (last smoothed level + dense coarsest solve, hx_amg_tail), or None
        when it does not apply: one level, complex128 cycle, that level stored as SELL, or not asked for.
        OFF by default (HX_AMG_TAIL=1 switches it on).  Measured on B200 (profiles/r2_tail2_*.json, 4 CTAs per SM):
        launched kernel by kernel one visit of the tail drops from 78 to 35 us at 1 M DoF and from 88 to 78 us at 8 M
        DoF -- but the production cycle is replayed from a CUDA graph, where the ten small kernels of the unfused
        tail already run back to back: the replayed cycle takes 498 us fused against 494 us unfused at 1 M DoF and the
        whole step 2.95 s against 2.92 s (8 M: 20.3 s against 20.1 s).  The graph had already removed the latency the
        fusion was after; kept as a tested option.  (A first version with one CTA per SM was slower.)"""
        if (len(self.levels) < 2 or not self.single or not getattr(self.be, "supports_tail", False)
                or os.environ.get("HX_AMG_TAIL", "0") != "1"):
            return None
        L, Lc = self.levels[-2], self.levels[-1]
        if getattr(L.Mop, "is_sell", False) or L.Mop.values.dtype != torch.complex64 or not 1 <= L.nu <= 4:
            return None
        if L.P.values.dtype != torch.float32 or L.R.values.dtype != torch.float32:
            return None
        d = TailDesc()
        d.n, d.nc, d.nu = L.n, Lc.n, L.nu
        d.a_ptr, d.a_idx, d.a_val = L.Mop.indptr.data_ptr(), L.Mop.indices.data_ptr(), L.Mop.values.data_ptr()
        d.dinv = L.dinv_w.data_ptr()
        d.r_ptr, d.r_idx, d.r_val = L.R.indptr.data_ptr(), L.R.indices.data_ptr(), L.R.values.data_ptr()
        d.p_ptr, d.p_idx, d.p_val = L.P.indptr.data_ptr(), L.P.indices.data_ptr(), L.P.values.data_ptr()
        d.coarse_inv = self.coarse_inv.data_ptr()
        for s_, om in enumerate(L.omegas):
            d.omega[s_] = float(om)
        if not hasattr(L, "tail_bar"):
            L.tail_bar = torch.zeros(2, dtype=torch.int32, device=L.dinv.device)
            L.tail_x, L.tail_t = L.x, L.t                # fixed roles: the result always lands in tail_x
        swaps = 2 * L.nu - 1
        d.buf0, d.buf1 = (L.tail_t.data_ptr(), L.tail_x.data_ptr()) if swaps % 2 else (L.tail_x.data_ptr(), L.tail_t.data_ptr())
        d.r, d.bc, d.xc = L.r.data_ptr(), Lc.b64.data_ptr(), Lc.x64.data_ptr()
        d.barrier = L.tail_bar.data_ptr()
        self._tail_keep = (L.Mop, L.dinv_w, self.coarse_inv)     # the descriptor holds raw pointers
        return d

    def fine_matrix(self):
        return self.levels[0].M

    def fine_operator(self):
        """The fine-level complex128 operator in its fastest SpMV format (SELL-32 when large)."""
        return self.levels[0].M64

    def _chebyshev_dampings(self, L, iters=20, safety=1.1, alpha=8.0):
        """1 / (roots of the degree-nu Chebyshev polynomial on [rho/alpha, rho]) for rho(D^-1 M_l)."""
        be = self.be
        gen = torch.Generator(device="cpu").manual_seed(1234 + L.n)
        v = torch.randn(L.n, dtype=f64, generator=gen).to(L.dinv.device).to(c128)
        w = be.zeros(L.n)
        rho = 0.0
        for _ in range(iters):
            v = v / torch.linalg.vector_norm(v)
            be.spmv(L.M, v.contiguous(), w)
            v = w * L.dinv
            rho = float(torch.linalg.vector_norm(v))
        rho *= safety
        a, b = rho / alpha, rho
        mid, half = 0.5 * (a + b), 0.5 * (b - a)
        roots = [mid + half * np.cos(np.pi * (2 * k + 1) / (2 * L.nu)) for k in range(L.nu)]
        L.rho = rho
        return [float(1.0 / r) for r in sorted(roots, reverse=True)]

    def _smooth(self, L, b, x, first_zero):
        """nu damped-Jacobi sweeps; result ends in L.x.  x is L.x."""
        be = self.be
        cur, other = L.x, L.t
        om = L.omegas
        k = 0
        if first_zero:
            be.jacobi_sweep(L.Mop, L.dinv_w, b, None, cur, om[0])
            k = 1
        for s in range(k, L.nu):
            be.jacobi_sweep(L.Mop, L.dinv_w, b, cur, other, om[s])
            cur, other = other, cur
        if cur is not L.x:
            L.x, L.t = cur, other      # swap the roles of the buffers

    def _cycle(self, i, b):
        """Solve approximately M_i x = b; result in self.levels[i].x."""
        be = self.be
        L = self.levels[i]
        if i == len(self.levels) - 2 and getattr(self, "_tail", None) is not None:
            be.amg_tail(self._tail, b)                    # this level and the coarsest one in one persistent kernel
            L.x, L.t = L.tail_x, L.tail_t
            return L.x
        if i == len(self.levels) - 1:
            if self.single:                       # the coarsest solve stays in double precision
                L.b64.copy_(b)
                be.dense_gemv(self.coarse_inv, L.b64, L.x64)
                L.x.copy_(L.x64)
            else:
                be.dense_gemv(self.coarse_inv, b, L.x)
            return L.x
        self._smooth(L, b, L.x, first_zero=True)
        be.spmv(L.Mop, L.x, L.r, alpha=-1.0, beta=1.0, y0=b)         # r = b - M x
        Lc = self.levels[i + 1]
        be.spmv(L.R_op, L.r, Lc.b_)
        xc = self._cycle(i + 1, Lc.b_)
        if self.w_from is not None and self.w_from <= i + 1 <= self.w_to and i + 1 < len(self.levels) - 1:
            # second visit: the cycle applied to the coarse residual corrects xc
            Lc.xs.copy_(xc)
            Lc.bs.copy_(Lc.b_)
            be.spmv(Lc.Mop, Lc.xs, Lc.b_, alpha=-1.0, beta=1.0, y0=Lc.bs)
            xc = self._cycle(i + 1, Lc.b_)
            xc.add_(Lc.xs)
        be.spmv(L.P_op, xc, L.x, alpha=1.0, beta=1.0, y0=L.x)          # x += P xc
        self._smooth(L, b, L.x, first_zero=False)
        return L.x

    def _capture(self):
        """Record one V-cycle (input levels[0].v_w, output the fine-level x buffer) into a CUDA graph."""
        be, L0 = self.be, self.levels[0]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._cycle(0, L0.v_w)                               # eager pass: lazy buffers exist before capture
            n0 = be.launch_count()
            g = torch.cuda.CUDAGraph()
            g.capture_begin(capture_error_mode="thread_local")
            try:
                x = self._cycle(0, L0.v_w)
            finally:
                g.capture_end()
            self._graph_kernels = be.launch_count() - n0
            be.add_launches(-self._graph_kernels)                # recorded, not run
        cur.wait_stream(side)
        self._graph, self._graph_x = g, x

    def apply(self, v, out):
        """out = V-cycle(v)."""
        L0 = self.levels[0]
        if self.single or self.use_graph:
            if not hasattr(L0, "v_w"):
                L0.v_w = self.be.zeros(L0.n, dtype=self.wdtype)
            L0.v_w.copy_(v)                       # complex128 -> complex64 (or the graph's fixed input)
            v = L0.v_w
        if self.use_graph and len(self.levels) > 1:
            if self._graph is None:
                self._capture()
            self._graph.replay()
            self.be.add_launches(self._graph_kernels)
            out.copy_(self._graph_x)
            return out
        x = self._cycle(0, v)
        out.copy_(x)                              # (complex64 ->) complex128
        return out
