"""Host-side Krylov logic: restarted GMRES (inner solve behind the spectral
transformation) and Krylov-Schur (what SLEPc's EPS 'krylovschur' does for
helmholtz_x/eigensolvers.py:41-67).  All vector work happens in backend kernels
(block classical Gram-Schmidt with re-orthogonalisation: multi_dot -> multi_axpy twice);
the host only sees the (m+1) x m projected matrix.
"""
from __future__ import annotations

import numpy as np
import torch

c128 = torch.complex128


class ArnoldiBasis:
    """(m+1) x n basis plus the device buffers of one CGS2 step."""

    def __init__(self, be, n, m):
        self.be, self.n, self.m = be, n, m
        self.V = be.zeros(m + 1, n)
        self.w = be.zeros(n)
        self.h1 = be.zeros(m + 2)
        self.h2 = be.zeros(m + 2)
        self._h1r = torch.view_as_real(self.h1)
        # one GPU: the Hessenberg column comes back through a pinned buffer on a copy stream, so the
        # caller can queue the next operator application before waiting for it
        self.pipelined = getattr(be, "supports_pipelining", False)
        if self.pipelined:
            self._host = torch.empty(m + 2, dtype=c128).pin_memory()
            self._copy_stream = torch.cuda.Stream(device=be.device)
            self._ready = torch.cuda.Event()
            self._done = torch.cuda.Event()

    def orthogonalize_begin(self, j, passes=2):
        """Queue classical Gram-Schmidt of self.w against V[0..j] (passes=2: with re-orthogonalisation,
        CGS2); the normalised result goes to V[j+1]."""
        be, V, w, k = self.be, self.V, self.w, j + 1
        nrm2 = self._h1r[k]          # real part of h1[k] receives ||w||^2
        be.multi_dot(V, k, w, self.h1)
        if passes == 1:
            # (beta from <w,w> - |h|^2 would save this second reduction, but single-pass classical
            # Gram-Schmidt loses orthogonality like eps * kappa^2 while the residual drops by 1e11 in one
            # cycle, and the Pythagorean beta then compounds the error: measured 63 instead of 26
            # iterations on the Rijke fixture.  The explicit norm stays.)
            be.multi_axpy(V, k, self.h1, w, nrm2=nrm2)
        else:
            be.multi_axpy(V, k, self.h1, w)
            be.multi_dot(V, k, w, self.h2)
            be.multi_axpy(V, k, self.h2, w, hacc=self.h1, nrm2=nrm2)
        be.scale_copy(w, V[j + 1], nrm2=nrm2)
        if self.pipelined:
            self._ready.record(torch.cuda.current_stream(be.device))
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._ready)
                self._host[:k + 1].copy_(self.h1[:k + 1], non_blocking=True)
                self._done.record(self._copy_stream)

    def orthogonalize_end(self, j):
        """(h[0..j], beta) of the step queued by orthogonalize_begin(j), on the host."""
        k = j + 1
        if self.pipelined:
            self._done.synchronize()
            host = self._host[:k + 1].numpy().copy()
        else:
            host = self.h1[:k + 1].cpu().numpy()
        beta = float(np.sqrt(max(host[k].real, 0.0)))
        return host[:k].copy(), beta

    def orthogonalize(self, j):
        self.orthogonalize_begin(j)
        return self.orthogonalize_end(j)


def gmres(be, apply_A, b, x, precond=None, rtol=1e-10, restart=60, maxiter=600, basis=None, work=None, zbasis=None,
          orth_passes=2, anorm=None, eta=1e-14, x0=False, info=None):
    """Right-preconditioned restarted GMRES: solves A x = b, x overwritten (start 0; x0=True: start
    from the x passed in).  apply_A(v, out), precond(v, out).  Returns (iterations, relative residual).
    zbasis (restart x n): flexible GMRES -- the preconditioned vectors z_j = M^-1 v_j are kept
    and x += Z y, so the Arnoldi relation A Z = V H holds exactly even when the preconditioner
    is not an exact linear operator (the complex64 multigrid cycle).

    Every solve ends on a recomputed true residual r = b - A x.  It is accepted when
    ||r|| <= rtol ||b||, or -- anorm given (an estimate of ||A||) -- when the normwise backward error
    ||r|| / (anorm ||x||) <= eta: next to an eigenvalue of A (shift-invert at a converged shift, the
    Newton iteration's singular L(omega_k)) ||x|| ~ ||b|| / |lambda_min| and rtol ||b|| lies below
    what double precision can represent in A x; a backward-stable x is all an exact LU delivers
    there, too.  info (dict): receives status ("converged" | "backward_stable" | "stagnated" |
    "maxiter"), rel, eta (the backward error reached, None without anorm)."""
    n = b.numel()
    if basis is None:
        basis = ArnoldiBasis(be, n, restart)
    m = basis.m
    V, w = basis.V, basis.w
    z = work if work is not None else be.zeros(n)
    nb = be.zeros(2)
    info = info if info is not None else {}

    def norm(v):
        be.multi_dot(v.view(1, -1), 1, v, nb)
        return float(np.sqrt(max(nb[:1].cpu().numpy()[0].real, 0.0)))

    bnorm = norm(b)
    if not x0:
        x.zero_()
    if bnorm == 0.0:
        x.zero_()
        info.update(status="converged", rel=0.0, eta=0.0)
        return 0, 0.0
    total = 0
    rel = 1.0
    x_is_zero = not x0
    claimed = False                 # the recurrence of the previous cycle reported convergence
    last_true = np.inf
    status = "maxiter"
    eta_now = None
    while True:
        # true residual r = b - A x (the recurrence estimate is never the last word: single-pass
        # Gram-Schmidt, complex64 preconditioner)
        if x_is_zero:
            beta = bnorm
        else:
            apply_A(x, w)
            be.axpby(1.0, b, -1.0, w)             # w = b - w
            beta = norm(w)
        rel = beta / bnorm
        if rel <= rtol:
            status = "converged"
            break
        if anorm is not None and not x_is_zero:
            xnorm = norm(x)
            eta_now = beta / (anorm * xnorm) if xnorm > 0.0 else None
            if eta_now is not None and eta_now <= eta:
                status = "backward_stable"
                break
        if total >= maxiter:
            status = "maxiter"
            break
        if claimed:
            # the recurrence said converged, the recomputed residual does not: refine with further
            # cycles while they still help
            if rel > 0.5 * last_true:
                status = "stagnated"
                break
            last_true = rel
        claimed = False
        if x_is_zero:
            be.scale_copy(b, V[0], alpha=1.0 / beta)
        else:
            be.scale_copy(w, V[0], alpha=1.0 / beta)
        # Givens QR of the Hessenberg matrix, column by column, in plain Python complex arithmetic
        # (NumPy scalar operations cost ~1 us each and this loop runs once per iteration)
        H = np.zeros((m + 1, m), complex)
        g = [0j] * (m + 1)
        g[0] = complex(beta)
        cs = [0j] * m
        sn = [0j] * m
        j_used = 0

        def operator(jj):
            """w = A M^-1 v_jj (queued)."""
            if precond is not None and zbasis is not None:
                precond(V[jj], zbasis[jj])
                apply_A(zbasis[jj], w)
            elif precond is not None:
                precond(V[jj], z)
                apply_A(z, w)
            else:
                apply_A(V[jj], w)

        queued = False
        for j in range(m):
            if not queued:
                operator(j)
            basis.orthogonalize_begin(j, orth_passes)
            queued = getattr(basis, "pipelined", False) and j + 1 < m and total + 2 <= maxiter
            if queued:
                # v_{j+1} is complete on the device: start the next operator application while the
                # host waits for column j and updates the QR (thrown away if column j converges)
                operator(j + 1)
            h, hb = basis.orthogonalize_end(j)
            total += 1
            col = h.tolist()
            col.append(complex(hb))
            for i in range(j):
                ci, si, u, v = cs[i], sn[i], col[i], col[i + 1]
                col[i] = ci * u + si * v
                col[i + 1] = -si.conjugate() * u + ci * v
            a, bb = col[j], col[j + 1]
            den = (abs(a) ** 2 + abs(bb) ** 2) ** 0.5
            if den == 0.0:
                cs[j], sn[j] = 1 + 0j, 0j
            elif a != 0:
                cs[j] = complex(abs(a) / den)
                sn[j] = (a / abs(a)) * bb.conjugate() / den
            else:
                cs[j], sn[j] = 0j, 1 + 0j
            col[j] = cs[j] * a + sn[j] * bb
            col[j + 1] = 0j
            g[j + 1] = -sn[j].conjugate() * g[j]
            g[j] = cs[j] * g[j]
            H[:j + 2, j] = col
            j_used = j + 1
            est = abs(g[j + 1]) / bnorm
            if est <= 0.9 * rtol or hb <= 1e-300:
                claimed = True
                break
            if total >= maxiter:
                break
        g = np.asarray(g)
        y = np.linalg.solve(np.triu(H[:j_used, :j_used]), g[:j_used])
        # x += M^{-1} (V y)
        yd = be.asarray(-y, dtype=c128)
        if precond is not None and zbasis is not None:
            be.multi_axpy(zbasis, j_used, yd, x)          # x += Z y
        else:
            w.zero_()
            be.multi_axpy(V, j_used, yd, w)
            if precond is not None:
                precond(w, z)
                be.axpby(1.0, z, 1.0, x)
            else:
                be.axpby(1.0, w, 1.0, x)
        x_is_zero = False
    info.update(status=status, rel=rel, eta=eta_now)
    return total, rel


def _reorder_schur(T, Z, order_key):
    """Reorder a complex Schur form so that diag(T) is sorted by order_key (descending)
    using adjacent Givens swaps (complex upper-triangular => always well defined)."""
    n = T.shape[0]
    T = T.copy(); Z = Z.copy()
    for i in range(n):
        # bring the best remaining eigenvalue to position i (bubble up)
        d = np.diag(T)
        j = i + int(np.argmax(order_key(d[i:])))
        for k in range(j - 1, i - 1, -1):
            a, b, c = T[k, k], T[k, k + 1], T[k + 1, k + 1]
            # Givens G with G^H [b; c-a] -> first column aligned so that T[k,k] <- c
            x = np.array([b, c - a])
            nx = np.linalg.norm(x)
            if nx == 0.0:
                continue
            cgs, sgs = x[0] / nx, x[1] / nx
            G = np.array([[cgs, -np.conj(sgs)], [sgs, np.conj(cgs)]])
            T[:, k:k + 2] = T[:, k:k + 2] @ G
            T[k:k + 2, :] = G.conj().T @ T[k:k + 2, :]
            Z[:, k:k + 2] = Z[:, k:k + 2] @ G
            T[k + 1, k] = 0.0
    return T, Z


def _triu_eigvecs(R):
    """Unit-norm eigenvectors of an upper-triangular matrix (back substitution)."""
    n = R.shape[0]
    Y = np.zeros((n, n), complex)
    small = np.finfo(float).eps * max(np.abs(R).max(), 1e-300)
    for i in range(n):
        y = np.zeros(n, complex)
        y[i] = 1.0
        for k in range(i - 1, -1, -1):
            d = R[k, k] - R[i, i]
            if abs(d) < small:
                d = small
            y[k] = -(R[k, k + 1:i + 1] @ y[k + 1:i + 1]) / d
        Y[:, i] = y / np.linalg.norm(y)
    return Y


def _wanted_residual(r, nev):
    theta, res, nconv = r[4], r[5], r[6]
    hi = min(nev, len(theta))
    if nconv >= hi:
        return None
    return float(max(res[i] / max(abs(theta[i]), 1e-300) for i in range(nconv, hi)))


class KrylovSchurResult:
    def __init__(self, theta, X, its, nconv, residuals, n_apply):
        self.theta, self.X, self.its, self.nconv, self.residuals, self.n_apply = theta, X, its, nconv, residuals, n_apply


def krylov_schur(be, apply_op, n, nev, ncv=None, tol=1e-10, maxit=100, v0=None, seed=0, n_global=None,
                 wanted_residual=None):
    """nev eigenpairs of largest |theta| of the operator apply_op(v, out) on C^n.
    SLEPc defaults: ncv = max(2 nev, nev+15) (eigensolvers.py:58 passes DECIDE),
    restart keeping half of the non-converged part.
    wanted_residual(r): called before every operator application with the largest relative Ritz
    residual of the wanted, not yet converged pairs (None while unknown) -- an inexact operator may
    loosen its own tolerance as 1/r (relaxed inexact Krylov, Bouras-Fraysse / Simoncini-Szyld)."""
    import scipy.linalg as sla
    if ncv is None:
        ncv = max(2 * nev, nev + 15)
    ng = n_global if n_global is not None else n           # n = local (owned) length on multi-GPU runs
    m = min(ncv, ng - 1) if ng > 2 else 1
    basis = ArnoldiBasis(be, n, m)
    V, w = basis.V, basis.w
    if v0 is None:
        part = getattr(be, "part", None)
        g = torch.Generator().manual_seed(seed + (1000 * part.rank if part is not None else 0))
        v0 = torch.randn(n, dtype=torch.float64, generator=g).to(c128)
    v0 = be.asarray(v0, dtype=c128).clone()
    nb = be.zeros(2)
    be.multi_dot(v0.view(1, -1), 1, v0, nb)
    be.scale_copy(v0, V[0], alpha=1.0 / float(np.sqrt(nb[:1].cpu().numpy()[0].real)))
    H = np.zeros((m + 1, m), complex)
    k = 0
    its = 0
    n_apply = 0
    Vnew = None
    r_wanted = None

    def ritz(mm):
        """Ordered Schur form of the current mm x mm projection and the Ritz residual estimates."""
        T, Z = sla.schur(H[:mm, :mm], output="complex")
        T, Z = _reorder_schur(T, Z, np.abs)
        bt = H[mm, :mm] @ Z
        Y = _triu_eigvecs(T)
        theta = np.diag(T).copy()
        res = np.abs(bt @ Y)
        nconv = 0
        while nconv < mm and res[nconv] <= tol * abs(theta[nconv]):
            nconv += 1
        return T, Z, bt, Y, theta, res, nconv

    while True:
        its += 1
        mm = m
        early = None
        for j in range(k, m):
            if wanted_residual is not None:
                wanted_residual(r_wanted)
            apply_op(V[j], w)
            n_apply += 1
            h, hb = basis.orthogonalize(j)
            H[:j + 1, j] = h
            H[j + 1, j] = hb
            if hb <= 1e-14 * max(np.abs(h).max(), 1e-300):
                mm = j + 1        # invariant subspace found
                break
            # early exit: with a good start vector (warm-started fixed-point iterates) the wanted
            # pairs converge long before the basis is full -- test the small projection as we go
            if j + 1 >= max(nev + 2, k + 2) and j + 1 < m:
                early = ritz(j + 1)
                r_wanted = _wanted_residual(early, nev)
                if early[-1] >= nev:
                    mm = j + 1
                    break
                early = None
        T, Z, bt, Y, theta, res, nconv = early if early is not None else ritz(mm)
        r_wanted = _wanted_residual((T, Z, bt, Y, theta, res, nconv), nev)
        if nconv >= nev or its >= maxit or mm < m:
            break
        keep = min(max(nconv + (m - nconv) // 2, nev), m - 1)
        if Vnew is None:
            Vnew = be.zeros(m + 1, n)
        Q = be.asarray(np.ascontiguousarray(Z[:, :keep].T), dtype=c128)      # (keep, m)
        be.basis_rotate(V, m, Q, keep, Vnew)
        Vnew[keep].copy_(V[m])
        V[:keep + 1].copy_(Vnew[:keep + 1])
        H[:] = 0.0
        H[:keep, :keep] = T[:keep, :keep]
        H[keep, :keep] = bt[:keep]
        k = keep
    nout = min(max(nev, nconv), mm)
    # Ritz vectors X = V Z Y
    ZY = Z @ Y[:, :nout]
    X = be.zeros(nout, n)
    Q = be.asarray(np.ascontiguousarray(ZY.T), dtype=c128)
    be.basis_rotate(V, mm, Q, nout, X)
    return KrylovSchurResult(theta[:nout], X, its, nconv, res[:nout], n_apply)
