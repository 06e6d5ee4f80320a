"""CPU test double of helmholtz_x_b200.backend.CudaBackend -- TEST INFRASTRUCTURE ONLY.

Implements the backend interface with torch-CPU / NumPy / SciPy so that the
host-side logic of the product (Krylov-Schur restarts, GMRES, AMG cycle order,
fixed-point / Newton recurrences, rank partitioning) can be exercised by the
``-m "not gpu"`` tests.  It is never imported by the product package.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch

c128 = torch.complex128
f64 = torch.float64


class HostBackend:
    name = "host-test-double"
    supports_sell = False
    # HX_HOST_MIXED=1: run the multigrid cycle in complex64 as the CUDA backend does (the NumPy/SciPy
    # calls below follow the dtype of their operands), to exercise flexible GMRES on the CPU
    supports_mixed = os.environ.get("HX_HOST_MIXED", "0") == "1"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def zeros(self, *shape, dtype=c128):
        return torch.zeros(*shape, dtype=dtype)

    def empty(self, *shape, dtype=c128):
        return torch.zeros(*shape, dtype=dtype)

    def asarray(self, a, dtype=None):
        return torch.as_tensor(a, dtype=dtype).contiguous()

    def synchronize(self):
        pass

    @staticmethod
    def _sp(M):
        return sp.csr_matrix((M.values.numpy(), M.indices.numpy(), M.indptr.numpy()), shape=M.shape)

    def spmv(self, M, x, y, alpha=1.0, beta=None, y0=None, lanes=None):
        r = alpha * (self._sp(M) @ x.numpy()[:M.n_cols])
        if beta is not None:
            r = r + beta * (y0 if y0 is not None else y).numpy()[:M.n_rows]
        y[:M.n_rows] = torch.from_numpy(np.asarray(r, dtype=complex))
        self.launches += 1
        return y

    def combine_abc(self, a, b, c, ca, cb, cc, out):
        r = torch.zeros_like(out)
        if a is not None:
            r += ca * a
        if b is not None:
            r += cb * b
        if c is not None:
            r += cc * c
        out.copy_(r)
        return out

    def lowrank_dots(self, lr, x, t):
        xn = x.numpy()
        for f in range(lr.r):
            s, e = int(lr.rptr[f]), int(lr.rptr[f + 1])
            t[f] = complex(np.dot(lr.rval[s:e].numpy(), xn[lr.ridx[s:e].numpy()]))
        return t

    def lowrank_update(self, lr, t, coef, y):
        tn = t.numpy()
        for i in range(int(lr.lrow.numel())):
            s, e = int(lr.lptr[i]), int(lr.lptr[i + 1])
            y[int(lr.lrow[i])] += coef * complex(np.dot(lr.lval[s:e].numpy(), tn[lr.lcol[s:e].numpy()]))
        return y

    def multi_dot(self, V, k, w, out, conj=True):
        Vk = V[:k].numpy()
        out[:k] = torch.from_numpy((Vk.conj() if conj else Vk) @ w.numpy())
        self.launches += 1
        return out

    def multi_axpy(self, V, k, h, w, hacc=None, nrm2=None):
        hh = h[:k].numpy().copy()
        w -= torch.from_numpy(hh @ V[:k].numpy())
        if hacc is not None:
            hacc[:k] += torch.from_numpy(hh)
        if nrm2 is not None:
            nrm2[0] = float(np.vdot(w.numpy(), w.numpy()).real)
        self.launches += 1
        return w

    def scale_copy(self, w, out, nrm2=None, alpha=None):
        a = 1.0 / np.sqrt(float(nrm2[0])) if nrm2 is not None else (alpha if alpha is not None else 1.0)
        out.copy_(a * w)
        return out

    def axpby(self, a, x, b, y):
        if b is None:
            y.copy_(a * x)
        else:
            y.copy_(a * x + b * y)
        return y

    def basis_rotate(self, V, m, Q, kout, Vout):
        Vout[:kout] = torch.from_numpy(Q[:kout, :m].numpy() @ V[:m].numpy())
        return Vout

    def diag_inv(self, M, out):
        out.copy_(torch.from_numpy(1.0 / self._sp(M).diagonal().astype(complex)))
        return out

    def jacobi_sweep(self, M, dinv, b, xin, xout, omega):
        if xin is None:
            n = b.numel()
            xout[:n] = omega * dinv[:n] * b
        else:
            n = M.n_rows                          # rectangular (owned rows x owned+ghost columns) allowed
            r = b.numpy()[:n] - self._sp(M) @ xin.numpy()[:M.n_cols]
            xout[:n] = xin[:n] + omega * dinv[:n] * torch.from_numpy(r)
        self.launches += 1
        return xout

    def dense_inverse(self, A):
        A.copy_(torch.from_numpy(np.linalg.inv(A.numpy().T).T.copy()))
        return torch.zeros(A.shape[0] + 1, dtype=torch.int32)

    def dense_gemv(self, A, x, y):
        y.copy_(torch.from_numpy(A.numpy().T @ x.numpy()))
        return y

    def launch_count(self):
        return self.launches

    def reset_launch_count(self):
        self.launches = 0
