"""Device backend: torch owns memory/streams, every numerical kernel is a call into
libhx_b200.so.  Host-side algorithms (krylov.py, amg.py, eigensolvers.py) are written
against this small interface; ``tests`` substitute a NumPy/torch-CPU test double
(``oracle/host_backend.py``) to exercise that host logic without a GPU.  The product
never constructs anything but :class:`CudaBackend`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib

c128 = torch.complex128
f64 = torch.float64


def _c2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


def choose_lanes(nnz, n_rows):
    """Threads per row of the CSR-vector kernel (measured on B200, 10M-DoF annulus:
    14.8 nnz/row -> 4 lanes 4.6 TB/s, 8 lanes 4.0 TB/s; profiles/r1_spmv_bench_10m_v1.json)."""
    mean = nnz / max(n_rows, 1)
    if mean <= 18:
        return 4
    if mean <= 40:
        return 8
    if mean <= 96:
        return 16
    return 32


@dataclass
class CsrMatrix:
    """CSR matrix on the device: int32 indptr/indices, float64 or complex128 values."""
    n_rows: int
    n_cols: int
    indptr: torch.Tensor
    indices: torch.Tensor
    values: torch.Tensor

    @property
    def nnz(self):
        return int(self.indices.numel())

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def lanes(self):
        return choose_lanes(self.nnz, self.n_rows)

    def with_values(self, values):
        return CsrMatrix(self.n_rows, self.n_cols, self.indptr, self.indices, values)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.values.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                             shape=self.shape)


@dataclass
class LowRank:
    """sum_f left_f right_f^T kept as sparse vectors (flame operator, never densified).

    right: CSR-like over flames (rptr, ridx, rval); left: rows of the union support
    (lrow) each with (flame, value) pairs (lptr, lcol, lval)."""
    n: int
    r: int
    rptr: torch.Tensor
    ridx: torch.Tensor
    rval: torch.Tensor
    lrow: torch.Tensor
    lptr: torch.Tensor
    lcol: torch.Tensor
    lval: torch.Tensor


class CudaBackend:
    name = "cuda"
    supports_sell = True
    supports_mixed = True        # complex64 kernels for the multigrid cycle
    supports_spgemm = True       # hx_spgemm_* for the multigrid set-up
    supports_graphs = True       # CUDA-graph capture of fixed launch sequences (the multigrid cycle)
    supports_pipelining = True   # Hessenberg columns return through a pinned buffer on a copy stream
    supports_tail = True         # last smoothed level + coarsest solve of the cycle as one persistent kernel

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.HxLibraryError("helmholtz_x_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._scratch = {}

    # ---- plumbing -------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def zeros(self, *shape, dtype=c128):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def empty(self, *shape, dtype=c128):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def asarray(self, a, dtype=None):
        return torch.as_tensor(a, dtype=dtype, device=self.device).contiguous()

    def scratch(self, k):
        key = max(8, 1 << (max(k, 1) - 1).bit_length())
        if key not in self._scratch:
            nbytes = _lib.call("hx_reduce_scratch_bytes", key)
            self._scratch[key] = torch.zeros(nbytes // 8 + 1, dtype=f64, device=self.device)
        return self._scratch[key]

    def synchronize(self):
        torch.cuda.synchronize(self.device)

    # ---- sparse -----------------------------------------------------------------
    def spmv(self, M: CsrMatrix, x, y, alpha=1.0, beta=None, y0=None, lanes=None):
        """y = alpha*M@x (+ beta*y0)."""
        if getattr(M, "is_sell", False):
            return M.spmv(x, y, alpha=None if (alpha == 1.0 and beta is None) else alpha, beta=beta, y0=y0)
        assert x.dtype == y.dtype and x.is_contiguous() and y.is_contiguous()
        assert x.numel() >= M.n_cols and y.numel() >= M.n_rows
        if x.dtype == torch.complex64:
            name = "hx_spmv_cc" if M.values.dtype == torch.complex64 else "hx_spmv_sc"
            assert M.values.dtype in (torch.complex64, torch.float32)
        else:
            assert x.dtype == c128 and M.values.dtype in (c128, f64)
            name = "hx_spmv_zz" if M.values.dtype == c128 else "hx_spmv_dz"
        y0p = None
        if beta is not None:
            y0p = (y0 if y0 is not None else y).data_ptr()
        _lib.call(name, M.n_rows, M.indptr.data_ptr(), M.indices.data_ptr(), M.values.data_ptr(), x.data_ptr(),
                  y.data_ptr(), _c2(alpha), _c2(beta if beta is not None else 0.0), y0p, lanes or M.lanes, self.stream)
        return y

    def combine_abc(self, a, b, c, ca, cb, cc, out):
        """out = ca*a + cb*b + cc*c on one pattern.  a, c real (the assembled A, C) with complex b,
        or all complex (Bloch-reduced operators); one fused kernel either way."""
        nnz = out.numel()
        ptr = lambda t: t.data_ptr() if t is not None else None
        cplx = [t.dtype == c128 for t in (a, c) if t is not None]
        if cplx and all(cplx):
            _lib.call("hx_combine_zzz", nnz, ptr(a), ptr(b), ptr(c), _c2(ca), _c2(cb), _c2(cc), out.data_ptr(), self.stream)
        elif not any(cplx):
            _lib.call("hx_combine_abc", nnz, ptr(a), ptr(b), ptr(c), _c2(ca), _c2(cb), _c2(cc), out.data_ptr(), self.stream)
        else:
            raise TypeError("combine_abc: A and C must both be real or both complex")
        return out

    def lowrank_dots(self, lr: LowRank, x, t):
        """t_f = right_f^T x  (the adjoint problem passes a LowRank with the roles swapped)."""
        _lib.call("hx_lowrank_dots", lr.r, lr.rptr.data_ptr(), lr.ridx.data_ptr(), lr.rval.data_ptr(), x.data_ptr(),
                  t.data_ptr(), self.stream)
        return t

    def lowrank_update(self, lr: LowRank, t, coef, y):
        """y += coef * sum_f left_f t_f."""
        _lib.call("hx_lowrank_update", int(lr.lrow.numel()), lr.lrow.data_ptr(), lr.lptr.data_ptr(), lr.lcol.data_ptr(),
                  lr.lval.data_ptr(), t.data_ptr(), _c2(coef), y.data_ptr(), self.stream)
        return y

    # ---- Krylov basis ---------------------------------------------------------------
    def multi_dot(self, V, k, w, out, conj=True):
        """out[j] = <V[j], w>, j<k; V is (m, n) row-major."""
        n = w.numel()
        _lib.call("hx_multi_dot", n, k, V.data_ptr(), V.stride(0), w.data_ptr(), 1 if conj else 0, out.data_ptr(),
                  self.scratch(k).data_ptr(), self.stream)
        return out

    def multi_axpy(self, V, k, h, w, hacc=None, nrm2=None):
        """w -= sum_j h[j] V[j]; hacc += h; nrm2 (device float64 view) = ||w||^2."""
        n = w.numel()
        _lib.call("hx_multi_axpy", n, k, V.data_ptr(), V.stride(0), h.data_ptr(), w.data_ptr(),
                  hacc.data_ptr() if hacc is not None else None, nrm2.data_ptr() if nrm2 is not None else None,
                  self.scratch(1).data_ptr(), self.stream)
        return w

    def scale_copy(self, w, out, nrm2=None, alpha=None):
        """out = w/sqrt(nrm2[0]) (device scalar) or alpha*w."""
        _lib.call("hx_scale_copy", w.numel(), w.data_ptr(), nrm2.data_ptr() if nrm2 is not None else None,
                  _c2(alpha if alpha is not None else 1.0), out.data_ptr(), self.stream)
        return out

    def axpby(self, a, x, b, y):
        """y = a*x + b*y."""
        _lib.call("hx_axpby", y.numel(), _c2(a), x.data_ptr(), _c2(b) if b is not None else None, y.data_ptr(), self.stream)
        return y

    #: restart rotation on the FP64 tensor cores (DMMA, 16 output vectors per pass over V) when more than 8
    #: vectors come out, on the FP64 CUDA cores (8 per pass) otherwise.  Measured on B200
    #: (profiles/r2_rotate_bench_dmma_vs_fma.json, n = 10 M): m=19,k=10 1.09 ms vs 1.69 ms (4.24 vs 2.74 TB/s),
    #: m=24,k=12 1.42 vs 2.08 ms, m=64,k=32 9.46 vs 10.77 ms; m=64,k=8 4.61 vs 3.11 ms (half of the tile is padding).
    ROTATE_TENSOR_CORES = None

    def basis_rotate(self, V, m, Q, kout, Vout, tensor_cores=None):
        """Vout[c] = sum_j Q[c, j] V[j]  (Q: (kout, m) row-major tensor = column-major m x kout)."""
        assert Q.is_contiguous() and Q.shape[1] >= m
        n = V.shape[1]
        tc = tensor_cores if tensor_cores is not None else self.ROTATE_TENSOR_CORES
        if tc is None:
            tc = kout > 8
        _lib.call("hx_basis_rotate_dmma" if tc else "hx_basis_rotate", n, m, kout, V.data_ptr(), V.stride(0), Q.data_ptr(),
                  Q.stride(0), Vout.data_ptr(), Vout.stride(0), self.stream)
        return Vout

    # ---- preconditioner pieces -----------------------------------------------------------
    def diag_inv(self, M: CsrMatrix, out):
        _lib.call("hx_extract_diag_inv", M.n_rows, M.indptr.data_ptr(), M.indices.data_ptr(), M.values.data_ptr(),
                  out.data_ptr(), self.stream)
        return out

    def jacobi_sweep(self, M, dinv, b, xin, xout, omega):
        """xout = xin + omega*dinv*(b - M xin); xin None => xout = omega*dinv*b (matrix unused)."""
        name = "hx_jacobi_sweep_c" if b.dtype == torch.complex64 else "hx_jacobi_sweep"
        if xin is None:
            _lib.call(name, M.n_rows, None, None, None, dinv.data_ptr(), b.data_ptr(), None, xout.data_ptr(),
                      float(omega), 8, self.stream)
        elif getattr(M, "is_sell", False):
            M.jacobi(dinv, b, xin, xout, omega)
        else:
            _lib.call(name, M.n_rows, M.indptr.data_ptr(), M.indices.data_ptr(), M.values.data_ptr(),
                      dinv.data_ptr(), b.data_ptr(), xin.data_ptr(), xout.data_ptr(), float(omega), M.lanes, self.stream)
        return xout

    def amg_tail(self, desc, b):
        """Run the fused tail of the multigrid cycle (amg.TailDesc) on the right-hand side b (complex64)."""
        _lib.call("hx_amg_tail", C.byref(desc), b.data_ptr(), self.stream)

    def dense_inverse(self, A):
        """In-place inverse of a column-major n x n complex matrix; returns info tensor (0 = ok)."""
        n = A.shape[0]
        info = torch.zeros(n + 1, dtype=torch.int32, device=self.device)
        _lib.call("hx_dense_inverse", n, A.data_ptr(), info.data_ptr(), self.stream)
        return info

    def dense_gemv(self, A, x, y):
        _lib.call("hx_dense_gemv", A.shape[0], A.data_ptr(), x.data_ptr(), y.data_ptr(), self.stream)
        return y

    def launch_count(self):
        return int(_lib.call("hx_launch_count"))

    def add_launches(self, n):
        """Account for kernels launched by a CUDA-graph replay (they bypass the C-ABI entry points)."""
        _lib.load().hx_launch_count_add(int(n))

    def reset_launch_count(self):
        _lib.load().hx_launch_count_reset()
