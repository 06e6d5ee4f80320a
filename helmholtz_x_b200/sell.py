"""SELL-32-sigma storage of a complex CSR matrix (K7 production format): rows are
sorted by length inside windows of `sigma` rows, grouped in slices of 32, and stored
column-major inside each slice so a warp's loads are fully coalesced (one thread per
row, no reduction).  Conversion and SpMV are libhx_b200 kernels; the row permutation is
integer plumbing done with torch."""
import torch

from . import _lib


class SellPattern:
    """Structure of a SELL-32 matrix + the map sell position -> csr position (src)."""

    def __init__(self, be, indptr, indices, n_rows, n_cols, sigma=1024):
        self.be, self.n, self.n_cols = be, n_rows, n_cols
        n = n_rows
        lens = (indptr[1:] - indptr[:-1]).long()
        maxlen = int(lens.max()) if n else 0
        win = torch.arange(n, device=lens.device) // sigma
        key = win * (maxlen + 1) + (maxlen - lens)          # descending length inside each window, stable
        self.row_perm = torch.sort(key, stable=True).indices.to(torch.int32).contiguous()
        self.n_slices = (n + 31) // 32
        widths = be.zeros(max(self.n_slices, 1), dtype=torch.int32)
        _lib.call("hx_sell_slice_widths", n, indptr.data_ptr(), self.row_perm.data_ptr(), self.n_slices, widths.data_ptr(),
                  be.stream)
        self.slice_ptr = be.zeros(self.n_slices + 1, dtype=torch.int64)
        self.slice_ptr[1:] = torch.cumsum(widths[:self.n_slices].long() * 32, 0)
        self.total = int(self.slice_ptr[-1])
        self.cols = be.empty(max(self.total, 1), dtype=torch.int32)
        self.src = be.empty(max(self.total, 1), dtype=torch.int32)
        _lib.call("hx_sell_fill", n, indptr.data_ptr(), indices.data_ptr(), None, self.row_perm.data_ptr(), self.n_slices,
                  self.slice_ptr.data_ptr(), self.cols.data_ptr(), None, self.src.data_ptr(), be.stream)
        self.nnz = int(indices.numel())

    @property
    def padding_ratio(self):
        return float(self.total) / max(self.nnz, 1)

    def values_from_csr(self, csr_vals, out=None, dtype=torch.complex128):
        """SELL-ordered copy of complex128 CSR values (optionally down-converted to complex64)."""
        if csr_vals.dtype == torch.float32:          # real transfer operator of the complex64 cycle
            out = out if out is not None else self.be.empty(max(self.total, 1), dtype=torch.float32)
            _lib.call("hx_sell_gather_s", self.total, self.src.data_ptr(), csr_vals.data_ptr(), out.data_ptr(), self.be.stream)
            return out
        out = out if out is not None else self.be.empty(max(self.total, 1), dtype=dtype)
        name = "hx_sell_gather_c" if out.dtype == torch.complex64 else "hx_sell_gather"
        _lib.call(name, self.total, self.src.data_ptr(), csr_vals.data_ptr(), out.data_ptr(), self.be.stream)
        return out


class SellMatrix:
    is_sell = True

    def __init__(self, pattern: SellPattern, vals, variant=0):
        self.p, self.vals, self.variant = pattern, vals, variant
        self.be = pattern.be
        self.n_rows, self.n_cols, self.nnz = pattern.n, pattern.n_cols, pattern.nnz

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def padding_ratio(self):
        return self.p.padding_ratio

    @classmethod
    def from_csr(cls, be, M, sigma=1024, variant=0):
        p = SellPattern(be, M.indptr, M.indices, M.n_rows, M.n_cols, sigma)
        return cls(p, p.values_from_csr(M.values), variant)

    def spmv(self, x, y, alpha=None, beta=None, y0=None, variant=None):
        from .backend import _c2
        p = self.p
        y0p = None
        if beta is not None:
            y0p = (y0 if y0 is not None else y).data_ptr()
        name = {torch.complex64: "hx_spmv_sell_cc", torch.float32: "hx_spmv_sell_sc"}.get(self.vals.dtype, "hx_spmv_sell_zz")
        _lib.call(name, p.n, p.n_slices, p.slice_ptr.data_ptr(), p.cols.data_ptr(), self.vals.data_ptr(),
                  p.row_perm.data_ptr(), x.data_ptr(), y.data_ptr(), _c2(alpha) if alpha is not None else None,
                  _c2(beta) if beta is not None else None, y0p, self.variant if variant is None else variant, self.be.stream)
        return y

    def jacobi(self, dinv, b, xin, xout, omega):
        p = self.p
        name = "hx_jacobi_sell_c" if self.vals.dtype == torch.complex64 else "hx_jacobi_sell"
        _lib.call(name, p.n, p.n_slices, p.slice_ptr.data_ptr(), p.cols.data_ptr(), self.vals.data_ptr(),
                  p.row_perm.data_ptr(), dinv.data_ptr(), b.data_ptr(), xin.data_ptr(), xout.data_ptr(), float(omega),
                  self.variant, self.be.stream)
        return xout
