"""SELL-32-sigma storage of a complex CSR matrix (K7 variant): rows are sorted by
length inside windows of `sigma` rows, grouped in slices of 32, and stored column-major
inside each slice so a warp's loads are fully coalesced (one thread per row, no
reduction).  Conversion and SpMV are libhx_b200 kernels; the row permutation is integer
plumbing done with torch."""
import torch

from . import _lib


class SellMatrix:
    def __init__(self, be, n, n_cols, slice_ptr, cols, vals, row_perm, nnz):
        self.be, self.n, self.n_cols = be, n, n_cols
        self.slice_ptr, self.cols, self.vals, self.row_perm, self.nnz = slice_ptr, cols, vals, row_perm, nnz

    @property
    def n_slices(self):
        return self.slice_ptr.numel() - 1

    @property
    def padding_ratio(self):
        return float(self.cols.numel()) / max(self.nnz, 1)

    @classmethod
    def from_csr(cls, be, M, sigma=1024):
        n = M.n_rows
        lens = (M.indptr[1:] - M.indptr[:-1]).long()
        # sort by descending length inside each window of sigma rows (stable)
        win = torch.arange(n, device=lens.device) // sigma
        key = win * (int(lens.max()) + 1 if n else 1) + (int(lens.max()) - lens if n else lens)
        row_perm = torch.sort(key, stable=True).indices.to(torch.int32).contiguous()
        n_slices = (n + 31) // 32
        widths = be.zeros(max(n_slices, 1), dtype=torch.int32)
        _lib.call("hx_sell_slice_widths", n, M.indptr.data_ptr(), row_perm.data_ptr(), n_slices, widths.data_ptr(), be.stream)
        slice_ptr = be.zeros(n_slices + 1, dtype=torch.int64)
        slice_ptr[1:] = torch.cumsum(widths[:n_slices].long() * 32, 0)
        total = int(slice_ptr[-1])
        cols = be.empty(max(total, 1), dtype=torch.int32)
        vals = be.empty(max(total, 1))
        _lib.call("hx_sell_fill", n, M.indptr.data_ptr(), M.indices.data_ptr(), M.values.data_ptr(), row_perm.data_ptr(),
                  n_slices, slice_ptr.data_ptr(), cols.data_ptr(), vals.data_ptr(), be.stream)
        return cls(be, n, M.n_cols, slice_ptr, cols[:total], vals[:total], row_perm, M.nnz)

    def spmv(self, x, y):
        _lib.call("hx_spmv_sell_zz", self.n, self.n_slices, self.slice_ptr.data_ptr(), self.cols.data_ptr(),
                  self.vals.data_ptr(), self.row_perm.data_ptr(), x.data_ptr(), y.data_ptr(), self.be.stream)
        return y
