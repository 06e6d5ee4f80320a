"""CSR x CSR products for the multigrid set-up through libhx_b200 (hx_spgemm_*)."""
import torch

from . import _lib
from .backend import CsrMatrix

f64 = torch.float64


_longest = {}


class Overflow(Exception):
    """A product row has more distinct columns than the kernel's shared-memory budget."""


SMALL_CAP = 512          # rows of at most this many distinct columns use the 8-warps-per-CTA kernels


def symbolic(be, A: CsrMatrix, B: CsrMatrix):
    """Pattern (indptr, indices) of A*B.  The small size class first (fine levels), the large one on overflow."""
    m = A.n_rows
    row_nnz = be.zeros(max(m, 1), dtype=torch.int32)
    mode = 2
    _lib.call("hx_spgemm_symbolic", m, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(), B.indices.data_ptr(),
              row_nnz.data_ptr(), None, None, mode, be.stream)
    if m and int(row_nnz.min()) < 0:
        mode = 0
        _lib.call("hx_spgemm_symbolic", m, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(), B.indices.data_ptr(),
                  row_nnz.data_ptr(), None, None, mode, be.stream)
    if m and int(row_nnz.min()) < 0:
        raise Overflow()
    indptr = be.zeros(m + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(row_nnz[:m].long(), 0)
    nnz = int(indptr[-1])
    if nnz >= 2 ** 31:
        raise Overflow()
    indptr = indptr.to(torch.int32).contiguous()
    indices = be.empty(max(nnz, 1), dtype=torch.int32)
    _lib.call("hx_spgemm_symbolic", m, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(), B.indices.data_ptr(),
              row_nnz.data_ptr(), indptr.data_ptr(), indices.data_ptr(), mode + 1, be.stream)
    return indptr, indices[:nnz]


def numeric(be, A: CsrMatrix, B: CsrMatrix, indptr, indices, out=None):
    """Values of A*B on the given pattern (float64)."""
    out = out if out is not None else be.empty(max(int(indices.numel()), 1), dtype=f64)
    key = (indptr.data_ptr(), int(indptr.numel()))
    if _longest.get("key") != key:                       # longest row of the pattern (one sync per pattern)
        _longest["key"] = key
        _longest["len"] = int((indptr[1:] - indptr[:-1]).max()) if indptr.numel() > 1 else 0
    name = "hx_spgemm_numeric_small" if _longest["len"] <= SMALL_CAP else "hx_spgemm_numeric"
    _lib.call(name, A.n_rows, A.indptr.data_ptr(), A.indices.data_ptr(), A.values.data_ptr(),
              B.indptr.data_ptr(), B.indices.data_ptr(), B.values.data_ptr(), indptr.data_ptr(), indices.data_ptr(),
              out.data_ptr(), be.stream)
    return out[:indices.numel()]


def multiply(be, A: CsrMatrix, B: CsrMatrix) -> CsrMatrix:
    ip, ix = symbolic(be, A, B)
    return CsrMatrix(A.n_rows, B.n_cols, ip, ix, numeric(be, A, B, ip, ix))


def transpose(M: CsrMatrix) -> CsrMatrix:
    """CSR transpose by a key sort (integer plumbing with torch)."""
    dev = M.indices.device
    rows = torch.repeat_interleave(torch.arange(M.n_rows, device=dev), (M.indptr[1:] - M.indptr[:-1]).long(),
                                   output_size=M.nnz)
    key = M.indices.long() * M.n_rows + rows
    order = torch.sort(key).indices
    counts = torch.bincount(M.indices.long(), minlength=M.n_cols)
    indptr = torch.zeros(M.n_cols + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    return CsrMatrix(M.n_cols, M.n_rows, indptr.to(torch.int32).contiguous(), rows[order].to(torch.int32).contiguous(),
                     M.values[order].contiguous())
