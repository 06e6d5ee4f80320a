"""Level-scheduled ILU(0) on the CSR pattern (K9 building block named by the north star;
the default inner preconditioner is the SA-AMG cycle in amg.py -- ILU(0) needs
hundreds of dependency levels on these meshes)."""
import numpy as np
import torch

from . import _lib


def _levels(indptr, indices, lower=True):
    """Dependency levels of the lower (or upper) triangular pattern by fixed-point sweeps."""
    n = indptr.numel() - 1
    dev = indptr.device
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (indptr[1:] - indptr[:-1]).long())
    cols = indices.long()
    mask = cols < rows if lower else cols > rows
    r, c = rows[mask], cols[mask]
    lev = torch.zeros(n, dtype=torch.int64, device=dev)
    while True:
        new = torch.zeros(n, dtype=torch.int64, device=dev)
        new.scatter_reduce_(0, r, lev[c] + 1, reduce="amax", include_self=True)
        if torch.equal(new, lev):
            break
        lev = new
    order = torch.sort(lev, stable=True).indices.to(torch.int32).contiguous()
    counts = torch.bincount(lev).cpu().numpy()
    ptr = np.zeros(len(counts) + 1, np.int32)
    ptr[1:] = np.cumsum(counts)
    return len(counts), ptr, order


class ILU0:
    def __init__(self, be, M):
        import ctypes as C
        self.be, self.M = be, M
        n = M.n_rows
        self.diag = be.zeros(n, dtype=torch.int32)
        _lib.call("hx_diag_positions", n, M.indptr.data_ptr(), M.indices.data_ptr(), self.diag.data_ptr(), be.stream)
        self.nl, self.lptr, self.lrows = _levels(M.indptr, M.indices, True)
        self.nu, self.uptr, self.urows = _levels(M.indptr, M.indices, False)
        self.lu = M.values.clone()
        self._c = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.call("hx_ilu0_factor", n, M.indptr.data_ptr(), M.indices.data_ptr(), self.diag.data_ptr(), self.lu.data_ptr(),
                  self.nl, self._c(self.lptr), self.lrows.data_ptr(), be.stream)

    def solve(self, b, x):
        M = self.M
        _lib.call("hx_ilu0_solve", M.n_rows, M.indptr.data_ptr(), M.indices.data_ptr(), self.diag.data_ptr(),
                  self.lu.data_ptr(), self.nl, self._c(self.lptr), self.lrows.data_ptr(), self.nu, self._c(self.uptr),
                  self.urows.data_ptr(), b.data_ptr(), x.data_ptr(), self.be.stream)
        return x
