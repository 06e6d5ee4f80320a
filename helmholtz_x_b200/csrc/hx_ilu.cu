// K9 building block: ILU(0) on the CSR pattern with level scheduling -- the
// stand-in for the exact LU (PETSc LU / MUMPS) behind SLEPc's ST sinvert
// (helmholtz_x/eigensolvers.py:49-50,102-103).  Rows of one dependency level are
// independent; one launch per level.
#include "hx_common.cuh"

namespace hx {

__device__ __forceinline__ int ilu_find(const int* __restrict__ indices, int lo, int hi, int col) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (indices[mid] < col) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per row of the level
__global__ void __launch_bounds__(128)
ilu0_factor_level_kernel(int n_rows, const int* __restrict__ rows, const int* __restrict__ indptr,
                         const int* __restrict__ indices, const int* __restrict__ diag, double2* __restrict__ lu) {
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const int i = rows[w];
    const int lo = indptr[i], hi = indptr[i + 1], di = diag[i];
    for (int kk = lo; kk < di; ++kk) {
        const int k = indices[kk];
        const double2 lik = cdiv(lu[kk], lu[diag[k]]);
        __syncwarp();
        if (lane == 0) lu[kk] = lik;
        const int ke = indptr[k + 1];
        for (int jj = diag[k] + 1 + lane; jj < ke; jj += 32) {
            const int j = indices[jj];
            const int p = ilu_find(indices, kk + 1, hi, j);
            if (p < hi && indices[p] == j) {
                const double2 u = lu[jj];
                double2 a = lu[p];
                a.x -= lik.x * u.x - lik.y * u.y;
                a.y -= lik.x * u.y + lik.y * u.x;
                lu[p] = a;
            }
        }
        __syncwarp();
    }
}

__global__ void ilu0_forward_level_kernel(int n_rows, const int* __restrict__ rows, const int* __restrict__ indptr,
                                          const int* __restrict__ indices, const int* __restrict__ diag,
                                          const double2* __restrict__ lu, const double2* __restrict__ b,
                                          double2* __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows) return;
    const int i = rows[t];
    double2 s = b[i];
    const int di = diag[i];
    for (int kk = indptr[i]; kk < di; ++kk) {
        const double2 l = lu[kk], xv = x[indices[kk]];
        s.x -= l.x * xv.x - l.y * xv.y;
        s.y -= l.x * xv.y + l.y * xv.x;
    }
    x[i] = s;
}

__global__ void ilu0_backward_level_kernel(int n_rows, const int* __restrict__ rows, const int* __restrict__ indptr,
                                           const int* __restrict__ indices, const int* __restrict__ diag,
                                           const double2* __restrict__ lu, double2* __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows) return;
    const int i = rows[t];
    double2 s = x[i];
    const int di = diag[i], hi = indptr[i + 1];
    for (int kk = di + 1; kk < hi; ++kk) {
        const double2 u = lu[kk], xv = x[indices[kk]];
        s.x -= u.x * xv.x - u.y * xv.y;
        s.y -= u.x * xv.y + u.y * xv.x;
    }
    x[i] = cdiv(s, lu[di]);
}

}  // namespace hx

using namespace hx;

extern "C" int hx_ilu0_factor(int n, const int32_t* indptr, const int32_t* indices, const int32_t* diag_pos, double* lu,
                              int n_levels, const int32_t* level_ptr_h, const int32_t* level_rows, hx_stream_t stream) {
    (void)n;
    for (int l = 0; l < n_levels; ++l) {
        const int nr = level_ptr_h[l + 1] - level_ptr_h[l];
        if (nr <= 0) continue;
        ilu0_factor_level_kernel<<<ceil_div(nr, 4), 128, 0, (cudaStream_t)stream>>>(nr, level_rows + level_ptr_h[l], indptr, indices,
                                                                                 diag_pos, (double2*)lu);
        int rc = check_launch("ilu0_factor_level_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}

extern "C" int hx_ilu0_solve(int n, const int32_t* indptr, const int32_t* indices, const int32_t* diag_pos, const double* lu,
                             int n_levels, const int32_t* level_ptr_h, const int32_t* level_rows, int n_levels_u,
                             const int32_t* level_ptr_u_h, const int32_t* level_rows_u, const double* b, double* x,
                             hx_stream_t stream) {
    (void)n;
    for (int l = 0; l < n_levels; ++l) {
        const int nr = level_ptr_h[l + 1] - level_ptr_h[l];
        if (nr <= 0) continue;
        ilu0_forward_level_kernel<<<ceil_div(nr, 128), 128, 0, (cudaStream_t)stream>>>(
            nr, level_rows + level_ptr_h[l], indptr, indices, diag_pos, (const double2*)lu, (const double2*)b, (double2*)x);
        int rc = check_launch("ilu0_forward_level_kernel");
        if (rc) return rc;
    }
    for (int l = 0; l < n_levels_u; ++l) {
        const int nr = level_ptr_u_h[l + 1] - level_ptr_u_h[l];
        if (nr <= 0) continue;
        ilu0_backward_level_kernel<<<ceil_div(nr, 128), 128, 0, (cudaStream_t)stream>>>(
            nr, level_rows_u + level_ptr_u_h[l], indptr, indices, diag_pos, (const double2*)lu, (double2*)x);
        int rc = check_launch("ilu0_backward_level_kernel");
        if (rc) return rc;
    }
    return HX_OK;
}
