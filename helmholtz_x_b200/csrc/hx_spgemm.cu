// Sparse x sparse product C = A * B on CSR (real double values) for the multigrid set-up:
// prolongator smoothing S*T and the Galerkin products X*P, R*(X*P).  One warp per row of A.
//  symbolic: the union of the B rows selected by the A row is collected in a shared-memory
//            hash set, compacted and bitonic-sorted (count pass, then fill pass);
//  numeric : for each A entry in order, lanes run over the B row and add a_ik*b_kj into the
//            accumulator slot found by binary search in the sorted C row -- no atomics, fixed
//            order => bitwise reproducible.  The pattern is computed once and reused for the
//            four value sets (A, C, Re B, Im B) that share it.
#include "hx_common.cuh"

namespace hx {

// Two size classes: rows of the fine levels have a few dozen distinct columns -- a 1024-slot table and a
// 512-entry row buffer per warp let 8 warps share a CTA (24 KB) and several CTAs an SM; the dense coarse
// rows need the large class (4096 / 2048, 2 warps per CTA).  The host tries the small class first and
// falls back on overflow.
template <int HASH>
__device__ __forceinline__ bool sg_insert(int* tab, int col) {
    unsigned h = ((unsigned)col * 2654435761u) & (HASH - 1);
    while (true) {
        const int old = atomicCAS(tab + h, -1, col);
        if (old == -1) return true;
        if (old == col) return false;
        h = (h + 1) & (HASH - 1);
    }
}

// mode 0: row_nnz[i] = number of distinct columns (or -1 on overflow)
// mode 1: write the sorted columns at indices_c[indptr_c[i] ...]
template <int kSgWarps, int kSgHash, int kSgCap>
__global__ void __launch_bounds__(kSgWarps * 32)
spgemm_symbolic_kernel(int m, const int* __restrict__ a_ptr, const int* __restrict__ a_idx,
                       const int* __restrict__ b_ptr, const int* __restrict__ b_idx, int* __restrict__ row_nnz,
                       const int* __restrict__ c_ptr, int* __restrict__ c_idx, int mode) {
    __shared__ int tab[kSgWarps][kSgHash];
    __shared__ int list[kSgWarps][kSgCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kSgWarps + warp;
    if (row >= m) return;
    int* t = tab[warp];
    for (int s = lane; s < kSgHash; s += 32) t[s] = -1;
    __syncwarp();
    int cnt = 0;
    bool overflow = false;
    const int as = a_ptr[row], ae = a_ptr[row + 1];
    for (int kk = as; kk < ae; ++kk) {
        const int k = a_idx[kk];
        const int bs = b_ptr[k], be_ = b_ptr[k + 1];
        if (be_ - bs > kSgCap) { overflow = true; break; }
        for (int jj = bs + lane; jj < be_; jj += 32) cnt += sg_insert<kSgHash>(t, b_idx[jj]) ? 1 : 0;
        // stop before the table can fill up (probing would not terminate on a full table)
        int tot = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (tot > kSgCap) { overflow = true; break; }
    }
    int tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (mode == 0) {
        if (lane == 0) row_nnz[row] = overflow ? -1 : tot;
        return;
    }
    if (overflow) return;
    // compact the table into list[] (order irrelevant), then bitonic sort
    int* L = list[warp];
    int base = 0;
    for (int s0 = 0; s0 < kSgHash; s0 += 32) {
        const int v = t[s0 + lane];
        const unsigned mask = __ballot_sync(0xffffffffu, v != -1);
        if (v != -1) L[base + __popc(mask & ((1u << lane) - 1))] = v;
        base += __popc(mask);
    }
    int np2 = 1;
    while (np2 < tot) np2 <<= 1;
    for (int s = tot + lane; s < np2; s += 32) L[s] = 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < np2; i += 32) {
                const int p = i ^ j;
                if (p > i) {
                    const int x = L[i], y = L[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { L[i] = y; L[p] = x; }
                }
            }
            __syncwarp();
        }
    const int cs = c_ptr[row];
    for (int s = lane; s < tot; s += 32) c_idx[cs + s] = L[s];
}

template <int kSgWarps, int kSgCap>
__global__ void __launch_bounds__(kSgWarps * 32)
spgemm_numeric_kernel(int m, const int* __restrict__ a_ptr, const int* __restrict__ a_idx,
                      const double* __restrict__ a_val, const int* __restrict__ b_ptr, const int* __restrict__ b_idx,
                      const double* __restrict__ b_val, const int* __restrict__ c_ptr, const int* __restrict__ c_idx,
                      double* __restrict__ c_val) {
    __shared__ int cols[kSgWarps][kSgCap];
    __shared__ double acc[kSgWarps][kSgCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kSgWarps + warp;
    if (row >= m) return;
    const int cs = c_ptr[row], n_c = c_ptr[row + 1] - cs;
    if (n_c > kSgCap) return;                 // the host picks the class from the longest row; never taken
    int* C = cols[warp];
    double* A = acc[warp];
    for (int s = lane; s < n_c; s += 32) { C[s] = c_idx[cs + s]; A[s] = 0.0; }
    __syncwarp();
    const int as = a_ptr[row], ae = a_ptr[row + 1];
    for (int kk = as; kk < ae; ++kk) {
        const int k = a_idx[kk];
        const double av = a_val[kk];
        const int bs = b_ptr[k], be_ = b_ptr[k + 1];
        for (int jj = bs + lane; jj < be_; jj += 32) {
            const int col = b_idx[jj];
            int lo = 0, hi = n_c;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (C[mid] < col) lo = mid + 1; else hi = mid; }
            A[lo] = fma(av, b_val[jj], A[lo]);      // columns of one B row are distinct: no conflict
        }
        __syncwarp();
    }
    for (int s = lane; s < n_c; s += 32) c_val[cs + s] = A[s];
}

}  // namespace hx

using namespace hx;

extern "C" int hx_spgemm_symbolic(int m, const int32_t* a_ptr, const int32_t* a_idx, const int32_t* b_ptr,
                                  const int32_t* b_idx, int32_t* row_nnz, const int32_t* c_ptr, int32_t* c_idx,
                                  int write_cols, hx_stream_t stream) {
    if (m <= 0) return HX_OK;
    // write_cols: 0 / 1 = count / fill with the large class (rows up to 2048 distinct columns);
    //             2 / 3 = count / fill with the small class (up to 512; -1 in row_nnz on overflow)
    if (write_cols >= 2)
        spgemm_symbolic_kernel<8, 1024, 512><<<ceil_div(m, 8), 8 * 32, 0, (cudaStream_t)stream>>>(
            m, a_ptr, a_idx, b_ptr, b_idx, row_nnz, c_ptr, c_idx, write_cols - 2);
    else
        spgemm_symbolic_kernel<2, 4096, 2048><<<ceil_div(m, 2), 2 * 32, 0, (cudaStream_t)stream>>>(
            m, a_ptr, a_idx, b_ptr, b_idx, row_nnz, c_ptr, c_idx, write_cols);
    return check_launch("spgemm_symbolic_kernel");
}

extern "C" int hx_spgemm_numeric(int m, const int32_t* a_ptr, const int32_t* a_idx, const double* a_val,
                                 const int32_t* b_ptr, const int32_t* b_idx, const double* b_val,
                                 const int32_t* c_ptr, const int32_t* c_idx, double* c_val, hx_stream_t stream) {
    if (m <= 0) return HX_OK;
    spgemm_numeric_kernel<2, 2048><<<ceil_div(m, 2), 2 * 32, 0, (cudaStream_t)stream>>>(
        m, a_ptr, a_idx, a_val, b_ptr, b_idx, b_val, c_ptr, c_idx, c_val);
    return check_launch("spgemm_numeric_kernel");
}

/* same for a pattern whose longest row has at most 512 entries (8 warps per CTA) */
extern "C" int hx_spgemm_numeric_small(int m, const int32_t* a_ptr, const int32_t* a_idx, const double* a_val,
                                       const int32_t* b_ptr, const int32_t* b_idx, const double* b_val,
                                       const int32_t* c_ptr, const int32_t* c_idx, double* c_val, hx_stream_t stream) {
    if (m <= 0) return HX_OK;
    spgemm_numeric_kernel<8, 512><<<ceil_div(m, 8), 8 * 32, 0, (cudaStream_t)stream>>>(
        m, a_ptr, a_idx, a_val, b_ptr, b_idx, b_val, c_ptr, c_idx, c_val);
    return check_launch("spgemm_numeric_kernel<small>");
}
