// K7/K8: complex-double CSR / SELL-32 SpMV, shifted-operator combine, matrix-free
// low-rank flame term, Jacobi sweep.  All HBM-bound: 16-byte value loads streamed
// past L1 (no_allocate) so L1/L2 hold the gathered x entries.
#include "hx_common.cuh"

namespace hx {

// ---- one row's dot product with LANES cooperating threads --------------------
// MV: matrix value type (double2 | double | float2 | float), V: vector type (double2 | float2)
template <int LANES, typename MV, typename V>
__device__ __forceinline__ V row_dot(const int* __restrict__ indices, const MV* __restrict__ vals,
                                     const V* __restrict__ x, int start, int end, int lane) {
    V acc0 = vzero<V>(), acc1 = vzero<V>();
    int k = start + lane;
    for (; k + LANES < end; k += 2 * LANES) {
        const int c0 = ld_stream(indices + k);
        const int c1 = ld_stream(indices + k + LANES);
        const MV v0 = ld_stream(vals + k);
        const MV v1 = ld_stream(vals + k + LANES);
        const V x0 = __ldg(x + c0);
        const V x1 = __ldg(x + c1);
        mac(acc0, v0, x0);
        mac(acc1, v1, x1);
    }
    if (k < end) {
        const int c0 = ld_stream(indices + k);
        const MV v0 = ld_stream(vals + k);
        mac(acc0, v0, __ldg(x + c0));
    }
    acc0 = cadd(acc0, acc1);
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        acc0.x += __shfl_xor_sync(0xffffffffu, acc0.x, o, LANES);
        acc0.y += __shfl_xor_sync(0xffffffffu, acc0.y, o, LANES);
    }
    return acc0;
}

constexpr int kSpmvThreads = 256;

template <int LANES, typename VT, typename V = double2>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_csr_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                const VT* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y,
                V alpha, V beta, const V* y0) {
    const int row = blockIdx.x * (kSpmvThreads / LANES) + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    const bool valid = row < n;   // no early exit: every lane reaches the full-mask shuffles
    const int start = valid ? __ldg(indptr + row) : 0, end = valid ? __ldg(indptr + row + 1) : 0;
    V acc = row_dot<LANES, VT, V>(indices, vals, x, start, end, lane);
    if (valid && lane == 0) {
        V r = cmul(alpha, acc);
        if (y0) r = cadd(r, cmul(beta, y0[row]));
        y[row] = r;
    }
}

template <int LANES, typename V = double2>
__global__ void __launch_bounds__(kSpmvThreads)
jacobi_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
              const V* __restrict__ vals, const V* __restrict__ dinv,
              const V* __restrict__ b, const V* __restrict__ xin, V* __restrict__ xout,
              typename scalar_of<V>::type omega) {
    const int row = blockIdx.x * (kSpmvThreads / LANES) + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    const bool valid = row < n;
    const int start = valid ? __ldg(indptr + row) : 0, end = valid ? __ldg(indptr + row + 1) : 0;
    V acc = row_dot<LANES, V, V>(indices, vals, xin, start, end, lane);
    if (valid && lane == 0) {
        V r = csub(b[row], acc);
        xout[row] = cadd(xin[row], cscale(omega, cmul(dinv[row], r)));
    }
}

template <typename V>
__global__ void jacobi_first_kernel(int n, const V* __restrict__ dinv, const V* __restrict__ b,
                                    V* __restrict__ xout, typename scalar_of<V>::type omega) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xout[i] = cscale(omega, cmul(dinv[i], b[i]));
}

__global__ void diag_inv_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                const double2* __restrict__ vals, double2* __restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 d = make_double2(1.0, 0.0);
    for (int k = indptr[i]; k < indptr[i + 1]; ++k)
        if (indices[k] == i) { d = vals[k]; break; }
    dinv[i] = cdiv(make_double2(1.0, 0.0), d);
}

__global__ void diag_pos_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                int* __restrict__ pos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = -1;
    for (int k = indptr[i]; k < indptr[i + 1]; ++k)
        if (indices[k] == i) { p = k; break; }
    pos[i] = p;
}

// ---- SELL-32: one thread per row, column-major inside a 32-row slice ----------
// MODE 0: y = acc; 1: y = alpha*acc + beta*y0; 2: Jacobi xout = xin + omega*dinv*(b - acc)
// MV: matrix value type (V, or the real scalar for the real-valued transfer operators of the cycle)
template <int UNROLL, int MINB, int MODE, typename V = double2, typename MV = V>
__global__ void __launch_bounds__(256, MINB)
sell_kernel(int n, int n_slices, const long long* __restrict__ slice_ptr, const int* __restrict__ cols,
            const MV* __restrict__ vals, const int* __restrict__ row_perm, const V* __restrict__ x,
            V* __restrict__ y, V alpha, V beta, const V* __restrict__ y0,
            const V* __restrict__ dinv, const V* __restrict__ b, typename scalar_of<V>::type omega) {
    const int slice = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const long long base = slice_ptr[slice];
    const int width = (int)((slice_ptr[slice + 1] - base) >> 5);
    const int* c = cols + base + lane;
    const MV* v = vals + base + lane;
    V acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u] = vzero<V>();
    int j = 0;
    for (; j + UNROLL <= width; j += UNROLL) {
        int cc[UNROLL];
        MV vv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) cc[u] = ld_stream(c + 32 * (j + u));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) vv[u] = ld_stream(v + 32 * (j + u));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) mac(acc[u], vv[u], __ldg(x + cc[u]));
    }
    for (; j < width; ++j) {
        const int c0 = ld_stream(c + 32 * j);
        const MV v0 = ld_stream(v + 32 * j);
        mac(acc[0], v0, __ldg(x + c0));
    }
#pragma unroll
    for (int u = 1; u < UNROLL; ++u) acc[0] = cadd(acc[0], acc[u]);
    const int r = slice * 32 + lane;
    if (r >= n) return;
    const int row = row_perm[r];
    if (MODE == 0) y[row] = acc[0];
    else if (MODE == 1) {
        V res = cmul(alpha, acc[0]);
        if (y0) res = cadd(res, cmul(beta, y0[row]));
        y[row] = res;
    } else {
        const V res = csub(b[row], acc[0]);
        y[row] = cadd(x[row], cscale(omega, cmul(dinv[row], res)));
    }
}

__global__ void sell_widths_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ row_perm,
                                   int n_slices, int* __restrict__ widths) {
    const int slice = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const int r = slice * 32 + lane;
    int len = 0;
    if (r < n) { const int row = row_perm[r]; len = indptr[row + 1] - indptr[row]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) widths[slice] = len;
}

__global__ void sell_fill_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                 const double2* __restrict__ vals, const int* __restrict__ row_perm, int n_slices,
                                 const long long* __restrict__ slice_ptr, int* __restrict__ cols,
                                 double2* __restrict__ svals, int* __restrict__ src) {
    const int slice = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const long long base = slice_ptr[slice];
    const int width = (int)((slice_ptr[slice + 1] - base) >> 5);
    const int r = slice * 32 + lane;
    int start = 0, len = 0, row = 0;
    if (r < n) { row = row_perm[r]; start = indptr[row]; len = indptr[row + 1] - start; }
    for (int j = 0; j < width; ++j) {
        const long long p = base + 32LL * j + lane;
        if (j < len) {
            cols[p] = indices[start + j];
            if (svals) svals[p] = vals[start + j];
            if (src) src[p] = start + j;
        } else {                                  // padding: value 0, column = the row's first column (in
            cols[p] = len > 0 ? indices[start] : 0;   // bounds for rectangular matrices too, e.g. prolongators)
            if (svals) svals[p] = make_double2(0.0, 0.0);
            if (src) src[p] = -1;
        }
    }
}

template <typename V>
__global__ void sell_gather_kernel(long long total, const int* __restrict__ src, const double2* __restrict__ csr_vals,
                                   V* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int s = ld_stream(src + i);
        out[i] = (s >= 0) ? from_c128<V>(csr_vals[s]) : vzero<V>();
    }
}

__global__ void sell_gather_real_kernel(long long total, const int* __restrict__ src, const float* __restrict__ csr_vals,
                                        float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int s = ld_stream(src + i);
        out[i] = (s >= 0) ? csr_vals[s] : 0.f;
    }
}

// ---- K8: P(sigma) values on the shared pattern ------------------------------------
__global__ void combine_abc_kernel(long long nnz, const double* __restrict__ a, const double2* __restrict__ b,
                                   const double* __restrict__ c, double2 ca, double2 cb, double2 cc,
                                   double2* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        double2 r = make_double2(0.0, 0.0);
        if (a) r = cscale(ld_stream(a + i), ca);
        if (c) r = cadd(r, cscale(ld_stream(c + i), cc));
        if (b) r = cadd(r, cmul(cb, ld_stream(b + i)));
        out[i] = r;
    }
}

// all three operands complex (Bloch-reduced A and C are Hermitian, not real)
__global__ void combine_zzz_kernel(long long nnz, const double2* __restrict__ a, const double2* __restrict__ b,
                                   const double2* __restrict__ c, double2 ca, double2 cb, double2 cc,
                                   double2* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        double2 r = make_double2(0.0, 0.0);
        if (a) r = cmul(ca, ld_stream(a + i));
        if (c) r = cadd(r, cmul(cc, ld_stream(c + i)));
        if (b) r = cadd(r, cmul(cb, ld_stream(b + i)));
        out[i] = r;
    }
}

// ---- matrix-free rank-r flame term ------------------------------------------------
// one warp per flame f: t_f = r_f^T x   (sparse r_f, fixed reduction order)
__global__ void lowrank_dots_kernel(int r, const int* __restrict__ rptr, const int* __restrict__ ridx,
                                    const double* __restrict__ rval, const double2* __restrict__ x,
                                    double2* __restrict__ t) {
    const int f = blockIdx.x;
    const int s = rptr[f], e = rptr[f + 1];
    double2 acc = make_double2(0.0, 0.0);
    for (int k = s + threadIdx.x; k < e; k += blockDim.x) rfma(acc, rval[k], x[ridx[k]]);
    __shared__ double2 sm[32];
    acc = warp_sum(acc);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = acc;
    __syncthreads();
    if (w == 0) {
        double2 v = (l < (blockDim.x >> 5)) ? sm[l] : make_double2(0.0, 0.0);
        v = warp_sum(v);
        if (l == 0) t[f] = v;
    }
}

// y[lrow[i]] += coef * sum_k lval[k] t[lcol[k]]  -- rows of the union support of all left vectors
__global__ void lowrank_update_kernel(int nrows, const int* __restrict__ lrow, const int* __restrict__ lptr,
                                      const int* __restrict__ lcol, const double* __restrict__ lval,
                                      const double2* __restrict__ t, double2 coef, double2* __restrict__ y) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int k = lptr[i]; k < lptr[i + 1]; ++k) rfma(acc, lval[k], t[lcol[k]]);
    const int row = lrow[i];
    y[row] = cadd(y[row], cmul(coef, acc));
}

template <typename VT, typename V = double2>
static int launch_spmv(int n, const int* indptr, const int* indices, const VT* vals, const V* x, V* y,
                       V alpha, V beta, const V* y0, int lanes, cudaStream_t st) {
    if (n <= 0) return HX_OK;
#define HX_SPMV_CASE(L)                                                                              \
    case L: {                                                                                        \
        const int rows_per_block = kSpmvThreads / L;                                                 \
        spmv_csr_kernel<L, VT, V><<<ceil_div(n, rows_per_block), kSpmvThreads, 0, st>>>(n, indptr, indices, vals, x, y, \
                                                                                   alpha, beta, y0); \
        break;                                                                                       \
    }
    switch (lanes) {
        HX_SPMV_CASE(2) HX_SPMV_CASE(4) HX_SPMV_CASE(8) HX_SPMV_CASE(16) HX_SPMV_CASE(32)
        default: return fail(HX_ERR_ARG, "spmv: lanes must be 2,4,8,16,32%s%s");
    }
#undef HX_SPMV_CASE
    return check_launch("spmv_csr_kernel");
}

}  // namespace hx

using namespace hx;

// the caller passes lanes=0 to let the library choose; it then needs nnz, which
// lives on the device.  To stay sync-free the Python shim passes an explicit lane
// count (it knows nnz); lanes=0 falls back to 8 (P1 tets: ~13 nnz/row).
extern "C" int hx_spmv_zz(int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                          const double* x, double* y, const double* alpha_h, const double* beta_h,
                          const double* y0, int lanes, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    if (!indptr || !indices || !vals || !x || !y) return fail(HX_ERR_ARG, "hx_spmv_zz: null pointer%s%s");
    double2 alpha = alpha_h ? h2c(alpha_h) : make_double2(1.0, 0.0);
    double2 beta = beta_h ? h2c(beta_h) : make_double2(1.0, 0.0);
    if (lanes == 0) lanes = 8;
    return launch_spmv<double2, double2>(n, indptr, indices, (const double2*)vals, (const double2*)x, (double2*)y, alpha, beta,
                                (const double2*)y0, lanes, (cudaStream_t)stream);
}

extern "C" int hx_spmv_dz(int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                          const double* x, double* y, const double* alpha_h, const double* beta_h,
                          const double* y0, int lanes, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    if (!indptr || !indices || !vals || !x || !y) return fail(HX_ERR_ARG, "hx_spmv_dz: null pointer%s%s");
    double2 alpha = alpha_h ? h2c(alpha_h) : make_double2(1.0, 0.0);
    double2 beta = beta_h ? h2c(beta_h) : make_double2(1.0, 0.0);
    if (lanes == 0) lanes = 8;
    return launch_spmv<double, double2>(n, indptr, indices, vals, (const double2*)x, (double2*)y, alpha, beta,
                               (const double2*)y0, lanes, (cudaStream_t)stream);
}

namespace hx {
template <int MODE, typename V = double2, typename MV = V>
static int launch_sell(int variant, int n, int n_slices, const long long* sp, const int* cols, const MV* vals,
                       const int* perm, const V* x, V* y, V alpha, V beta, const V* y0,
                       const V* dinv, const V* b, typename scalar_of<V>::type omega, cudaStream_t st) {
    const int warps = 8;
    const int grid = ceil_div(n_slices, warps);
#define HX_SELL(U, MB)                                                                                          \
    sell_kernel<U, MB, MODE, V, MV><<<grid, warps * 32, 0, st>>>(n, n_slices, sp, cols, vals, perm, x, y, alpha, beta, y0, dinv, b, omega)
    switch (variant) {
        case 1: HX_SELL(4, 4); break;
        case 2: HX_SELL(4, 6); break;
        case 3: HX_SELL(2, 8); break;
        case 4: HX_SELL(4, 8); break;
        case 5: HX_SELL(8, 4); break;
        case 0: default: HX_SELL(4, 4); break;   // best on B200: 6.0 TB/s at 1M and 10M DoF (profiles/README.md)
    }
#undef HX_SELL
    return check_launch("sell_kernel");
}
}  // namespace hx

extern "C" int hx_spmv_sell_zz(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols, const double* vals,
                               const int32_t* row_perm, const double* x, double* y, const double* alpha_h,
                               const double* beta_h, const double* y0, int variant, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const double2 one = make_double2(1.0, 0.0);
    if (!alpha_h && !y0)
        return launch_sell<0, double2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const double2*)vals, row_perm,
                              (const double2*)x, (double2*)y, one, one, nullptr, nullptr, nullptr, 0.0, (cudaStream_t)stream);
    return launch_sell<1, double2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const double2*)vals, row_perm,
                          (const double2*)x, (double2*)y, alpha_h ? h2c(alpha_h) : one, beta_h ? h2c(beta_h) : one,
                          (const double2*)y0, nullptr, nullptr, 0.0, (cudaStream_t)stream);
}

extern "C" int hx_jacobi_sell(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols, const double* vals,
                              const int32_t* row_perm, const double* dinv, const double* b, const double* xin,
                              double* xout, double omega, int variant, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const double2 one = make_double2(1.0, 0.0);
    return launch_sell<2, double2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const double2*)vals, row_perm,
                          (const double2*)xin, (double2*)xout, one, one, nullptr, (const double2*)dinv, (const double2*)b,
                          omega, (cudaStream_t)stream);
}

extern "C" int hx_sell_slice_widths(int n, const int32_t* indptr, const int32_t* row_perm, int n_slices,
                                    int32_t* widths, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    sell_widths_kernel<<<ceil_div(n_slices, 8), 256, 0, (cudaStream_t)stream>>>(n, indptr, row_perm, n_slices, widths);
    return check_launch("sell_widths_kernel");
}

extern "C" int hx_sell_fill(int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                            const int32_t* row_perm, int n_slices, const int64_t* slice_ptr, int32_t* cols,
                            double* svals, int32_t* src, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    sell_fill_kernel<<<ceil_div(n_slices, 8), 256, 0, (cudaStream_t)stream>>>(
        n, indptr, indices, (const double2*)vals, row_perm, n_slices, (const long long*)slice_ptr, cols, (double2*)svals, src);
    return check_launch("sell_fill_kernel");
}

extern "C" int hx_sell_gather(int64_t total, const int32_t* src, const double* csr_vals, double* sell_vals,
                              hx_stream_t stream) {
    if (total <= 0) return HX_OK;
    long long blocks = ceil_div<long long>(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    sell_gather_kernel<double2><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, src, (const double2*)csr_vals, (double2*)sell_vals);
    return check_launch("sell_gather_kernel");
}

// ---- complex64 variants (mixed-precision AMG cycle) ------------------------------------------
static inline float2 h2cf(const double* p) { return make_float2((float)p[0], (float)p[1]); }

extern "C" int hx_sell_gather_c(int64_t total, const int32_t* src, const double* csr_vals, float* sell_vals,
                                hx_stream_t stream) {
    if (total <= 0) return HX_OK;
    long long blocks = ceil_div<long long>(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    sell_gather_kernel<float2><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, src, (const double2*)csr_vals, (float2*)sell_vals);
    return check_launch("sell_gather_kernel<float2>");
}

extern "C" int hx_spmv_cc(int n, const int32_t* indptr, const int32_t* indices, const float* vals, const float* x,
                          float* y, const double* alpha_h, const double* beta_h, const float* y0, int lanes,
                          hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const float2 one = make_float2(1.f, 0.f);
    if (lanes == 0) lanes = 8;
    return launch_spmv<float2, float2>(n, indptr, indices, (const float2*)vals, (const float2*)x, (float2*)y,
                                       alpha_h ? h2cf(alpha_h) : one, beta_h ? h2cf(beta_h) : one, (const float2*)y0, lanes,
                                       (cudaStream_t)stream);
}

extern "C" int hx_spmv_sc(int n, const int32_t* indptr, const int32_t* indices, const float* vals, const float* x,
                          float* y, const double* alpha_h, const double* beta_h, const float* y0, int lanes,
                          hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const float2 one = make_float2(1.f, 0.f);
    if (lanes == 0) lanes = 8;
    return launch_spmv<float, float2>(n, indptr, indices, vals, (const float2*)x, (float2*)y,
                                      alpha_h ? h2cf(alpha_h) : one, beta_h ? h2cf(beta_h) : one, (const float2*)y0, lanes,
                                      (cudaStream_t)stream);
}

extern "C" int hx_spmv_sell_cc(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols, const float* vals,
                               const int32_t* row_perm, const float* x, float* y, const double* alpha_h,
                               const double* beta_h, const float* y0, int variant, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const float2 one = make_float2(1.f, 0.f);
    if (!alpha_h && !y0)
        return launch_sell<0, float2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const float2*)vals, row_perm,
                                      (const float2*)x, (float2*)y, one, one, nullptr, nullptr, nullptr, 0.f, (cudaStream_t)stream);
    return launch_sell<1, float2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const float2*)vals, row_perm,
                                  (const float2*)x, (float2*)y, alpha_h ? h2cf(alpha_h) : one, beta_h ? h2cf(beta_h) : one,
                                  (const float2*)y0, nullptr, nullptr, 0.f, (cudaStream_t)stream);
}

extern "C" int hx_sell_gather_s(int64_t total, const int32_t* src, const float* csr_vals, float* sell_vals, hx_stream_t stream) {
    if (total <= 0) return HX_OK;
    long long blocks = ceil_div<long long>(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    sell_gather_real_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, src, csr_vals, sell_vals);
    return check_launch("sell_gather_real_kernel");
}

extern "C" int hx_spmv_sell_sc(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols, const float* vals,
                               const int32_t* row_perm, const float* x, float* y, const double* alpha_h,
                               const double* beta_h, const float* y0, int variant, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const float2 one = make_float2(1.f, 0.f);
    if (!alpha_h && !y0)
        return launch_sell<0, float2, float>(variant, n, n_slices, (const long long*)slice_ptr, cols, vals, row_perm,
                                             (const float2*)x, (float2*)y, one, one, nullptr, nullptr, nullptr, 0.f, (cudaStream_t)stream);
    return launch_sell<1, float2, float>(variant, n, n_slices, (const long long*)slice_ptr, cols, vals, row_perm,
                                         (const float2*)x, (float2*)y, alpha_h ? h2cf(alpha_h) : one, beta_h ? h2cf(beta_h) : one,
                                         (const float2*)y0, nullptr, nullptr, 0.f, (cudaStream_t)stream);
}

extern "C" int hx_jacobi_sell_c(int n, int n_slices, const int64_t* slice_ptr, const int32_t* cols, const float* vals,
                                const int32_t* row_perm, const float* dinv, const float* b, const float* xin,
                                float* xout, double omega, int variant, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const float2 one = make_float2(1.f, 0.f);
    return launch_sell<2, float2>(variant, n, n_slices, (const long long*)slice_ptr, cols, (const float2*)vals, row_perm,
                                  (const float2*)xin, (float2*)xout, one, one, nullptr, (const float2*)dinv, (const float2*)b,
                                  (float)omega, (cudaStream_t)stream);
}

extern "C" int hx_jacobi_sweep_c(int n, const int32_t* indptr, const int32_t* indices, const float* vals,
                                 const float* dinv, const float* b, const float* xin, float* xout, double omega,
                                 int lanes, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!xin) {
        jacobi_first_kernel<float2><<<ceil_div(n, 256), 256, 0, st>>>(n, (const float2*)dinv, (const float2*)b, (float2*)xout, (float)omega);
        return check_launch("jacobi_first_kernel<float2>");
    }
    if (lanes == 0) lanes = 8;
#define HX_JACF_CASE(L)                                                                                   \
    case L:                                                                                               \
        jacobi_kernel<L, float2><<<ceil_div(n, kSpmvThreads / L), kSpmvThreads, 0, st>>>(                 \
            n, indptr, indices, (const float2*)vals, (const float2*)dinv, (const float2*)b, (const float2*)xin, \
            (float2*)xout, (float)omega);                                                                 \
        break;
    switch (lanes) {
        HX_JACF_CASE(2) HX_JACF_CASE(4) HX_JACF_CASE(8) HX_JACF_CASE(16) HX_JACF_CASE(32)
        default: return fail(HX_ERR_ARG, "jacobi: lanes must be 2,4,8,16,32%s%s");
    }
#undef HX_JACF_CASE
    return check_launch("jacobi_kernel<float2>");
}

extern "C" int hx_combine_abc(int64_t nnz, const double* a, const double* b, const double* c, const double* ca_h,
                              const double* cb_h, const double* cc_h, double* out, hx_stream_t stream) {
    if (nnz <= 0) return HX_OK;
    const double2 z = make_double2(0.0, 0.0);
    long long blocks = ceil_div<long long>(nnz, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    combine_abc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        nnz, a, (const double2*)b, c, ca_h ? h2c(ca_h) : z, cb_h ? h2c(cb_h) : z, cc_h ? h2c(cc_h) : z, (double2*)out);
    return check_launch("combine_abc_kernel");
}

extern "C" int hx_combine_zzz(int64_t nnz, const double* a, const double* b, const double* c, const double* ca_h,
                              const double* cb_h, const double* cc_h, double* out, hx_stream_t stream) {
    if (nnz <= 0) return HX_OK;
    const double2 z = make_double2(0.0, 0.0);
    long long blocks = ceil_div<long long>(nnz, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    combine_zzz_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        nnz, (const double2*)a, (const double2*)b, (const double2*)c, ca_h ? h2c(ca_h) : z, cb_h ? h2c(cb_h) : z,
        cc_h ? h2c(cc_h) : z, (double2*)out);
    return check_launch("combine_zzz_kernel");
}

extern "C" int hx_lowrank_dots(int r, const int32_t* rptr, const int32_t* ridx, const double* rval, const double* x,
                               double* t, hx_stream_t stream) {
    if (r <= 0) return HX_OK;
    lowrank_dots_kernel<<<r, 256, 0, (cudaStream_t)stream>>>(r, rptr, ridx, rval, (const double2*)x, (double2*)t);
    return check_launch("lowrank_dots_kernel");
}

extern "C" int hx_lowrank_update(int nrows, const int32_t* lrow, const int32_t* lptr, const int32_t* lcol,
                                 const double* lval, const double* t, const double* coef_h, double* y,
                                 hx_stream_t stream) {
    if (nrows <= 0) return HX_OK;
    lowrank_update_kernel<<<ceil_div(nrows, 256), 256, 0, (cudaStream_t)stream>>>(
        nrows, lrow, lptr, lcol, lval, (const double2*)t, h2c(coef_h), (double2*)y);
    return check_launch("lowrank_update_kernel");
}

extern "C" int hx_jacobi_sweep(int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                               const double* dinv, const double* b, const double* xin, double* xout, double omega,
                               int lanes, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!xin) {
        jacobi_first_kernel<double2><<<ceil_div(n, 256), 256, 0, st>>>(n, (const double2*)dinv, (const double2*)b, (double2*)xout, omega);
        return check_launch("jacobi_first_kernel");
    }
    if (lanes == 0) lanes = 8;
#define HX_JAC_CASE(L)                                                                                    \
    case L:                                                                                               \
        jacobi_kernel<L, double2><<<ceil_div(n, kSpmvThreads / L), kSpmvThreads, 0, st>>>(                \
            n, indptr, indices, (const double2*)vals, (const double2*)dinv, (const double2*)b, (const double2*)xin, \
            (double2*)xout, omega);                                                                       \
        break;
    switch (lanes) {
        HX_JAC_CASE(2) HX_JAC_CASE(4) HX_JAC_CASE(8) HX_JAC_CASE(16) HX_JAC_CASE(32)
        default: return fail(HX_ERR_ARG, "jacobi: lanes must be 2,4,8,16,32%s%s");
    }
#undef HX_JAC_CASE
    return check_launch("jacobi_kernel");
}

extern "C" int hx_extract_diag_inv(int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                                   double* dinv, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    diag_inv_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, (const double2*)vals, (double2*)dinv);
    return check_launch("diag_inv_kernel");
}

extern "C" int hx_diag_positions(int n, const int32_t* indptr, const int32_t* indices, int32_t* diag_pos,
                                 hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    diag_pos_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, diag_pos);
    return check_launch("diag_pos_kernel");
}
