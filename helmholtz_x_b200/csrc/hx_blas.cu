// K10/K11: Krylov basis kernels (block classical Gram-Schmidt pieces, restart
// rotation) and the dense coarse-level solve.  All reductions are two-stage with a
// fixed summation tree (per-block partials, then the last block to finish sums
// them in block order), so results are bitwise run-to-run reproducible.
#include "hx_common.cuh"

namespace hx {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = kNumSMs * 4;   // persistent grid: 4 CTAs per SM
constexpr int kDotTile = 8;                  // basis vectors per pass

struct ScratchHeader { unsigned int counter; unsigned int pad[3]; };

__device__ __forceinline__ bool last_block_done(unsigned int* counter, unsigned int total) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned int prev = atomicAdd(counter, 1u);
        is_last = (prev == total - 1);
    }
    __syncthreads();
    return is_last;
}

// out[j] = sum_i op(V[j][i]) * w[i];  grid = (G, ceil(k/kDotTile))
template <bool CONJ>
__global__ void __launch_bounds__(kRedThreads)
multi_dot_kernel(long long n, int k, const double2* __restrict__ V, long long ld, const double2* __restrict__ w,
                 double2* __restrict__ out, ScratchHeader* hdr, double2* __restrict__ partial) {
    const int j0 = blockIdx.y * kDotTile;
    const int kt = min(kDotTile, k - j0);
    double2 acc[kDotTile];
#pragma unroll
    for (int t = 0; t < kDotTile; ++t) acc[t] = make_double2(0.0, 0.0);
    const long long stride = (long long)gridDim.x * kRedThreads;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
        const double2 wi = __ldg(w + i);
#pragma unroll
        for (int t = 0; t < kDotTile; ++t) {
            if (t < kt) {
                const double2 v = ld_stream(V + (long long)(j0 + t) * ld + i);
                if (CONJ) cfmac(acc[t], v, wi); else cfma(acc[t], v, wi);
            }
        }
    }
    __shared__ double2 sm[kDotTile][kRedThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < kDotTile; ++t) {
        double2 v = warp_sum(acc[t]);
        if (lane == 0) sm[t][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kDotTile) {
        double2 s = make_double2(0.0, 0.0);
        for (int wv = 0; wv < kRedThreads / 32; ++wv) s = cadd(s, sm[threadIdx.x][wv]);
        if ((int)threadIdx.x < kt) partial[(long long)blockIdx.x * k + j0 + threadIdx.x] = s;
    }
    if (last_block_done(&hdr->counter, gridDim.x * gridDim.y)) {
        // one warp per output: lanes stride over the per-block partials, then a butterfly sum --
        // a fixed tree, so the result does not depend on which block finished last
        for (int j = warp; j < k; j += kRedThreads / 32) {
            double2 s = make_double2(0.0, 0.0);
            for (unsigned int b = lane; b < gridDim.x; b += 32) s = cadd(s, partial[(long long)b * k + j]);
            s = warp_sum(s);
            if (lane == 0) out[j] = s;
        }
        if (threadIdx.x == 0) hdr->counter = 0;
    }
}

// w -= sum_j h[j] V[j]; optional hacc += h; optional nrm2 = ||w||^2
__global__ void __launch_bounds__(kRedThreads)
multi_axpy_kernel(long long n, int k, const double2* __restrict__ V, long long ld, const double2* __restrict__ h,
                  double2* __restrict__ w, double2* hacc, double* nrm2_out, ScratchHeader* hdr,
                  double* __restrict__ partial) {
    extern __shared__ double2 hs[];
    for (int j = threadIdx.x; j < k; j += kRedThreads) hs[j] = h[j];
    __syncthreads();
    double nrm = 0.0;
    const long long stride = (long long)gridDim.x * kRedThreads;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
        double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
        int j = 0;
        for (; j + 1 < k; j += 2) {
            const double2 v0 = ld_stream(V + (long long)j * ld + i);
            const double2 v1 = ld_stream(V + (long long)(j + 1) * ld + i);
            cfma(acc0, hs[j], v0);
            cfma(acc1, hs[j + 1], v1);
        }
        if (j < k) cfma(acc0, hs[j], ld_stream(V + (long long)j * ld + i));
        double2 wi = w[i];
        wi = csub(wi, cadd(acc0, acc1));
        w[i] = wi;
        nrm = fma(wi.x, wi.x, nrm);
        nrm = fma(wi.y, wi.y, nrm);
    }
    if (blockIdx.x == 0 && hacc)
        for (int j = threadIdx.x; j < k; j += kRedThreads) hacc[j] = cadd(hacc[j], hs[j]);
    if (!nrm2_out) return;
    __shared__ double smn[kRedThreads / 32];
    nrm = warp_sum(nrm);
    if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int wv = 0; wv < kRedThreads / 32; ++wv) s += smn[wv];
        partial[blockIdx.x] = s;
    }
    if (last_block_done(&hdr->counter, gridDim.x)) {
        double s = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += kRedThreads) s += partial[b];
        s = warp_sum(s);
        __syncthreads();                       // smn is reused
        if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int wv = 0; wv < kRedThreads / 32; ++wv) t += smn[wv];
            *nrm2_out = t;
            hdr->counter = 0;
        }
    }
}

__global__ void scale_copy_kernel(long long n, const double2* __restrict__ w, const double* nrm2, double2 alpha,
                                  double2* __restrict__ out) {
    double2 a = alpha;
    if (nrm2) a = make_double2(1.0 / sqrt(*nrm2), 0.0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = cmul(a, w[i]);
}

__global__ void axpby_kernel(long long n, double2 a, const double2* __restrict__ x, double2 b, double2* __restrict__ y,
                             bool use_y) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 r = cmul(a, x[i]);
        if (use_y) r = cadd(r, cmul(b, y[i]));
        y[i] = r;
    }
}

// Vout[c][i] = sum_j Q[j + c*ldq] * V[j][i]; grid = (G, ceil(kout/kRotTile))
constexpr int kRotTile = 8;
__global__ void __launch_bounds__(kRedThreads)
basis_rotate_kernel(long long n, int m, int kout, const double2* __restrict__ V, long long ld,
                    const double2* __restrict__ Q, int ldq, double2* __restrict__ Vout, long long ldout) {
    extern __shared__ double2 qs[];   // m x kRotTile
    const int c0 = blockIdx.y * kRotTile;
    const int ct = min(kRotTile, kout - c0);
    for (int e = threadIdx.x; e < m * kRotTile; e += kRedThreads) {
        const int j = e / kRotTile, t = e % kRotTile;
        qs[e] = (t < ct) ? Q[j + (long long)(c0 + t) * ldq] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * kRedThreads;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride) {
        double2 acc[kRotTile];
#pragma unroll
        for (int t = 0; t < kRotTile; ++t) acc[t] = make_double2(0.0, 0.0);
        for (int j = 0; j < m; ++j) {
            const double2 v = ld_stream(V + (long long)j * ld + i);
#pragma unroll
            for (int t = 0; t < kRotTile; ++t) cfma(acc[t], qs[j * kRotTile + t], v);
        }
#pragma unroll
        for (int t = 0; t < kRotTile; ++t)
            if (t < ct) Vout[(long long)(c0 + t) * ldout + i] = acc[t];
    }
}

// ---- the same contraction on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64) --------------------
// Vout (kout x n) = Q^T (kout x m) * V (m x n), complex, as four real m8n8k4 products per k-step:
// Cr += Qr*Vr + (-Qi)*Vi, Ci += Qr*Vi + Qi*Vr.  A warp owns 16 output vectors (two m8 tiles) and walks
// the vector index in chunks of 8; the V fragment of a k-step is four rows x 8 consecutive complex
// entries (4 x 128 B per load instruction).  Q^T lives in shared memory (split re / -im / im).
// Per loaded V entry: 16 complex FMAs on the tensor pipe instead of 8 on the FP64 CUDA cores, and half
// the re-reads of V for kout > 8.
constexpr int kMmaWarps = 8;
constexpr int kMmaTileM = 16;

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kMmaWarps * 32)
basis_rotate_dmma_kernel(long long n, int m, int kout, const double2* __restrict__ V, long long ld,
                         const double2* __restrict__ Q, int ldq, double2* __restrict__ Vout, long long ldout) {
    extern __shared__ double qsm[];               // [3][kMmaTileM][m4]: re, -im, im of Q^T (zero padded)
    const int m4 = (m + 3) & ~3;
    const int c0 = blockIdx.y * kMmaTileM;
    double* qr = qsm;
    double* qni = qsm + kMmaTileM * m4;
    double* qi = qsm + 2 * kMmaTileM * m4;
    for (int e = threadIdx.x; e < kMmaTileM * m4; e += blockDim.x) {
        const int c = e / m4, j = e % m4;
        double2 q = make_double2(0.0, 0.0);
        if (c0 + c < kout && j < m) q = Q[j + (long long)(c0 + c) * ldq];
        qr[e] = q.x; qni[e] = -q.y; qi[e] = q.y;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ar = lane >> 2, ak = lane & 3;      // A fragment: row (output vector), k
    const long long tiles = (n + 7) / 8;
    for (long long t = (long long)blockIdx.x * kMmaWarps + warp; t < tiles; t += (long long)gridDim.x * kMmaWarps) {
        const long long i0 = t * 8;
        const long long ib = i0 + (lane >> 2);    // B fragment: column (vector entry) of this lane
        const bool inb = ib < n;
        double cr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, ci[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int j0 = 0; j0 < m4; j0 += 4) {
            const int j = j0 + ak;
            double2 v = make_double2(0.0, 0.0);
            if (inb && j < m) v = ld_stream(V + (long long)j * ld + ib);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int qe = (h * 8 + ar) * m4 + j;
                const double a_r = qr[qe], a_ni = qni[qe], a_i = qi[qe];
                dmma_8x8x4(cr[h][0], cr[h][1], a_r, v.x);
                dmma_8x8x4(cr[h][0], cr[h][1], a_ni, v.y);
                dmma_8x8x4(ci[h][0], ci[h][1], a_r, v.y);
                dmma_8x8x4(ci[h][0], ci[h][1], a_i, v.x);
            }
        }
        // C fragment: row = lane>>2 (output vector), columns 2*(lane&3) + {0,1}
        const long long ic = i0 + 2 * (lane & 3);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + h * 8 + (lane >> 2);
            if (c < kout) {
                double2* o = Vout + (long long)c * ldout + ic;
                if (ic < n) o[0] = make_double2(cr[h][0], ci[h][0]);
                if (ic + 1 < n) o[1] = make_double2(cr[h][1], ci[h][1]);
            }
        }
    }
}

// ---- dense coarse solve: in-place Gauss-Jordan inverse, partial pivoting, one CTA -----
constexpr int kGJThreads = 1024;
__global__ void __launch_bounds__(kGJThreads)
dense_inverse_kernel(int n, double2* __restrict__ A, int* __restrict__ info, int* __restrict__ piv) {
    extern __shared__ double2 sh[];          // rowk[n], colk[n]
    double2* rowk = sh;
    double2* colk = sh + n;
    __shared__ double red_val[kGJThreads / 32];
    __shared__ int red_idx[kGJThreads / 32];
    __shared__ int pivot_row;
    const int tid = threadIdx.x;
    if (tid == 0) *info = 0;
    for (int k = 0; k < n; ++k) {
        // pivot search in column k, rows >= k
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < n; i += kGJThreads) {
            const double2 v = A[i + (long long)k * n];
            const double a = v.x * v.x + v.y * v.y;
            if (a > best) { best = a; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if ((tid & 31) == 0) { red_val[tid >> 5] = best; red_idx[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = red_val[0]; int ix = red_idx[0];
            for (int w = 1; w < kGJThreads / 32; ++w)
                if (red_val[w] > b || (red_val[w] == b && red_idx[w] < ix)) { b = red_val[w]; ix = red_idx[w]; }
            pivot_row = ix;
            piv[k] = ix;
            if (!(b > 0.0)) *info = k + 1;
        }
        __syncthreads();
        const int p = pivot_row;
        if (p != k)
            for (int j = tid; j < n; j += kGJThreads) {
                const double2 t = A[k + (long long)j * n];
                A[k + (long long)j * n] = A[p + (long long)j * n];
                A[p + (long long)j * n] = t;
            }
        __syncthreads();
        const double2 pv = A[k + (long long)k * n];
        const double2 pinv = cdiv(make_double2(1.0, 0.0), pv);
        for (int j = tid; j < n; j += kGJThreads) {
            const double2 a = (j == k) ? make_double2(1.0, 0.0) : A[k + (long long)j * n];
            rowk[j] = cmul(a, pinv);
            colk[j] = (j == k) ? make_double2(0.0, 0.0) : A[j + (long long)k * n];
        }
        __syncthreads();
        // rank-1 update; threads run down columns' rows for coalescing
        for (long long e = tid; e < (long long)n * n; e += kGJThreads) {
            const int i = (int)(e % n), j = (int)(e / n);
            double2 a;
            if (i == k) a = rowk[j];
            else {
                a = (j == k) ? make_double2(0.0, 0.0) : A[e];
                const double2 f = colk[i], r = rowk[j];
                a.x -= f.x * r.x - f.y * r.y;
                a.y -= f.x * r.y + f.y * r.x;
            }
            A[e] = a;
        }
        __syncthreads();
    }
    // undo the row interchanges as column swaps, in reverse order
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int i = tid; i < n; i += kGJThreads) {
                const double2 t = A[i + (long long)k * n];
                A[i + (long long)k * n] = A[i + (long long)p * n];
                A[i + (long long)p * n] = t;
            }
        __syncthreads();
    }
}

// y = A x, A column-major n x n: one warp per 32-row strip would be uncoalesced for
// column-major, so each thread owns one row and the block walks the columns.
constexpr int kGemvRows = 32, kGemvParts = 32;
__global__ void __launch_bounds__(kGemvRows * kGemvParts)
dense_gemv_kernel(int n, const double2* __restrict__ A, const double2* __restrict__ x, double2* __restrict__ y) {
    // block = 32 rows x 32 column groups: the column walk of one row is split 32 ways (the matrix is a
    // few hundred columns wide and the solve sits on the critical path of every multigrid cycle), then
    // the groups are summed in fixed order
    extern __shared__ double2 xs[];
    __shared__ double2 part[kGemvParts][kGemvRows + 1];
    const int tid = threadIdx.y * kGemvRows + threadIdx.x;
    for (int j = tid; j < n; j += kGemvRows * kGemvParts) xs[j] = x[j];
    __syncthreads();
    const int i = blockIdx.x * kGemvRows + threadIdx.x;
    double2 a0 = make_double2(0.0, 0.0);
    if (i < n)
        for (int j = threadIdx.y; j < n; j += kGemvParts) cfma(a0, A[i + (long long)j * n], xs[j]);
    part[threadIdx.y][threadIdx.x] = a0;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double2 s = make_double2(0.0, 0.0);
#pragma unroll 8
        for (int g = 0; g < kGemvParts; ++g) s = cadd(s, part[g][threadIdx.x]);
        y[i] = s;
    }
}

static int red_blocks(long long n) {
    long long b = ceil_div<long long>(n, kRedThreads);
    if (b > kRedMaxBlocks) b = kRedMaxBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace hx

using namespace hx;

extern "C" int64_t hx_reduce_scratch_bytes(int k) {
    if (k < 1) k = 1;
    return (int64_t)sizeof(ScratchHeader) + (int64_t)kRedMaxBlocks * k * sizeof(double2);
}

extern "C" int hx_multi_dot(int64_t n, int k, const double* V, int64_t ld, const double* w, int conj, double* out,
                            void* scratch, hx_stream_t stream) {
    if (k <= 0) return HX_OK;
    if (!V || !w || !out || !scratch) return fail(HX_ERR_ARG, "hx_multi_dot: null pointer%s%s");
    ScratchHeader* hdr = (ScratchHeader*)scratch;
    double2* partial = (double2*)(hdr + 1);
    dim3 grid(red_blocks(n), ceil_div(k, kDotTile));
    if (conj)
        multi_dot_kernel<true><<<grid, kRedThreads, 0, (cudaStream_t)stream>>>(n, k, (const double2*)V, ld, (const double2*)w,
                                                                             (double2*)out, hdr, partial);
    else
        multi_dot_kernel<false><<<grid, kRedThreads, 0, (cudaStream_t)stream>>>(n, k, (const double2*)V, ld, (const double2*)w,
                                                                              (double2*)out, hdr, partial);
    return check_launch("multi_dot_kernel");
}

extern "C" int hx_multi_axpy(int64_t n, int k, const double* V, int64_t ld, const double* h, double* w, double* hacc,
                             double* nrm2_out, void* scratch, hx_stream_t stream) {
    if (k < 0) return fail(HX_ERR_ARG, "hx_multi_axpy: k<0%s%s");
    if (nrm2_out && !scratch) return fail(HX_ERR_ARG, "hx_multi_axpy: scratch required for the norm%s%s");
    ScratchHeader* hdr = (ScratchHeader*)scratch;
    double* partial = scratch ? (double*)(hdr + 1) : nullptr;
    multi_axpy_kernel<<<red_blocks(n), kRedThreads, (size_t)(k > 0 ? k : 1) * sizeof(double2), (cudaStream_t)stream>>>(
        n, k, (const double2*)V, ld, (const double2*)h, (double2*)w, (double2*)hacc, nrm2_out, hdr, partial);
    return check_launch("multi_axpy_kernel");
}

extern "C" int hx_scale_copy(int64_t n, const double* w, const double* nrm2_dev, const double* alpha_h, double* out,
                             hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    double2 a = alpha_h ? h2c(alpha_h) : make_double2(1.0, 0.0);
    scale_copy_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(n, (const double2*)w, nrm2_dev, a, (double2*)out);
    return check_launch("scale_copy_kernel");
}

extern "C" int hx_axpby(int64_t n, const double* a_h, const double* x, const double* b_h, double* y, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    const double2 b = b_h ? h2c(b_h) : make_double2(0.0, 0.0);
    const bool use_y = b_h && (b.x != 0.0 || b.y != 0.0);
    axpby_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(n, h2c(a_h), (const double2*)x, b, (double2*)y, use_y);
    return check_launch("axpby_kernel");
}

extern "C" int hx_basis_rotate(int64_t n, int m, int kout, const double* V, int64_t ld, const double* Q, int ldq,
                               double* Vout, int64_t ldout, hx_stream_t stream) {
    if (n <= 0 || kout <= 0) return HX_OK;
    if (m > 512) return fail(HX_ERR_CAPACITY, "hx_basis_rotate: m > 512%s%s");
    dim3 grid(red_blocks(n), ceil_div(kout, kRotTile));
    basis_rotate_kernel<<<grid, kRedThreads, (size_t)m * kRotTile * sizeof(double2), (cudaStream_t)stream>>>(
        n, m, kout, (const double2*)V, ld, (const double2*)Q, ldq, (double2*)Vout, ldout);
    return check_launch("basis_rotate_kernel");
}

extern "C" int hx_basis_rotate_dmma(int64_t n, int m, int kout, const double* V, int64_t ld, const double* Q, int ldq,
                                    double* Vout, int64_t ldout, hx_stream_t stream) {
    if (n <= 0 || kout <= 0) return HX_OK;
    if (m > 512) return fail(HX_ERR_CAPACITY, "hx_basis_rotate_dmma: m > 512%s%s");
    const int m4 = (m + 3) & ~3;
    const size_t smem = (size_t)3 * kMmaTileM * m4 * sizeof(double);
    if (smem > 48 * 1024)
        HX_CUDA(cudaFuncSetAttribute(basis_rotate_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long tiles = (n + 7) / 8;
    long long gx = ceil_div<long long>(tiles, kMmaWarps);
    if (gx > kNumSMs * 8) gx = kNumSMs * 8;
    dim3 grid((unsigned)gx, ceil_div(kout, kMmaTileM));
    basis_rotate_dmma_kernel<<<grid, kMmaWarps * 32, smem, (cudaStream_t)stream>>>(
        n, m, kout, (const double2*)V, ld, (const double2*)Q, ldq, (double2*)Vout, ldout);
    return check_launch("basis_rotate_dmma_kernel");
}

extern "C" int hx_dense_inverse(int n, double* a, int32_t* info_dev, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    if (n > 1400) return fail(HX_ERR_CAPACITY, "hx_dense_inverse: n > 1400%s%s");
    // info_dev: int32[1 + n] (info, then pivot scratch)
    const size_t smem = (size_t)2 * n * sizeof(double2);
    if (smem > 48 * 1024)
        HX_CUDA(cudaFuncSetAttribute(dense_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_inverse_kernel<<<1, kGJThreads, smem, (cudaStream_t)stream>>>(n, (double2*)a, info_dev, info_dev + 1);
    return check_launch("dense_inverse_kernel");
}

extern "C" int hx_dense_gemv(int n, const double* a, const double* x, double* y, hx_stream_t stream) {
    if (n <= 0) return HX_OK;
    dense_gemv_kernel<<<ceil_div(n, kGemvRows), dim3(kGemvRows, kGemvParts), (size_t)n * sizeof(double2), (cudaStream_t)stream>>>(
        n, (const double2*)a, (const double2*)x, (double2*)y);
    return check_launch("dense_gemv_kernel");
}
