// The tail of the multigrid cycle -- the last smoothed level and the dense coarsest solve below it -- as ONE
// persistent kernel: pre-smoothing, residual, restriction, dense coarse solve, prolongation, post-smoothing,
// separated by a grid-wide barrier instead of kernel boundaries.  On the levels this serves (a few thousand
// rows) every kernel of the unfused sequence is launch / dependency-latency bound (~8 us each, ~80 us per
// visit, and a W-cycle visits the tail four times per GMRES iteration); fused, a visit costs the barriers.
// What it replaces: the coarse part of PETSc's PC behind ST sinvert (helmholtz_x/eigensolvers.py:49-50).
#include "hx_common.cuh"

namespace hx {

constexpr int kTailThreads = 256;
constexpr int kTailCtasPerSm = 4;      // co-resident by construction: 4 x 256 threads, <= 64 registers, no shared arrays

// sense-reversing grid barrier; the grid is at most kTailCtasPerSm CTAs per SM (all co-resident)
__device__ __forceinline__ void tail_barrier(unsigned int* bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* gen = bar + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) {}
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float2 warp_sum(float2 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// one warp per row: sum_k val[k] * x[col[k]]; loads bypass L1 where other CTAs wrote the vector before a barrier
template <typename MV>
__device__ __forceinline__ float2 tail_row(const int* __restrict__ ptr, const int* __restrict__ idx,
                                           const MV* __restrict__ val, const float2* x, int row, int lane) {
    float2 acc = make_float2(0.f, 0.f);
    const int s = ptr[row], e = ptr[row + 1];
    for (int k = s + lane; k < e; k += 32) mac(acc, val[k], __ldcg(x + idx[k]));
    return warp_sum(acc);
}

__global__ void __launch_bounds__(kTailThreads, kTailCtasPerSm)
tail_kernel(hx_tail_desc a, const float2* __restrict__ b) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kTailThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kTailThreads) >> 5;
    const int tid = blockIdx.x * kTailThreads + threadIdx.x, nthreads = gridDim.x * kTailThreads;
    auto dinv = reinterpret_cast<const float2*>(a.dinv);
    auto av = reinterpret_cast<const float2*>(a.a_val);
    float2* cur = reinterpret_cast<float2*>(a.buf0);
    float2* oth = reinterpret_cast<float2*>(a.buf1);
    float2* r = reinterpret_cast<float2*>(a.r);
    unsigned int* bar = reinterpret_cast<unsigned int*>(a.barrier);
    // pre-smoothing: first sweep from zero, then nu-1 Jacobi sweeps
    for (int i = tid; i < a.n; i += nthreads) cur[i] = cscale(a.omega[0], cmul(dinv[i], b[i]));
    tail_barrier(bar, gridDim.x);
    for (int s = 1; s < a.nu; ++s) {
        for (int i = warp; i < a.n; i += nwarps) {
            const float2 ax = tail_row(a.a_ptr, a.a_idx, av, cur, i, lane);
            if (lane == 0) oth[i] = cadd(__ldcg(cur + i), cscale(a.omega[s], cmul(dinv[i], csub(b[i], ax))));
        }
        tail_barrier(bar, gridDim.x);
        float2* t = cur; cur = oth; oth = t;
    }
    // residual
    for (int i = warp; i < a.n; i += nwarps) {
        const float2 ax = tail_row(a.a_ptr, a.a_idx, av, cur, i, lane);
        if (lane == 0) r[i] = csub(b[i], ax);
    }
    tail_barrier(bar, gridDim.x);
    // restriction (float32 R, nc rows), result in complex128 for the dense solve
    double2* bc = reinterpret_cast<double2*>(a.bc);
    double2* xc = reinterpret_cast<double2*>(a.xc);
    for (int k = warp; k < a.nc; k += nwarps) {
        const float2 v = tail_row(a.r_ptr, a.r_idx, a.r_val, r, k, lane);
        if (lane == 0) bc[k] = make_double2((double)v.x, (double)v.y);
    }
    tail_barrier(bar, gridDim.x);
    // coarsest level: xc = Ainv bc (column-major complex128 inverse)
    auto cinv = reinterpret_cast<const double2*>(a.coarse_inv);
    for (int k = warp; k < a.nc; k += nwarps) {
        double2 acc = make_double2(0.0, 0.0);
        for (int j = lane; j < a.nc; j += 32) cfma(acc, cinv[k + (long long)j * a.nc], __ldcg(bc + j));
        acc = warp_sum(acc);
        if (lane == 0) xc[k] = acc;
    }
    tail_barrier(bar, gridDim.x);
    // prolongation: cur += P xc (float32 P, n rows)
    for (int i = warp; i < a.n; i += nwarps) {
        double2 acc = make_double2(0.0, 0.0);
        const int s = a.p_ptr[i], e = a.p_ptr[i + 1];
        for (int k = s + lane; k < e; k += 32) rfma(acc, (double)a.p_val[k], __ldcg(xc + a.p_idx[k]));
        acc = warp_sum(acc);
        if (lane == 0) {
            const float2 c = __ldcg(cur + i);
            cur[i] = make_float2(c.x + (float)acc.x, c.y + (float)acc.y);
        }
    }
    tail_barrier(bar, gridDim.x);
    // post-smoothing: nu sweeps
    for (int s = 0; s < a.nu; ++s) {
        for (int i = warp; i < a.n; i += nwarps) {
            const float2 ax = tail_row(a.a_ptr, a.a_idx, av, cur, i, lane);
            if (lane == 0) oth[i] = cadd(__ldcg(cur + i), cscale(a.omega[s], cmul(dinv[i], csub(b[i], ax))));
        }
        if (s + 1 < a.nu) tail_barrier(bar, gridDim.x);
        float2* t = cur; cur = oth; oth = t;
    }
}

}  // namespace hx

using namespace hx;

extern "C" int hx_amg_tail(const hx_tail_desc* d, const float* b_c64, hx_stream_t stream) {
    if (!d || d->n <= 0 || d->nc <= 0 || d->nu < 1 || d->nu > 4) return fail(HX_ERR_ARG, "hx_amg_tail: bad descriptor%s%s");
    int grid = ceil_div(d->n * 32, kTailThreads);          // one warp per row is all the parallelism there is
    if (grid > kNumSMs * kTailCtasPerSm) grid = kNumSMs * kTailCtasPerSm;   // the barrier needs co-residency
    if (grid < 1) grid = 1;
    tail_kernel<<<grid, kTailThreads, 0, (cudaStream_t)stream>>>(*d, (const float2*)b_c64);
    return check_launch("tail_kernel");
}
