// Multi-GPU data path over NVLink peer memory (row e of the scope table): the SpMV halo exchange
// and the small all-reduces of the Krylov solvers as plain kernels that store into the
// neighbours' HBM through NVSwitch and synchronise with sequence flags -- no NCCL call, no host
// round trip, capturable in a CUDA graph with the kernels around them.
//
// What this replaces: PETSc's VecScatter (MatMult on MPIAIJ: helmholtz_x/petsc4py_utils.py:86,96
// under mpirun) and the MPI_Allreduce inside VecDot/VecNorm/BVOrthogonalize
// (helmholtz_x/eigensolvers.py:62,113).
//
// Memory model used throughout: data stores to peer memory, then __threadfence_system(), then a
// st.release.sys of a monotonically increasing sequence number into the peer's flag block; the
// consumer spins with ld.acquire.sys on its OWN flag block (local HBM/L2, cheap to poll) and
// reads the payload with ld.global.cv / in a later kernel.  Every wait has a wall-clock timeout
// that raises the error word instead of hanging the GPU.
#include "hx_common.cuh"

namespace hx {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long kWaitTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

// spin until *flag >= seq; false (and *err = code) on timeout
__device__ __forceinline__ bool wait_ge(const unsigned long long* flag, unsigned long long seq, int* err, int code) {
    if (ld_acquire_sys(flag) >= seq) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(40);
        if (global_ns() - t0 > kWaitTimeoutNs) {
            atomicExch(err, code);
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ bool last_block(unsigned int* counter, unsigned int total) {
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(counter, 1u);
        is_last = (prev == total - 1);
    }
    __syncthreads();
    return is_last;
}

// flag block of one rank: [HX_PEER_FLAG_KINDS][world] u64.  kind 0: "entered exchange s" (all
// halo reads of exchange s-1 are done), kind 1: "halo data of exchange s has landed", kind 2: all-reduce
__device__ __forceinline__ unsigned long long* flag_at(unsigned long long* base, int kind, int world, int who) {
    return base + (long long)kind * world + who;
}

// One neighbour exchange = one launch.  Per channel (this rank, neighbour q) the sequence number is
// kept in device memory (chan_seq[q]) so a CUDA-graph replay keeps counting.
template <typename T>
__global__ void __launch_bounds__(256)
halo_exchange_kernel(hx_peer_halo_desc a, const T* __restrict__ x) {
    __shared__ unsigned long long s_seq[HX_PEER_MAX];
    const int t = threadIdx.x;
    auto my_flags = reinterpret_cast<unsigned long long*>(a.my_flags);
    auto chan_seq = reinterpret_cast<unsigned long long*>(a.chan_seq);
    if (t < a.n_nb) s_seq[t] = chan_seq[a.nb_rank[t]] + 1ull;
    __syncthreads();
    // phase A: every neighbour has entered this exchange => it no longer reads the ghost values I overwrite
    if (blockIdx.x == 0 && t < a.n_nb)
        st_release_sys(flag_at(reinterpret_cast<unsigned long long*>(a.nb_flags[t]), 0, a.world, a.rank), s_seq[t]);
    if (t < a.n_nb) wait_ge(flag_at(my_flags, 0, a.world, a.nb_rank[t]), s_seq[t], a.err, 1);
    __syncthreads();
    // phase B: my interface values straight into the neighbours' ghost tails (NVLink stores)
    const long long total = a.send_ptr[a.n_nb];
    for (long long i = (long long)blockIdx.x * blockDim.x + t; i < total; i += (long long)gridDim.x * blockDim.x) {
        int nb = 0;
        while (i >= a.send_ptr[nb + 1]) ++nb;
        reinterpret_cast<T*>(a.dst[nb])[i - a.send_ptr[nb]] = x[a.send_idx[i]];
    }
    if (last_block(reinterpret_cast<unsigned int*>(a.block_counter), gridDim.x)) {
        if (t < a.n_nb) {
            st_release_sys(flag_at(reinterpret_cast<unsigned long long*>(a.nb_flags[t]), 1, a.world, a.rank), s_seq[t]);
            wait_ge(flag_at(my_flags, 1, a.world, a.nb_rank[t]), s_seq[t], a.err, 2);
            chan_seq[a.nb_rank[t]] = s_seq[t];
        }
        if (t == 0) *reinterpret_cast<unsigned int*>(a.block_counter) = 0u;
    }
}

// One-shot all-reduce (sum) of `count` scalars: every rank stores its contribution into slot
// [parity][rank] of every rank's slot area, then sums the slots in rank order -- the same order on
// every rank, so the result is bitwise identical everywhere (the replicated host logic relies on it).
// Slots are double-buffered by the parity of the sequence number: a rank can only start all-reduce
// s+2 after every rank has finished summing s.  gridDim.x <= 32 (all CTAs must be co-resident).
template <typename T>
__global__ void __launch_bounds__(256)
allreduce_kernel(hx_peer_allreduce_desc a, const T* in, T* out, long long count) {
    const int t = threadIdx.x;
    auto my_flags = reinterpret_cast<unsigned long long*>(a.my_flags);
    auto seq_p = reinterpret_cast<unsigned long long*>(a.seq);
    const unsigned long long s = *seq_p + 1ull;
    const long long par = (long long)(s & 1ull);
    const long long per = (count + gridDim.x - 1) / gridDim.x;
    const long long c0 = per * blockIdx.x, c1 = min(count, c0 + per);
    for (long long i = c0 + t; i < c1; i += blockDim.x) {
        const T v = in[i];
        for (int q = 0; q < a.world; ++q)
            reinterpret_cast<T*>(reinterpret_cast<char*>(a.slots[q]) + (par * a.world + a.rank) * a.slot_bytes)[i] = v;
    }
    auto counters = reinterpret_cast<unsigned int*>(a.block_counter);
    if (last_block(counters, gridDim.x)) {
        if (t < a.world) st_release_sys(flag_at(reinterpret_cast<unsigned long long*>(a.flags[t]), 2, a.world, a.rank), s);
        if (t == 0) counters[0] = 0u;
    }
    if (t < a.world) wait_ge(flag_at(my_flags, 2, a.world, t), s, a.err, 3);
    __syncthreads();
    const char* mine = reinterpret_cast<const char*>(a.slots[a.rank]);
    for (long long i = c0 + t; i < c1; i += blockDim.x) {
        T acc = __ldcv(reinterpret_cast<const T*>(mine + (par * a.world) * a.slot_bytes) + i);
        for (int r = 1; r < a.world; ++r)
            acc += __ldcv(reinterpret_cast<const T*>(mine + (par * a.world + r) * a.slot_bytes) + i);
        out[i] = acc;
    }
    if (last_block(counters + 1, gridDim.x)) {
        if (t == 0) {
            *seq_p = s;
            counters[1] = 0u;
        }
    }
}

}  // namespace hx

using namespace hx;

extern "C" {

int hx_peer_alloc(int64_t bytes, void** ptr_out_h, unsigned char* handle64_h) {
    if (bytes <= 0 || !ptr_out_h || !handle64_h) return fail(HX_ERR_ARG, "hx_peer_alloc: bad arguments");
    void* p = nullptr;
    HX_CUDA(cudaMalloc(&p, (size_t)bytes));
    HX_CUDA(cudaMemset(p, 0, (size_t)bytes));
    HX_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(HX_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64_h, &h, 64);
    *ptr_out_h = p;
    return HX_OK;
}

int hx_peer_open(const unsigned char* handle64_h, void** ptr_out_h) {
    if (!handle64_h || !ptr_out_h) return fail(HX_ERR_ARG, "hx_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_h, 64);
    void* p = nullptr;
    HX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out_h = p;
    return HX_OK;
}

int hx_peer_close(void* ptr) {
    HX_CUDA(cudaIpcCloseMemHandle(ptr));
    return HX_OK;
}

int hx_peer_free(void* ptr) {
    HX_CUDA(cudaFree(ptr));
    return HX_OK;
}

int hx_peer_halo_exchange(const hx_peer_halo_desc* plan_h, const void* x_local, int elem_bytes, hx_stream_t stream) {
    if (!plan_h || plan_h->n_nb < 0 || plan_h->n_nb > HX_PEER_MAX) return fail(HX_ERR_ARG, "hx_peer_halo_exchange: bad plan");
    if (plan_h->n_nb == 0) return HX_OK;
    const long long total = plan_h->send_ptr[plan_h->n_nb];
    long long g_ = ceil_div<long long>(total, 2048);
    int grid = (int)(g_ < 1 ? 1 : (g_ > 32 ? 32 : g_));
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_bytes == 16)
        halo_exchange_kernel<double2><<<grid, 256, 0, st>>>(*plan_h, (const double2*)x_local);
    else if (elem_bytes == 8)
        halo_exchange_kernel<float2><<<grid, 256, 0, st>>>(*plan_h, (const float2*)x_local);
    else
        return fail(HX_ERR_ARG, "hx_peer_halo_exchange: elem_bytes must be 8 or 16");
    return check_launch("halo_exchange_kernel");
}

int hx_peer_allreduce(const hx_peer_allreduce_desc* ar_h, const void* in, void* out, int64_t count, int is_f32,
                      hx_stream_t stream) {
    if (!ar_h || ar_h->world < 1 || ar_h->world > HX_PEER_MAX + 1) return fail(HX_ERR_ARG, "hx_peer_allreduce: bad descriptor");
    if (count <= 0) return HX_OK;
    if (count * (is_f32 ? 4 : 8) > ar_h->slot_bytes) return fail(HX_ERR_CAPACITY, "hx_peer_allreduce: count exceeds the slot size");
    long long g_ = ceil_div<long long>(count, 4096);
    int grid = (int)(g_ < 1 ? 1 : (g_ > 32 ? 32 : g_));
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f32)
        allreduce_kernel<float><<<grid, 256, 0, st>>>(*ar_h, (const float*)in, (float*)out, count);
    else
        allreduce_kernel<double><<<grid, 256, 0, st>>>(*ar_h, (const double*)in, (double*)out, count);
    return check_launch("allreduce_kernel");
}

}  // extern "C"
