// Shared device/host helpers for libhx_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hx_b200.h"

namespace hx {

extern thread_local char g_err[512];
extern int64_t g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

inline int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return HX_OK;
}

#define HX_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return hx::fail(HX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- complex arithmetic on double2 (x=re, y=im) ------------------------------
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // conj(a)*b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(double2& acc, double2 a, double2 b) {  // acc += conj(a)*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ void rfma(double2& acc, double a, double2 b) {  // acc += a*b, a real
    acc.x = fma(a, b.x, acc.x);
    acc.y = fma(a, b.y, acc.y);
}
__host__ __device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
inline double2 h2c(const double* p) { return make_double2(p[0], p[1]); }

// ---- complex64 (float2) overloads: the mixed-precision multigrid cycle ---------------
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void rfma(float2& acc, float a, float2 b) {
    acc.x = fmaf(a, b.x, acc.x);
    acc.y = fmaf(a, b.y, acc.y);
}
// generic multiply-accumulate picked by the matrix value type
__device__ __forceinline__ void mac(double2& acc, double2 a, double2 x) { cfma(acc, a, x); }
__device__ __forceinline__ void mac(double2& acc, double a, double2 x) { rfma(acc, a, x); }
__device__ __forceinline__ void mac(float2& acc, float2 a, float2 x) { cfma(acc, a, x); }
__device__ __forceinline__ void mac(float2& acc, float a, float2 x) { rfma(acc, a, x); }
template <typename V> __device__ __forceinline__ V vzero();
template <> __device__ __forceinline__ double2 vzero<double2>() { return make_double2(0.0, 0.0); }
template <> __device__ __forceinline__ float2 vzero<float2>() { return make_float2(0.f, 0.f); }
template <typename V> struct scalar_of;
template <> struct scalar_of<double2> { typedef double type; };
template <> struct scalar_of<float2> { typedef float type; };
template <typename V> __host__ __device__ __forceinline__ V from_c128(double2 z);
template <> __host__ __device__ __forceinline__ double2 from_c128<double2>(double2 z) { return z; }
template <> __host__ __device__ __forceinline__ float2 from_c128<float2>(double2 z) { return make_float2((float)z.x, (float)z.y); }

// ---- streaming (read-once) loads: keep matrix streams out of L1 so the gathered
// x vector stays cached (guide: ld.global.nc.L1::no_allocate) -------------------
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// ---- warp / block reductions (fixed order => deterministic) ---------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double2 warp_sum(double2 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace hx
