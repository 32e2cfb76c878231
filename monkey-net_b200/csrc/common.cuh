// Shared helpers for libmonkey_b200 (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MK_EXPORT extern "C" __attribute__((visibility("default")))
// the public header declares the same symbols without the visibility attribute; re-declaring with it first keeps them exported

void mk_set_error(const char* fmt, ...);
int mk_check_launch(const char* what);

#define MK_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            mk_set_error(__VA_ARGS__);   \
            return -1;                   \
        }                                \
    } while (0)

static inline int mk_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

static inline long long mk_cdiv(long long a, long long b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per function AND per device: remember, per kernel, on which devices it has been applied
// (`done` is that kernel's static mask).  One process drives one GPU in this design, but a process that touches a
// second device must not launch with the first device's attribute state.
static inline unsigned long long mk_attr_needed(unsigned long long done) {  // 0 = already applied on this device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    return (done & bit) ? 0ull : bit;
}

// ---------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ void fma4(float4& acc, float4 a, float s) {
    acc.x = fmaf(a.x, s, acc.x); acc.y = fmaf(a.y, s, acc.y); acc.z = fmaf(a.z, s, acc.z); acc.w = fmaf(a.w, s, acc.w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NV values per thread; result valid in every thread.  `red` needs NV*32 floats of smem.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float t = (lane < nw) ? red[i * 32 + lane] : 0.f;
        v[i] = warp_sum(t);
    }
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : -INFINITY;
    return warp_max(t);
}

// reference coordinate grid (modules/util.py:26-42): x_j = 2*(j/(w-1)) - 1
__device__ __forceinline__ float grid_coord(int j, int n) { return 2.f * ((float)j / (float)(n - 1)) - 1.f; }

// nearest source index of F.interpolate(size=...) (ATen nearest_neighbor_compute_source_index)
__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
    float scale = (float)in / (float)out;
    int s = (int)floorf((float)dst * scale);
    return s < in - 1 ? s : in - 1;
}
// linear source index, align_corners=False (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void linear_src(int dst, int in, int out, int& i0, int& i1, float& l1) {
    float scale = (float)in / (float)out;
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

// ---------------------------------------------------------------------------------------------- index arithmetic
// The elementwise / gather kernels decode (n, h, w, channel-vector) from a flat thread index.  64-bit div/mod costs
// ~100 SASS instructions each on sm_100 and made those "HBM-bound" kernels issue-bound (profiles/r1: 45 % SM busy at
// 14 % DRAM).  All extents here are < 2^31, so the index math is 32-bit, and divisions by power-of-two extents
// (every channel count and resolution the configs use) are shifts.
struct FastDiv {
    unsigned d;
    int sh;  // >= 0: d == 1 << sh
};
static inline FastDiv make_fastdiv(long long d) {
    FastDiv f;
    f.d = (unsigned)d;
    f.sh = -1;
    if (d > 0 && (d & (d - 1)) == 0) {
        f.sh = 0;
        while ((1LL << f.sh) < d) ++f.sh;
    }
    return f;
}
// q = n / f.d, r = n % f.d
__device__ __forceinline__ unsigned fd_divmod(unsigned n, const FastDiv f, unsigned& r) {
    unsigned q = f.sh >= 0 ? n >> f.sh : n / f.d;
    r = n - q * f.d;
    return q;
}
