// Keypoint-side kernels: spatial softmax + soft-argmax moments (keypoint_detector.py:43-78,101-107), gaussian
// heatmap rendering + movement embedding (keypoint_detector.py:7-40, movement_embedding.py:42-92) and the
// dense-motion head (dense_motion_module.py:52-76), each with a hand-written backward.  All HBM/latency bound
// warp-/block-reduction kernels; the per-keypoint 2x2 algebra (inverse, smallest singular value and their
// derivatives) is closed form in registers (replaces torch.gesv, modules/util.py:220-224).
#include "common.cuh"
#include "../../include/monkey_b200.h"

// ================================================================================================ keypoint head
// One block per (frame n, keypoint k).  Three passes over the H*W logits of that channel (L2 resident).
__global__ void __launch_bounds__(256) k_kp_head_fwd(const float* __restrict__ logits, int H, int W, int K, int ld,
                                                     float invT, int var_mode, float clip, float* __restrict__ mean,
                                                     float* __restrict__ var, float* __restrict__ aux) {
    __shared__ float red[8 * 32];
    const int n = blockIdx.x / K, k = blockIdx.x % K;
    const int hw = H * W;
    const float* base = logits + (long long)n * hw * ld + k;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) m = fmaxf(m, base[(long long)i * ld] * invT);
    m = block_max(m, red);
    float v5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // sum e, sum e*gx, sum e*gy, sum gx, sum gy
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        float e = expf(base[(long long)i * ld] * invT - m);
        float gx = grid_coord(i % W, W), gy = grid_coord(i / W, H);
        v5[0] += e; v5[1] += e * gx; v5[2] += e * gy; v5[3] += gx; v5[4] += gy;
    }
    block_sum<5>(v5, red);
    const float se = v5[0];
    const float mx = v5[1] / se + 1e-7f * v5[3];
    const float my = v5[2] / se + 1e-7f * v5[4];
    float c3[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        float p = expf(base[(long long)i * ld] * invT - m) / se + 1e-7f;
        float dx = grid_coord(i % W, W) - mx, dy = grid_coord(i / W, H) - my;
        c3[0] += p * dx * dx; c3[1] += p * dx * dy; c3[2] += p * dy * dy;
    }
    block_sum<3>(c3, red);
    if (threadIdx.x == 0) {
        const long long o = (long long)n * K + k;
        mean[o * 2 + 0] = mx;
        mean[o * 2 + 1] = my;
        float a = c3[0], b = c3[1], c = c3[1], d = c3[2];
        float* ax = aux + o * 8;
        ax[0] = m; ax[1] = se; ax[2] = a; ax[3] = b; ax[4] = c; ax[5] = d; ax[6] = 0.f; ax[7] = 0.f;
        if (var_mode == 0) {
            if (clip > 0.f) {
                float s1 = a * a + b * b + c * c + d * d;
                float u = a * a + b * b - c * c - d * d, vv = a * c + b * d;
                float s2 = sqrtf(u * u + 4.f * vv * vv);
                float sg = sqrtf((s1 - s2) / 2.f);
                float f = fmaxf(clip, sg);
                a = f * a / sg; b = f * b / sg; c = f * c / sg; d = f * d / sg;
                ax[6] = sg;
            }
            var[o * 4 + 0] = a; var[o * 4 + 1] = b; var[o * 4 + 2] = c; var[o * 4 + 3] = d;
        } else {
            var[o] = (a + d) * 0.5f;
        }
    }
}

__global__ void __launch_bounds__(256) k_kp_head_bwd(const float* __restrict__ logits, int H, int W, int K, int ld,
                                                     float invT, int var_mode, float clip,
                                                     const float* __restrict__ mean, const float* __restrict__ aux,
                                                     const float* __restrict__ dmean, const float* __restrict__ dvar,
                                                     float* __restrict__ dlogits) {
    __shared__ float red[32];
    const int n = blockIdx.x / K, k = blockIdx.x % K;
    const int hw = H * W;
    const long long o = (long long)n * K + k;
    const float* ax = aux + o * 8;
    const float m = ax[0], se = ax[1];
    const float a = ax[2], b = ax[3], c = ax[4], d = ax[5];
    const float mx = mean[o * 2], my = mean[o * 2 + 1];
    // ---- gradient w.r.t. the raw covariance
    float g00, g01, g10, g11;
    if (var_mode == 0) {
        g00 = dvar[o * 4]; g01 = dvar[o * 4 + 1]; g10 = dvar[o * 4 + 2]; g11 = dvar[o * 4 + 3];
        if (clip > 0.f) {
            const float sg = ax[6];
            if (sg < clip) {
                // out = clip * V / sg ;  d/dV = clip/sg  -  clip/sg^2 * <dvar, V> * dsg/dV
                const float f = clip / sg;
                const float inner = g00 * a + g01 * b + g10 * c + g11 * d;
                const float dsg = -clip / (sg * sg) * inner;
                const float u = a * a + b * b - c * c - d * d, vv = a * c + b * d;
                const float s2 = sqrtf(u * u + 4.f * vv * vv);
                const float ds1 = dsg / (4.f * sg), ds2 = -dsg / (4.f * sg);
                const float du = s2 > 0.f ? ds2 * u / s2 : 0.f, dv = s2 > 0.f ? ds2 * 4.f * vv / s2 : 0.f;
                g00 = f * g00 + ds1 * 2.f * a + du * 2.f * a + dv * c;
                g01 = f * g01 + ds1 * 2.f * b + du * 2.f * b + dv * d;
                g10 = f * g10 + ds1 * 2.f * c - du * 2.f * c + dv * a;
                g11 = f * g11 + ds1 * 2.f * d - du * 2.f * d + dv * b;
            }
        }
    } else {
        g00 = g11 = dvar[o] * 0.5f;
        g01 = g10 = 0.f;
    }
    // the reference builds var with b == c from one accumulation each: raw[0,1] and raw[1,0] both receive gradient
    const float gxy = g01 + g10;
    // mean gradient, including the (tiny) path through the centring of the covariance: sum p' = 1 + HW*1e-7
    const float leak = (float)hw * 1e-7f;
    const float amx = dmean[o * 2] + (2.f * g00 * mx + gxy * my) * leak;
    const float amy = dmean[o * 2 + 1] + (gxy * mx + 2.f * g11 * my) * leak;
    const float* base = logits + (long long)n * hw * ld + k;
    float s[1] = {0.f};
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        float p = expf(base[(long long)i * ld] * invT - m) / se;
        float gx = grid_coord(i % W, W), gy = grid_coord(i / W, H);
        float dx = gx - mx, dy = gy - my;
        float dp = gx * amx + gy * amy + g00 * dx * dx + gxy * dx * dy + g11 * dy * dy;
        s[0] += p * dp;
    }
    block_sum<1>(s, red);
    float* ob = dlogits + (long long)n * hw * ld + k;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        float p = expf(base[(long long)i * ld] * invT - m) / se;
        float gx = grid_coord(i % W, W), gy = grid_coord(i / W, H);
        float dx = gx - mx, dy = gy - my;
        float dp = gx * amx + gy * amy + g00 * dx * dx + gxy * dx * dy + g11 * dy * dy;
        ob[(long long)i * ld] = invT * p * (dp - s[0]);
    }
}

// ------------------------------------------------------------------------------------------------ chunked keypoint head
// The one-block-per-(frame, keypoint) kernels above read their channel with a 4-byte load every `ld` floats and run
// N*K blocks: at 256x256 x 16 frames x 10 keypoints that is 160 blocks x 3 dependent passes = 410 us (fwd) + 390 us
// (bwd) for 50 MB of logits.  The chunked kernels give one block a CHUNK of pixels and ALL channels of them (float4
// loads of whole pixels: coalesced, every byte used), N x C blocks; the softmax statistics are combined across chunks
// from per-chunk partials in a fixed order (deterministic, no float atomics).  `scratch` holds the partials.
//   fwd 1: per chunk  local max m_c, se_c = sum e^{v - m_c}, sx_c, sy_c            -> part1[n][c][k][4], gsum[n][c][2]
//   fwd 2: combine -> m, se, mean; per chunk covariance partials                      -> part2[n][c][k][3];
//          the LAST block of a frame (atomic counter) adds them up and writes mean / var / aux
//   bwd 1: per chunk  s_c = sum p * dp                                                 -> part1[n][c][k][0]
//   bwd 2: combine s; dlogits = invT * p * (dp - s) for the chunk, pad channels zeroed (no memset)
constexpr int KPH_THREADS = 256;
constexpr int KPH_MAXC = 64;

static inline int kph_chunks(int hw) {
    int c = hw / 1024;
    if (c < 1) c = 1;
    if (c > KPH_MAXC) c = KPH_MAXC;
    return c;
}
// floats of `scratch` for mk_kp_head_fwd / mk_kp_head_bwd
MK_EXPORT int mk_kp_head_scratch_floats(int N, int H, int W, int K, long long* out) {
    MK_REQUIRE(out != nullptr, "mk_kp_head_scratch_floats: out is NULL");
    const int C = kph_chunks(H * W);
    const int K4 = (K + 3) & ~3;
    *out = (long long)N * ((long long)C * (7 * K4 + 4) + 4);
    return 0;
}

struct KphP {
    const float* logits; int H, W, K, ld, C, chunk; float invT; int var_mode; float clip;
    float* part1; float* gsum; float* part2; unsigned* counter;   // views into scratch
};

__device__ __forceinline__ float f4get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// block-wide reduction of NV per-thread values into red_out[NV] (shared), sum or max
template <int NV, bool MAX>
__device__ __forceinline__ void kph_reduce(float (&v)[NV], float* red /* NV * 8 */, float* out /* NV */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = MAX ? warp_max(v[i]) : warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 8 + wid] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float t = red[threadIdx.x * 8];
#pragma unroll
        for (int w = 1; w < KPH_THREADS / 32; ++w) t = MAX ? fmaxf(t, red[threadIdx.x * 8 + w]) : t + red[threadIdx.x * 8 + w];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

template <int KV>
__global__ void __launch_bounds__(KPH_THREADS) k_kph_fwd1(const KphP p) {
    constexpr int K4 = 4 * KV;
    __shared__ float red[(3 * K4 + 2) * 8];
    __shared__ float bm[K4];
    __shared__ float outv[3 * K4 + 2];
    const int n = blockIdx.y, c = blockIdx.x;
    const int hw = p.H * p.W;
    const int i0 = c * p.chunk, i1 = min(hw, i0 + p.chunk);
    const float* base = p.logits + (long long)n * hw * p.ld;
    if (c == 0 && threadIdx.x == 0) p.counter[n] = 0u;
    float m[K4];
#pragma unroll
    for (int j = 0; j < K4; ++j) m[j] = -INFINITY;
    for (int i = i0 + threadIdx.x; i < i1; i += KPH_THREADS) {
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const float4 v = ldg4(base + (long long)i * p.ld + 4 * q);
            m[4 * q] = fmaxf(m[4 * q], v.x * p.invT); m[4 * q + 1] = fmaxf(m[4 * q + 1], v.y * p.invT);
            m[4 * q + 2] = fmaxf(m[4 * q + 2], v.z * p.invT); m[4 * q + 3] = fmaxf(m[4 * q + 3], v.w * p.invT);
        }
    }
    kph_reduce<K4, true>(m, red, bm);
    float acc[3 * K4 + 2];
#pragma unroll
    for (int j = 0; j < 3 * K4 + 2; ++j) acc[j] = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += KPH_THREADS) {
        const float gx = grid_coord(i % p.W, p.W), gy = grid_coord(i / p.W, p.H);
        acc[3 * K4] += gx; acc[3 * K4 + 1] += gy;
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const float4 v = ldg4(base + (long long)i * p.ld + 4 * q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = expf(f4get(v, j) * p.invT - bm[4 * q + j]);
                acc[4 * q + j] += e; acc[K4 + 4 * q + j] += e * gx; acc[2 * K4 + 4 * q + j] += e * gy;
            }
        }
    }
    kph_reduce<3 * K4 + 2, false>(acc, red, outv);
    if (threadIdx.x < K4) {
        float* o = p.part1 + (((long long)n * p.C + c) * K4 + threadIdx.x) * 4;
        o[0] = bm[threadIdx.x]; o[1] = outv[threadIdx.x]; o[2] = outv[K4 + threadIdx.x]; o[3] = outv[2 * K4 + threadIdx.x];
    }
    if (threadIdx.x == 0) {
        p.gsum[((long long)n * p.C + c) * 2] = outv[3 * K4];
        p.gsum[((long long)n * p.C + c) * 2 + 1] = outv[3 * K4 + 1];
    }
}

template <int KV>
__global__ void __launch_bounds__(KPH_THREADS) k_kph_fwd2(const KphP p, float* __restrict__ mean, float* __restrict__ var,
                                                          float* __restrict__ aux) {
    constexpr int K4 = 4 * KV;
    __shared__ float red[3 * K4 * 8];
    __shared__ float sm[K4], sse[K4], smx[K4], smy[K4];
    __shared__ float outv[3 * K4];
    __shared__ int last;
    const int n = blockIdx.y, c = blockIdx.x;
    const int hw = p.H * p.W;
    if (threadIdx.x < K4) {   // combine the chunk partials of this frame (fixed order)
        const int k = threadIdx.x;
        const float* pp = p.part1 + ((long long)n * p.C * K4 + k) * 4;
        float M = -INFINITY;
        for (int cc = 0; cc < p.C; ++cc) M = fmaxf(M, pp[(long long)cc * K4 * 4]);
        float se = 0.f, sx = 0.f, sy = 0.f, gx = 0.f, gy = 0.f;
        for (int cc = 0; cc < p.C; ++cc) {
            const float* q = pp + (long long)cc * K4 * 4;
            const float f = expf(q[0] - M);
            se += q[1] * f; sx += q[2] * f; sy += q[3] * f;
            gx += p.gsum[((long long)n * p.C + cc) * 2]; gy += p.gsum[((long long)n * p.C + cc) * 2 + 1];
        }
        sm[k] = M; sse[k] = se;
        smx[k] = sx / se + 1e-7f * gx;
        smy[k] = sy / se + 1e-7f * gy;
    }
    __syncthreads();
    const int i0 = c * p.chunk, i1 = min(hw, i0 + p.chunk);
    const float* base = p.logits + (long long)n * hw * p.ld;
    float acc[3 * K4];
#pragma unroll
    for (int j = 0; j < 3 * K4; ++j) acc[j] = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += KPH_THREADS) {
        const float gx = grid_coord(i % p.W, p.W), gy = grid_coord(i / p.W, p.H);
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const float4 v = ldg4(base + (long long)i * p.ld + 4 * q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * q + j;
                const float pr = expf(f4get(v, j) * p.invT - sm[k]) / sse[k] + 1e-7f;
                const float dx = gx - smx[k], dy = gy - smy[k];
                acc[k] += pr * dx * dx; acc[K4 + k] += pr * dx * dy; acc[2 * K4 + k] += pr * dy * dy;
            }
        }
    }
    kph_reduce<3 * K4, false>(acc, red, outv);
    if (threadIdx.x < K4) {
        float* o = p.part2 + (((long long)n * p.C + c) * K4 + threadIdx.x) * 3;
        o[0] = outv[threadIdx.x]; o[1] = outv[K4 + threadIdx.x]; o[2] = outv[2 * K4 + threadIdx.x];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(p.counter + n, 1u) == (unsigned)(p.C - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < p.K) {   // the last block of the frame: covariance totals, clip, outputs
        const int k = threadIdx.x;
        float a = 0.f, b = 0.f, d = 0.f;
        for (int cc = 0; cc < p.C; ++cc) {
            const volatile float* q = p.part2 + (((long long)n * p.C + cc) * K4 + k) * 3;
            a += q[0]; b += q[1]; d += q[2];
        }
        float cq = b;
        const long long o = (long long)n * p.K + k;
        mean[o * 2] = smx[k];
        mean[o * 2 + 1] = smy[k];
        float* ax = aux + o * 8;
        ax[0] = sm[k]; ax[1] = sse[k]; ax[2] = a; ax[3] = b; ax[4] = cq; ax[5] = d; ax[6] = 0.f; ax[7] = 0.f;
        if (p.var_mode == 0) {
            if (p.clip > 0.f) {
                const float s1 = a * a + b * b + cq * cq + d * d;
                const float u = a * a + b * b - cq * cq - d * d, vv = a * cq + b * d;
                const float s2 = sqrtf(u * u + 4.f * vv * vv);
                const float sg = sqrtf((s1 - s2) / 2.f);
                const float f = fmaxf(p.clip, sg);
                a = f * a / sg; b = f * b / sg; cq = f * cq / sg; d = f * d / sg;
                ax[6] = sg;
            }
            var[o * 4] = a; var[o * 4 + 1] = b; var[o * 4 + 2] = cq; var[o * 4 + 3] = d;
        } else {
            var[o] = (a + d) * 0.5f;
        }
    }
}

// per-channel constants of the backward pass (same algebra as k_kp_head_bwd), threads k < K
struct KphB { float m, se, mx, my, amx, amy, g00, gxy, g11; };
__device__ __forceinline__ KphB kph_bwd_consts(int hw, int K, int var_mode, float clip, long long o, const float* mean,
                                               const float* aux, const float* dmean, const float* dvar) {
    KphB r;
    const float* ax = aux + o * 8;
    r.m = ax[0]; r.se = ax[1];
    const float a = ax[2], b = ax[3], c = ax[4], d = ax[5];
    r.mx = mean[o * 2]; r.my = mean[o * 2 + 1];
    float g00, g01, g10, g11;
    if (var_mode == 0) {
        g00 = dvar[o * 4]; g01 = dvar[o * 4 + 1]; g10 = dvar[o * 4 + 2]; g11 = dvar[o * 4 + 3];
        if (clip > 0.f) {
            const float sg = ax[6];
            if (sg < clip) {
                const float f = clip / sg;
                const float inner = g00 * a + g01 * b + g10 * c + g11 * d;
                const float dsg = -clip / (sg * sg) * inner;
                const float u = a * a + b * b - c * c - d * d, vv = a * c + b * d;
                const float s2 = sqrtf(u * u + 4.f * vv * vv);
                const float ds1 = dsg / (4.f * sg), ds2 = -dsg / (4.f * sg);
                const float du = s2 > 0.f ? ds2 * u / s2 : 0.f, dv = s2 > 0.f ? ds2 * 4.f * vv / s2 : 0.f;
                g00 = f * g00 + ds1 * 2.f * a + du * 2.f * a + dv * c;
                g01 = f * g01 + ds1 * 2.f * b + du * 2.f * b + dv * d;
                g10 = f * g10 + ds1 * 2.f * c - du * 2.f * c + dv * a;
                g11 = f * g11 + ds1 * 2.f * d - du * 2.f * d + dv * b;
            }
        }
    } else {
        g00 = g11 = dvar[o] * 0.5f;
        g01 = g10 = 0.f;
    }
    r.g00 = g00; r.g11 = g11; r.gxy = g01 + g10;
    const float leak = (float)hw * 1e-7f;
    r.amx = dmean[o * 2] + (2.f * g00 * r.mx + r.gxy * r.my) * leak;
    r.amy = dmean[o * 2 + 1] + (r.gxy * r.mx + 2.f * g11 * r.my) * leak;
    return r;
}

template <int KV, int PHASE>
__global__ void __launch_bounds__(KPH_THREADS) k_kph_bwd(const KphP p, const float* __restrict__ mean,
                                                         const float* __restrict__ aux, const float* __restrict__ dmean,
                                                         const float* __restrict__ dvar, float* __restrict__ dlogits) {
    constexpr int K4 = 4 * KV;
    __shared__ float red[K4 * 8];
    __shared__ KphB cb[K4];
    __shared__ float stot[K4];
    const int n = blockIdx.y, c = blockIdx.x;
    const int hw = p.H * p.W;
    if (threadIdx.x < K4) {
        if (threadIdx.x < p.K) {
            cb[threadIdx.x] = kph_bwd_consts(hw, p.K, p.var_mode, p.clip, (long long)n * p.K + threadIdx.x, mean, aux, dmean, dvar);
        } else {   // pad channels: p = exp(0 - 0) / 1, dp = 0 -> gradient 0
            KphB z = {0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            cb[threadIdx.x] = z;
        }
        if (PHASE == 2) {
            float t = 0.f;
            for (int cc = 0; cc < p.C; ++cc) t += p.part1[((long long)n * p.C + cc) * K4 + threadIdx.x];
            stot[threadIdx.x] = t;
        }
    }
    __syncthreads();
    const int i0 = c * p.chunk, i1 = min(hw, i0 + p.chunk);
    const float* base = p.logits + (long long)n * hw * p.ld;
    float acc[K4];
#pragma unroll
    for (int j = 0; j < K4; ++j) acc[j] = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += KPH_THREADS) {
        const float gx = grid_coord(i % p.W, p.W), gy = grid_coord(i / p.W, p.H);
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const float4 v = ldg4(base + (long long)i * p.ld + 4 * q);
            float o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const KphB& k = cb[4 * q + j];
                const float pr = expf(f4get(v, j) * p.invT - k.m) / k.se;
                const float dx = gx - k.mx, dy = gy - k.my;
                const float dp = gx * k.amx + gy * k.amy + k.g00 * dx * dx + k.gxy * dx * dy + k.g11 * dy * dy;
                if (PHASE == 1) acc[4 * q + j] += pr * dp;
                else o4[j] = (4 * q + j < p.K) ? p.invT * pr * (dp - stot[4 * q + j]) : 0.f;
            }
            if (PHASE == 2) st4(dlogits + ((long long)n * hw + i) * p.ld + 4 * q, make_float4(o4[0], o4[1], o4[2], o4[3]));
        }
        if (PHASE == 2) {
            for (int q = KV; 4 * q < p.ld; ++q)   // (ld > K4 never happens with ld = pad4(K); kept for safety)
                st4(dlogits + ((long long)n * hw + i) * p.ld + 4 * q, f4zero());
        }
    }
    if (PHASE == 1) {
        __shared__ float outv[K4];
        kph_reduce<K4, false>(acc, red, outv);
        if (threadIdx.x < K4) p.part1[((long long)n * p.C + c) * K4 + threadIdx.x] = outv[threadIdx.x];
    }
}

static void kph_fill(KphP& p, const float* logits, int N, int H, int W, int K, int ld, float invT, int var_mode,
                     float clip, float* scratch) {
    const int K4 = (K + 3) & ~3;
    p.logits = logits; p.H = H; p.W = W; p.K = K; p.ld = ld; p.invT = invT; p.var_mode = var_mode; p.clip = clip;
    p.C = kph_chunks(H * W);
    p.chunk = (H * W + p.C - 1) / p.C;
    p.part1 = scratch;
    p.part2 = p.part1 + (long long)N * p.C * K4 * 4;
    p.gsum = p.part2 + (long long)N * p.C * K4 * 3;
    p.counter = reinterpret_cast<unsigned*>(p.gsum + (long long)N * p.C * 2);
}

MK_EXPORT int mk_kp_head_fwd(const float* logits, int N, int H, int W, int K, int ld, float inv_temperature,
                             int var_mode, float clip, float* mean, float* var, float* aux, float* scratch,
                             void* stream) {
    MK_REQUIRE(var_mode == 0 || var_mode == 1, "mk_kp_head_fwd: var_mode");
    if (N * K == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int K4 = (K + 3) & ~3;
    if (scratch && ld == K4 && K4 <= 16 && N <= 65535) {
        KphP p;
        kph_fill(p, logits, N, H, W, K, ld, inv_temperature, var_mode, clip, scratch);
        dim3 grid((unsigned)p.C, (unsigned)N);
        switch (K4 / 4) {
            case 1: k_kph_fwd1<1><<<grid, KPH_THREADS, 0, st>>>(p); k_kph_fwd2<1><<<grid, KPH_THREADS, 0, st>>>(p, mean, var, aux); break;
            case 2: k_kph_fwd1<2><<<grid, KPH_THREADS, 0, st>>>(p); k_kph_fwd2<2><<<grid, KPH_THREADS, 0, st>>>(p, mean, var, aux); break;
            case 3: k_kph_fwd1<3><<<grid, KPH_THREADS, 0, st>>>(p); k_kph_fwd2<3><<<grid, KPH_THREADS, 0, st>>>(p, mean, var, aux); break;
            default: k_kph_fwd1<4><<<grid, KPH_THREADS, 0, st>>>(p); k_kph_fwd2<4><<<grid, KPH_THREADS, 0, st>>>(p, mean, var, aux); break;
        }
        return mk_check_launch("mk_kp_head_fwd(chunked)");
    }
    k_kp_head_fwd<<<N * K, 256, 0, st>>>(logits, H, W, K, ld, inv_temperature, var_mode, clip, mean, var, aux);
    return mk_check_launch("mk_kp_head_fwd");
}

MK_EXPORT int mk_kp_head_bwd(const float* logits, int N, int H, int W, int K, int ld, float inv_temperature,
                             int var_mode, float clip, const float* mean, const float* aux, const float* dmean,
                             const float* dvar, float* dlogits, float* scratch, void* stream) {
    if (N * K == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int K4 = (K + 3) & ~3;
    if (scratch && ld == K4 && K4 <= 16 && N <= 65535) {
        KphP p;
        kph_fill(p, logits, N, H, W, K, ld, inv_temperature, var_mode, clip, scratch);
        dim3 grid((unsigned)p.C, (unsigned)N);
#define KPH_BWD(KV)                                                                                              \
    do {                                                                                                         \
        k_kph_bwd<KV, 1><<<grid, KPH_THREADS, 0, st>>>(p, mean, aux, dmean, dvar, dlogits);                      \
        k_kph_bwd<KV, 2><<<grid, KPH_THREADS, 0, st>>>(p, mean, aux, dmean, dvar, dlogits);                      \
    } while (0)
        switch (K4 / 4) {
            case 1: KPH_BWD(1); break;
            case 2: KPH_BWD(2); break;
            case 3: KPH_BWD(3); break;
            default: KPH_BWD(4); break;
        }
#undef KPH_BWD
        return mk_check_launch("mk_kp_head_bwd(chunked)");
    }
    if (ld != K) {
        cudaError_t e = cudaMemsetAsync(dlogits, 0, sizeof(float) * (size_t)N * H * W * ld, st);
        if (e != cudaSuccess) { mk_set_error("mk_kp_head_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    k_kp_head_bwd<<<N * K, 256, 0, st>>>(logits, H, W, K, ld, inv_temperature, var_mode, clip, mean, aux, dmean, dvar,
                                         dlogits);
    return mk_check_launch("mk_kp_head_bwd");
}

// ================================================================================================ gaussians
struct KpGauss {  // one keypoint's gaussian in registers
    float mx, my;
    float a00, a01, a10, a11;  // inverse covariance (matrix mode) or 1/var on the diagonal
};

__device__ __forceinline__ KpGauss load_gauss(const float* __restrict__ mean, const float* __restrict__ var,
                                              long long idx, int var_mode, float const_var) {
    KpGauss g;
    g.mx = mean[idx * 2]; g.my = mean[idx * 2 + 1];
    if (var_mode == 0) {
        float a = var[idx * 4], b = var[idx * 4 + 1], c = var[idx * 4 + 2], d = var[idx * 4 + 3];
        float det = a * d - b * c;
        g.a00 = d / det; g.a01 = -b / det; g.a10 = -c / det; g.a11 = a / det;
    } else {
        float v = var_mode == 1 ? var[idx] : const_var;
        g.a00 = g.a11 = 1.f / v;
        g.a01 = g.a10 = 0.f;
    }
    return g;
}
__device__ __forceinline__ float gauss_eval(const KpGauss& g, float x, float y) {
    float dx = x - g.mx, dy = y - g.my;
    float q = (dx * g.a00 + dy * g.a10) * dx + (dx * g.a01 + dy * g.a11) * dy;
    return expf(-0.5f * q);
}

// heat_sums[0][n][k] = sum over pixels of the driving gaussian, heat_sums[1][n][k] of the source gaussian
__global__ void __launch_bounds__(256) k_kp_heat_sums(const float* __restrict__ kd_mean, const float* __restrict__ kd_var,
                                                      const float* __restrict__ ks_mean, const float* __restrict__ ks_var,
                                                      int NK, int d, int K, int h, int w, int var_mode, float const_var,
                                                      float* __restrict__ sums) {
    __shared__ float red[32];
    const int which = blockIdx.x / NK, idx = blockIdx.x % NK;
    const int n = idx / K, k = idx % K;
    KpGauss g = which == 0 ? load_gauss(kd_mean, kd_var, idx, var_mode, const_var)
                           : load_gauss(ks_mean, ks_var, (long long)(n / d) * K + k, var_mode, const_var);
    float s[1] = {0.f};
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) s[0] += gauss_eval(g, grid_coord(i % w, w), grid_coord(i / w, h));
    block_sum<1>(s, red);
    if (threadIdx.x == 0) sums[blockIdx.x] = s[0];
}

MK_EXPORT int mk_kp_heat_sums(const float* kd_mean, const float* kd_var, const float* ks_mean, const float* ks_var,
                              int B, int d, int K, int h, int w, int var_mode, float const_var, float* heat_sums,
                              void* stream) {
    const int NK = B * d * K;
    if (NK == 0) return 0;
    k_kp_heat_sums<<<2 * NK, 256, 0, (cudaStream_t)stream>>>(kd_mean, kd_var, ks_mean, ks_var, NK, d, K, h, w, var_mode,
                                                             const_var, heat_sums);
    return mk_check_launch("mk_kp_heat_sums");
}

// ================================================================================================ movement embedding
enum { F_HEAT = 1, F_DIFF = 2, F_DEFORMED = 4, F_BG = 8, F_HEATDIFF = 16 };

struct EmbP {
    const float* src; int lds, C;
    const float* kd_mean; const float* kd_var; const float* ks_mean; const float* ks_var;
    int B, d, K, h, w, flags, var_mode; float const_var, norm_const; const float* heat_sums;
    int slots, F;
};

// bilinear sample of C channels of the (translated) source image at normalised (x,y); optionally the x/y derivative
// dotted with an upstream gradient
__device__ __forceinline__ void sample_src(const EmbP& p, long long b, float x, float y, float* out3,
                                           const float* gout, float* gxy) {
    float ix = ((x + 1.f) / 2.f) * (float)(p.w - 1), iy = ((y + 1.f) / 2.f) * (float)(p.h - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const float* s = p.src + b * (long long)p.h * p.w * p.lds;
    bool xin0 = x0 >= 0 && x0 < p.w, xin1 = x0 + 1 >= 0 && x0 + 1 < p.w;
    bool yin0 = y0 >= 0 && y0 < p.h, yin1 = y0 + 1 >= 0 && y0 + 1 < p.h;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < p.C; ++c) {
        float nw = (yin0 && xin0) ? s[((long long)y0 * p.w + x0) * p.lds + c] : 0.f;
        float ne = (yin0 && xin1) ? s[((long long)y0 * p.w + x0 + 1) * p.lds + c] : 0.f;
        float sw = (yin1 && xin0) ? s[((long long)(y0 + 1) * p.w + x0) * p.lds + c] : 0.f;
        float se = (yin1 && xin1) ? s[((long long)(y0 + 1) * p.w + x0 + 1) * p.lds + c] : 0.f;
        if (out3) out3[c] = nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1);
        if (gout) {
            float g = gout[c];
            gix += g * (-nw * wy0 + ne * wy0 - sw * wy1 + se * wy1);
            giy += g * (-nw * wx0 - ne * wx1 + sw * wx0 + se * wx1);
        }
    }
    if (gxy) {
        gxy[0] = gix * 0.5f * (float)(p.w - 1);
        gxy[1] = giy * 0.5f * (float)(p.h - 1);
    }
}

__global__ void __launch_bounds__(256) k_movement_embed_fwd(const EmbP p, float* __restrict__ out, int Cout_p, int ldo,
                                                            long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(i % p.slots);
        const long long op = i / p.slots;
        const int x = (int)(op % p.w);
        const long long t = op / p.w;
        const int y = (int)(t % p.h);
        const long long n = t / p.h;
        const long long b = n / p.d;
        const bool bg = (p.flags & F_BG) != 0;
        const int k = slot - (bg ? 1 : 0);  // -1 for the background slot
        float* o = out + op * ldo + slot * p.F;
        const float gx = grid_coord(x, p.w), gy = grid_coord(y, p.h);
        int f = 0;
        if (p.flags & F_HEAT) {
            float v = 0.f;
            if (k >= 0) {
                const long long idd = n * p.K + k, ids = b * p.K + k;
                float hd = gauss_eval(load_gauss(p.kd_mean, p.kd_var, idd, p.var_mode, p.const_var), gx, gy);
                hd = p.norm_const > 0.f ? hd / p.norm_const : hd / p.heat_sums[idd];
                v = hd;
                if (p.flags & F_HEATDIFF) {
                    float hs = gauss_eval(load_gauss(p.ks_mean, p.ks_var, ids, p.var_mode, p.const_var), gx, gy);
                    hs = p.norm_const > 0.f ? hs / p.norm_const
                                            : hs / p.heat_sums[(long long)p.B * p.d * p.K + idd];
                    v = hd - hs;
                }
            }
            o[f++] = v;
        }
        float sx = 0.f, sy = 0.f;
        if (k >= 0 && (p.flags & (F_DIFF | F_DEFORMED))) {
            sx = p.ks_mean[(b * p.K + k) * 2] - p.kd_mean[(n * p.K + k) * 2];
            sy = p.ks_mean[(b * p.K + k) * 2 + 1] - p.kd_mean[(n * p.K + k) * 2 + 1];
        }
        if (p.flags & F_DIFF) {
            o[f++] = sx;
            o[f++] = sy;
        }
        if (p.flags & F_DEFORMED) {
            float tmp[8];
            sample_src(p, b, gx + sx, gy + sy, tmp, nullptr, nullptr);
            for (int c = 0; c < p.C; ++c) o[f++] = tmp[c];
        }
        if (slot == 0)
            for (int c = p.slots * p.F; c < Cout_p; ++c) out[op * ldo + c] = 0.f;
    }
}

static int fill_embp(EmbP& p, const float* src, int lds, int C, const float* kd_mean, const float* kd_var,
                     const float* ks_mean, const float* ks_var, int B, int d, int K, int h, int w, int flags,
                     int var_mode, float const_var, float norm_const, const float* heat_sums) {
    p.src = src; p.lds = lds; p.C = C; p.kd_mean = kd_mean; p.kd_var = kd_var; p.ks_mean = ks_mean; p.ks_var = ks_var;
    p.B = B; p.d = d; p.K = K; p.h = h; p.w = w; p.flags = flags; p.var_mode = var_mode; p.const_var = const_var;
    p.norm_const = norm_const; p.heat_sums = heat_sums;
    p.slots = K + ((flags & F_BG) ? 1 : 0);
    p.F = ((flags & F_HEAT) ? 1 : 0) + ((flags & F_DIFF) ? 2 : 0) + ((flags & F_DEFORMED) ? C : 0);
    MK_REQUIRE(p.F > 0, "movement_embed: no features selected");
    MK_REQUIRE(!(flags & F_DEFORMED) || (src && C <= 8), "movement_embed: deformed source needs src, C <= 8");
    MK_REQUIRE(norm_const > 0.f || heat_sums || !(flags & F_HEAT), "movement_embed: 'sum' norm needs heat_sums");
    MK_REQUIRE(var_mode != 0 || (kd_var && ks_var) || !(flags & F_HEAT), "movement_embed: matrix mode needs var");
    return 0;
}

MK_EXPORT int mk_movement_embed_fwd(const float* src, int lds, int C, const float* kd_mean, const float* kd_var,
                                    const float* ks_mean, const float* ks_var, int B, int d, int K, int h, int w,
                                    int flags, int var_mode, float const_var, float norm_const, const float* heat_sums,
                                    float* out, int Cout_p, int ldo, void* stream) {
    EmbP p;
    int rc = fill_embp(p, src, lds, C, kd_mean, kd_var, ks_mean, ks_var, B, d, K, h, w, flags, var_mode, const_var,
                       norm_const, heat_sums);
    if (rc) return rc;
    MK_REQUIRE(p.slots * p.F <= Cout_p && Cout_p <= ldo, "movement_embed_fwd: output too narrow");
    const long long total = (long long)B * d * h * w * p.slots;
    if (total == 0) return 0;
    long long blocks = mk_cdiv(total, 256);
    const long long cap = 16LL * mk_num_sms();
    if (blocks > cap) blocks = cap;
    k_movement_embed_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, out, Cout_p, ldo, total);
    return mk_check_launch("mk_movement_embed_fwd");
}

// backward: one block per (frame n, keypoint k); 14 block-reduced sums, closed-form chain in thread 0.  1024 threads:
// the B*d*K blocks (80 at taichi@256) leave half the SMs idle and each walks h*w strided pixels - the block size is the
// only parallelism this kernel has (256 threads: 225 us per launch at 256x256).
__global__ void __launch_bounds__(1024) k_movement_embed_bwd(const EmbP p, const float* __restrict__ dout, int ldo,
                                                            float* __restrict__ d_kd_mean, float* __restrict__ d_kd_var,
                                                            float* __restrict__ d_ks_mean,
                                                            float* __restrict__ d_ks_var) {
    __shared__ float red[14 * 32];
    const int n = blockIdx.x / p.K, k = blockIdx.x % p.K;
    const long long b = n / p.d;
    const bool bg = (p.flags & F_BG) != 0;
    const int slot = k + (bg ? 1 : 0);
    const long long idd = (long long)n * p.K + k, ids = b * p.K + k;
    const int hw = p.h * p.w;
    const float* g0 = dout + (long long)n * hw * ldo + slot * p.F;
    const bool heat = (p.flags & F_HEAT) != 0, hdiff = (p.flags & F_HEATDIFF) != 0;
    KpGauss gd = {}, gs = {};
    float Sd = 1.f, Ss = 1.f, Td = 0.f, Ts = 0.f;
    if (heat) {
        gd = load_gauss(p.kd_mean, p.kd_var, idd, p.var_mode, p.const_var);
        if (hdiff) gs = load_gauss(p.ks_mean, p.ks_var, ids, p.var_mode, p.const_var);
        if (p.norm_const > 0.f) {
            Sd = Ss = p.norm_const;
        } else {
            Sd = p.heat_sums[idd];
            Ss = p.heat_sums[(long long)p.B * p.d * p.K + idd];
            float t2[2] = {0.f, 0.f};
            for (int i = threadIdx.x; i < hw; i += blockDim.x) {
                float g = g0[(long long)i * ldo];
                float x = grid_coord(i % p.w, p.w), y = grid_coord(i / p.w, p.h);
                t2[0] += g * gauss_eval(gd, x, y);
                if (hdiff) t2[1] += g * gauss_eval(gs, x, y);
            }
            block_sum<2>(t2, red);
            Td = t2[0]; Ts = t2[1];
        }
    }
    float sx = 0.f, sy = 0.f;
    if (p.flags & (F_DIFF | F_DEFORMED)) {
        sx = p.ks_mean[ids * 2] - p.kd_mean[idd * 2];
        sy = p.ks_mean[ids * 2 + 1] - p.kd_mean[idd * 2 + 1];
    }
    // 0,1: dmu_d  2..5: dA_d  6,7: dmu_s  8..11: dA_s  12,13: dshift
    float acc[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) acc[j] = 0.f;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        const float* g = g0 + (long long)i * ldo;
        const float x = grid_coord(i % p.w, p.w), y = grid_coord(i / p.w, p.h);
        int f = 0;
        if (heat) {
            const float G = g[f++];
            {
                float e = gauss_eval(gd, x, y);
                float wgt = p.norm_const > 0.f ? G / Sd : (G - Td / Sd) / Sd;  // dL/de
                float we = wgt * e;
                float dx = x - gd.mx, dy = y - gd.my;
                // dq/dmu = -(A + A^T) delta ; de = -0.5 e dq
                acc[0] += 0.5f * we * (2.f * gd.a00 * dx + (gd.a01 + gd.a10) * dy);
                acc[1] += 0.5f * we * ((gd.a01 + gd.a10) * dx + 2.f * gd.a11 * dy);
                acc[2] += -0.5f * we * dx * dx; acc[3] += -0.5f * we * dx * dy;
                acc[4] += -0.5f * we * dy * dx; acc[5] += -0.5f * we * dy * dy;
            }
            if (hdiff) {
                float e = gauss_eval(gs, x, y);
                float wgt = p.norm_const > 0.f ? -G / Ss : -(G - Ts / Ss) / Ss;
                float we = wgt * e;
                float dx = x - gs.mx, dy = y - gs.my;
                acc[6] += 0.5f * we * (2.f * gs.a00 * dx + (gs.a01 + gs.a10) * dy);
                acc[7] += 0.5f * we * ((gs.a01 + gs.a10) * dx + 2.f * gs.a11 * dy);
                acc[8] += -0.5f * we * dx * dx; acc[9] += -0.5f * we * dx * dy;
                acc[10] += -0.5f * we * dy * dx; acc[11] += -0.5f * we * dy * dy;
            }
        }
        if (p.flags & F_DIFF) {
            acc[12] += g[f++];
            acc[13] += g[f++];
        }
        if (p.flags & F_DEFORMED) {
            float gxy[2];
            sample_src(p, b, x + sx, y + sy, nullptr, g + f, gxy);
            acc[12] += gxy[0];
            acc[13] += gxy[1];
        }
    }
    block_sum<14>(acc, red);
    if (threadIdx.x != 0) return;
    // chain dA -> dSigma (matrix) or d(1/v) -> dv (single):  dSigma = -A^T G A^T
    auto chain_var = [&](const KpGauss& g, const float* G, float* dvar_out, long long idx, bool atomic) {
        if (p.var_mode == 0) {
            // M = A^T G ; R = M A^T ; dSigma = -R
            float m00 = g.a00 * G[0] + g.a10 * G[2], m01 = g.a00 * G[1] + g.a10 * G[3];
            float m10 = g.a01 * G[0] + g.a11 * G[2], m11 = g.a01 * G[1] + g.a11 * G[3];
            float r00 = m00 * g.a00 + m01 * g.a01, r01 = m00 * g.a10 + m01 * g.a11;
            float r10 = m10 * g.a00 + m11 * g.a01, r11 = m10 * g.a10 + m11 * g.a11;
            float* q = dvar_out + idx * 4;
            if (atomic) { atomicAdd(q, -r00); atomicAdd(q + 1, -r01); atomicAdd(q + 2, -r10); atomicAdd(q + 3, -r11); }
            else { q[0] = -r00; q[1] = -r01; q[2] = -r10; q[3] = -r11; }
        } else if (p.var_mode == 1) {
            // A = I / v : dL/dv = -(G00 + G11) / v^2
            float dv = -(G[0] + G[3]) * g.a00 * g.a00;
            if (atomic) atomicAdd(dvar_out + idx, dv); else dvar_out[idx] = dv;
        }
    };
    d_kd_mean[idd * 2] = acc[0] - acc[12];
    d_kd_mean[idd * 2 + 1] = acc[1] - acc[13];
    atomicAdd(d_ks_mean + ids * 2, acc[6] + acc[12]);
    atomicAdd(d_ks_mean + ids * 2 + 1, acc[7] + acc[13]);
    if (heat) {
        if (d_kd_var) chain_var(gd, acc + 2, d_kd_var, idd, false);
        if (hdiff && d_ks_var) chain_var(gs, acc + 8, d_ks_var, ids, true);
    }
}

MK_EXPORT int mk_movement_embed_bwd(const float* src, int lds, int C, const float* kd_mean, const float* kd_var,
                                    const float* ks_mean, const float* ks_var, int B, int d, int K, int h, int w,
                                    int flags, int var_mode, float const_var, float norm_const, const float* heat_sums,
                                    const float* dout, int ldo, float* d_kd_mean, float* d_kd_var, float* d_ks_mean,
                                    float* d_ks_var, void* stream) {
    EmbP p;
    int rc = fill_embp(p, src, lds, C, kd_mean, kd_var, ks_mean, ks_var, B, d, K, h, w, flags, var_mode, const_var,
                       norm_const, heat_sums);
    if (rc) return rc;
    if (B * d * K == 0) return 0;
    k_movement_embed_bwd<<<B * d * K, 1024, 0, (cudaStream_t)stream>>>(p, dout, ldo, d_kd_mean, d_kd_var, d_ks_mean,
                                                                      d_ks_var);
    return mk_check_launch("mk_movement_embed_bwd");
}

// ================================================================================================ dense-motion head
constexpr int MAXS = 32;  // max mask slots (K+1)

__global__ void __launch_bounds__(256) k_flow_head_fwd(const float* __restrict__ pred, int ld,
                                                       const float* __restrict__ kd_mean,
                                                       const float* __restrict__ ks_mean, int d, int K, int h, int w,
                                                       int use_mask, int use_corr, float* __restrict__ deform,
                                                       long long total) {
    for (long long op = (long long)blockIdx.x * blockDim.x + threadIdx.x; op < total;
         op += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(op % w);
        const long long t = op / w;
        const int y = (int)(t % h);
        const long long n = t / h, b = n / d;
        const float* pr = pred + op * ld;
        float fx = 0.f, fy = 0.f;
        int off = 0;
        if (use_mask) {
            const int S = K + 1;
            float l[MAXS];
            float m = -INFINITY;
            for (int s = 0; s < S; ++s) { l[s] = pr[s]; m = fmaxf(m, l[s]); }
            float se = 0.f;
            for (int s = 0; s < S; ++s) { l[s] = expf(l[s] - m); se += l[s]; }
            for (int s = 1; s < S; ++s) {
                float mk = l[s] / se;
                fx += (ks_mean[(b * K + s - 1) * 2] - kd_mean[(n * K + s - 1) * 2]) * mk;
                fy += (ks_mean[(b * K + s - 1) * 2 + 1] - kd_mean[(n * K + s - 1) * 2 + 1]) * mk;
            }
            off = S;
        }
        if (use_corr) { fx += pr[off]; fy += pr[off + 1]; }
        deform[op * 2] = fx + grid_coord(x, w);
        deform[op * 2 + 1] = fy + grid_coord(y, h);
    }
}

MK_EXPORT int mk_flow_head_fwd(const float* pred, int ld, const float* kd_mean, const float* ks_mean, int B, int d,
                               int K, int h, int w, int use_mask, int use_correction, float* deform, void* stream) {
    MK_REQUIRE(K + 1 <= MAXS, "mk_flow_head: too many keypoints");
    const long long total = (long long)B * d * h * w;
    if (total == 0) return 0;
    long long blocks = mk_cdiv(total, 256);
    k_flow_head_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, ld, kd_mean, ks_mean, d, K, h, w, use_mask,
                                                                        use_correction, deform, total);
    return mk_check_launch("mk_flow_head_fwd");
}

// backward: grid (pixel blocks, frame n) so the per-keypoint shift gradients reduce within a block
__global__ void __launch_bounds__(256) k_flow_head_bwd(const float* __restrict__ pred, int ld,
                                                       const float* __restrict__ kd_mean,
                                                       const float* __restrict__ ks_mean, int d, int K, int h, int w,
                                                       int use_mask, int use_corr, const float* __restrict__ ddeform,
                                                       float* __restrict__ dpred, int P, float* __restrict__ d_kd_mean,
                                                       float* __restrict__ d_ks_mean) {
    __shared__ float red[2 * 32];
    __shared__ float sh_shift[MAXS * 2];
    const int n = blockIdx.y;
    const long long b = n / d;
    const int hw = h * w;
    const int S = K + 1;
    if (use_mask) {
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            sh_shift[s * 2] = s ? ks_mean[(b * K + s - 1) * 2] - kd_mean[((long long)n * K + s - 1) * 2] : 0.f;
            sh_shift[s * 2 + 1] = s ? ks_mean[(b * K + s - 1) * 2 + 1] - kd_mean[((long long)n * K + s - 1) * 2 + 1] : 0.f;
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < hw;
    const long long op = (long long)n * hw + (valid ? i : 0);
    float dfx = 0.f, dfy = 0.f;
    float mk[MAXS];
    if (valid) {
        dfx = ddeform[op * 2];
        dfy = ddeform[op * 2 + 1];
    }
    float* dp = dpred + op * ld;
    int off = 0;
    if (use_mask) {
        const float* pr = pred + op * ld;
        float m = -INFINITY, se = 0.f;
        for (int s = 0; s < S; ++s) { mk[s] = valid ? pr[s] : 0.f; m = fmaxf(m, mk[s]); }
        for (int s = 0; s < S; ++s) { mk[s] = expf(mk[s] - m); se += mk[s]; }
        float dot = 0.f;
        for (int s = 0; s < S; ++s) {
            mk[s] /= se;
            dot += mk[s] * (sh_shift[s * 2] * dfx + sh_shift[s * 2 + 1] * dfy);
        }
        if (valid)
            for (int s = 0; s < S; ++s) dp[s] = mk[s] * (sh_shift[s * 2] * dfx + sh_shift[s * 2 + 1] * dfy - dot);
        off = S;
    }
    if (valid) {
        if (use_corr) { dp[off] = dfx; dp[off + 1] = dfy; off += 2; }
        for (int c = off; c < P; ++c) dp[c] = 0.f;
    }
    if (use_mask) {
        for (int s = 1; s < S; ++s) {
            float v[2] = {valid ? mk[s] * dfx : 0.f, valid ? mk[s] * dfy : 0.f};
            block_sum<2>(v, red);
            if (threadIdx.x == 0) {
                atomicAdd(d_kd_mean + ((long long)n * K + s - 1) * 2, -v[0]);
                atomicAdd(d_kd_mean + ((long long)n * K + s - 1) * 2 + 1, -v[1]);
                atomicAdd(d_ks_mean + (b * K + s - 1) * 2, v[0]);
                atomicAdd(d_ks_mean + (b * K + s - 1) * 2 + 1, v[1]);
            }
        }
    }
}

MK_EXPORT int mk_flow_head_bwd(const float* pred, int ld, const float* kd_mean, const float* ks_mean, int B, int d,
                               int K, int h, int w, int use_mask, int use_correction, const float* ddeform,
                               float* dpred, float* d_kd_mean, float* d_ks_mean, void* stream) {
    MK_REQUIRE(K + 1 <= MAXS, "mk_flow_head: too many keypoints");
    const int N = B * d;
    if (N * h * w == 0) return 0;
    dim3 grid((unsigned)mk_cdiv((long long)h * w, 256), (unsigned)N);
    k_flow_head_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, ld, kd_mean, ks_mean, d, K, h, w, use_mask,
                                                            use_correction, ddeform, dpred, ld, d_kd_mean, d_ks_mean);
    return mk_check_launch("mk_flow_head_bwd");
}
