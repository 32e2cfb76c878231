// Per-sample loss means (modules/losses.py:4-24) over logical 5-D (B,C,D,H,W) operands given by element strides, so
// the reference-layout NCDHW inputs and the internal NHWC feature maps (exposed as permuted views) mix freely.
// HBM-bound reductions: grid (blocks per sample, B), block partial -> one atomic per block.
#include "common.cuh"
#include "../../include/monkey_b200.h"

struct LossP {
    const float* a; const float* b;
    long long sa[5], sb[5];
    int C, D, H, W, c_fast;  // c_fast: iterate channels fastest (NHWC-backed operands)
    long long per_sample;
    float scale;  // weight / per_sample
};

__device__ __forceinline__ void loss_offsets(const LossP& p, long long s, long long j, long long& oa, long long& ob) {
    int c, d, h, w;
    if (p.c_fast) {
        c = (int)(j % p.C); j /= p.C;
        w = (int)(j % p.W); j /= p.W;
        h = (int)(j % p.H); d = (int)(j / p.H);
    } else {
        w = (int)(j % p.W); j /= p.W;
        h = (int)(j % p.H); j /= p.H;
        d = (int)(j % p.D); c = (int)(j / p.D);
    }
    oa = s * p.sa[0] + c * p.sa[1] + d * p.sa[2] + h * p.sa[3] + w * p.sa[4];
    ob = s * p.sb[0] + c * p.sb[1] + d * p.sb[2] + h * p.sb[3] + w * p.sb[4];
}

template <int KIND>
__global__ void __launch_bounds__(256) k_loss_fwd(const LossP p, float* __restrict__ out) {
    __shared__ float red[32];
    const long long s = blockIdx.y;
    float acc[1] = {0.f};
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < p.per_sample;
         j += (long long)gridDim.x * blockDim.x) {
        long long oa, ob;
        loss_offsets(p, s, j, oa, ob);
        float a = p.a[oa];
        if (KIND == 0) acc[0] += fabsf(a - p.b[ob]);
        else if (KIND == 1) acc[0] += (1.f - a) * (1.f - a);
        else { float b = p.b[ob]; acc[0] += (1.f - a) * (1.f - a) + b * b; }
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + s, acc[0] * p.scale);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_loss_bwd(const LossP p, const float* __restrict__ gout, float* __restrict__ da,
                                                  float* __restrict__ db) {
    const long long s = blockIdx.y;
    const float g = gout[s] * p.scale;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < p.per_sample;
         j += (long long)gridDim.x * blockDim.x) {
        long long oa, ob;
        loss_offsets(p, s, j, oa, ob);
        float a = p.a[oa];
        if (KIND == 0) {
            float df = a - p.b[ob];
            float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
            if (da) da[oa] = sg * g;
            if (db) db[ob] = -sg * g;
        } else if (KIND == 1) {
            if (da) da[oa] = -2.f * (1.f - a) * g;
        } else {
            if (da) da[oa] = -2.f * (1.f - a) * g;
            if (db) db[ob] = 2.f * p.b[ob] * g;
        }
    }
}

static int fill_lossp(LossP& p, int kind, const float* a, const long long* sa, const float* b, const long long* sb,
                      int B, int C, int D, int H, int W, float weight) {
    MK_REQUIRE(kind >= 0 && kind <= 2, "mk_loss: bad kind");
    MK_REQUIRE(kind == 1 || b, "mk_loss: operand b required");
    p.a = a; p.b = b ? b : a;
    for (int i = 0; i < 5; ++i) { p.sa[i] = sa[i]; p.sb[i] = b ? sb[i] : sa[i]; }
    p.C = C; p.D = D; p.H = H; p.W = W;
    p.c_fast = (sa[1] == 1 && C > 1) ? 1 : 0;
    p.per_sample = (long long)C * D * H * W;
    p.scale = weight / (float)p.per_sample;
    return 0;
}

MK_EXPORT int mk_loss_fwd(int kind, const float* a, const long long* stride_a, const float* b,
                          const long long* stride_b, int B, int C, int D, int H, int W, float weight, float* out,
                          void* stream) {
    LossP p;
    int rc = fill_lossp(p, kind, a, stride_a, b, stride_b, B, C, D, H, W, weight);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B, st);
    if (e != cudaSuccess) { mk_set_error("mk_loss_fwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (B == 0 || p.per_sample == 0) return 0;
    long long bx = mk_cdiv(p.per_sample, 256 * 8);
    long long cap = mk_cdiv(4LL * mk_num_sms(), B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)B);
    if (kind == 0) k_loss_fwd<0><<<grid, 256, 0, st>>>(p, out);
    else if (kind == 1) k_loss_fwd<1><<<grid, 256, 0, st>>>(p, out);
    else k_loss_fwd<2><<<grid, 256, 0, st>>>(p, out);
    return mk_check_launch("mk_loss_fwd");
}

MK_EXPORT int mk_loss_bwd(int kind, const float* a, const long long* stride_a, const float* b,
                          const long long* stride_b, int B, int C, int D, int H, int W, float weight,
                          const float* gout, float* da, float* db, void* stream) {
    LossP p;
    int rc = fill_lossp(p, kind, a, stride_a, b, stride_b, B, C, D, H, W, weight);
    if (rc) return rc;
    if (B == 0 || p.per_sample == 0) return 0;
    long long bx = mk_cdiv(p.per_sample, 256 * 4);
    long long cap = mk_cdiv(8LL * mk_num_sms(), B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 0) k_loss_bwd<0><<<grid, 256, 0, st>>>(p, gout, da, db);
    else if (kind == 1) k_loss_bwd<1><<<grid, 256, 0, st>>>(p, gout, da, db);
    else k_loss_bwd<2><<<grid, 256, 0, st>>>(p, gout, da, db);
    return mk_check_launch("mk_loss_bwd");
}
