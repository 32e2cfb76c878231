// Batch / instance normalisation + activation + 2x2 average pool, forward and backward.  All HBM-bound:
// float4 over the channel axis (NHWC), grid-stride over pixels, grids sized in multiples of the SM count.
//
// Training-mode batch norm is a two-phase op: column statistics -> [2][Cp] sums (the only thing that
// crosses GPUs: one all-reduce per layer per direction) -> finalize -> apply.  Instance norm is the same code with
// one statistics group per frame.  params layout everywhere: [groups][4][Cp] = mean, invstd, scale, shift.
#include "common.cuh"
#include "../../include/monkey_b200.h"

// ------------------------------------------------------------------------------------------------ column statistics
// MODE 0: sum x, sum x^2.   MODE 1 (backward): sum dz, sum dz*xhat  with dz = act'(z)*unpool(dout).
struct StatP {
    const float* x; int ldx; int N, H, W, Cp; long long hw;
    const float* dout; int ldd; int Hp, Wp; const float* params; int per_frame; float slope; int pool;
    float* sums;
};

template <int MODE>
__global__ void __launch_bounds__(256) k_colstats(const StatP p, int cv, int cvb, int rows) {
    __shared__ float4 s0[256];
    __shared__ float4 s1[256];
    const int tid = threadIdx.x;
    const int tv = tid % cvb, prow = tid / cvb;
    const int vec = blockIdx.y * cvb + tv;
    const int g = blockIdx.z;  // statistics group (frame) when per_frame
    const bool active = prow < rows && vec < cv;
    float4 a0 = f4zero(), a1 = f4zero();
    if (active) {
        const int c = vec * 4;
        const long long pix_total = p.per_frame ? p.hw : (long long)p.N * p.hw;
        const long long pix_base = p.per_frame ? (long long)g * p.hw : 0;
        float4 mean = f4zero(), invstd = f4zero(), scale = make_float4(1.f, 1.f, 1.f, 1.f), shift = f4zero();
        if (MODE == 1 && p.params) {
            const float* pr = p.params + (long long)(p.per_frame ? g : 0) * 4 * p.Cp;
            mean = ldg4(pr + c); invstd = ldg4(pr + p.Cp + c); scale = ldg4(pr + 2 * p.Cp + c); shift = ldg4(pr + 3 * p.Cp + c);
        }
        for (long long q = (long long)blockIdx.x * rows + prow; q < pix_total; q += (long long)gridDim.x * rows) {
            const long long pix = pix_base + q;
            float4 v = ldg4(p.x + pix * p.ldx + c);
            if (MODE == 0) {
                a0 = a0 + v;
                a1 = a1 + v * v;
            } else {
                float4 d;
                if (p.pool) {
                    int w = (int)(pix % p.W);
                    long long t = pix / p.W;
                    int h = (int)(t % p.H);
                    long long n = t / p.H;
                    int hp = h >> 1, wp = w >> 1;
                    if (hp < p.Hp && wp < p.Wp) d = ldg4(p.dout + ((n * p.Hp + hp) * p.Wp + wp) * p.ldd + c) * 0.25f;
                    else d = f4zero();
                } else {
                    d = ldg4(p.dout + pix * p.ldd + c);
                }
                if (p.slope >= 0.f) {
                    float4 z = make_float4(fmaf(v.x, scale.x, shift.x), fmaf(v.y, scale.y, shift.y),
                                           fmaf(v.z, scale.z, shift.z), fmaf(v.w, scale.w, shift.w));
                    d.x = z.x > 0.f ? d.x : d.x * p.slope; d.y = z.y > 0.f ? d.y : d.y * p.slope;
                    d.z = z.z > 0.f ? d.z : d.z * p.slope; d.w = z.w > 0.f ? d.w : d.w * p.slope;
                }
                float4 xh = make_float4((v.x - mean.x) * invstd.x, (v.y - mean.y) * invstd.y,
                                        (v.z - mean.z) * invstd.z, (v.w - mean.w) * invstd.w);
                a0 = a0 + d;
                a1 = a1 + d * xh;
            }
        }
    }
    s0[tid] = a0;
    s1[tid] = a1;
    __syncthreads();
    int span = 1;
    while (span < rows) span <<= 1;
    for (int s = span >> 1; s > 0; s >>= 1) {
        if (active && prow < s && prow + s < rows) {
            s0[tid] = s0[tid] + s0[tid + s * cvb];
            s1[tid] = s1[tid] + s1[tid + s * cvb];
        }
        __syncthreads();
    }
    if (active && prow == 0) {
        float* o = p.sums + (long long)g * 2 * p.Cp + vec * 4;
        float4 r0 = s0[tid], r1 = s1[tid];
        atomicAdd(o + 0, r0.x); atomicAdd(o + 1, r0.y); atomicAdd(o + 2, r0.z); atomicAdd(o + 3, r0.w);
        o += p.Cp;
        atomicAdd(o + 0, r1.x); atomicAdd(o + 1, r1.y); atomicAdd(o + 2, r1.z); atomicAdd(o + 3, r1.w);
    }
}

// Forward statistics of a normalisation layer in DOUBLE precision: var = E[x^2] - E[x]^2 cancels catastrophically in
// fp32 when mean^2 >> var (measured up to 90x on the dense-motion hourglass of shapes.yaml -> 1e-5 relative error of
// the variance; ATen uses Welford for the same reason).  B200 runs FP64 at half the FP32 rate, and this kernel is
// HBM-bound, so the double accumulators are free.  sums[g][2][Cp] doubles.
struct D4 { double x, y, z, w; };
__device__ __forceinline__ void d4acc(D4& a, float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ void d4acc2(D4& a, float4 v) {
    a.x += (double)v.x * v.x; a.y += (double)v.y * v.y; a.z += (double)v.z * v.z; a.w += (double)v.w * v.w;
}
__device__ __forceinline__ void d4add(D4& a, const D4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__global__ void __launch_bounds__(256) k_colstats_f64(const float* __restrict__ x, int ldx, long long pix_total,
                                                      long long hw, int Cp, int per_frame, double* __restrict__ sums,
                                                      int cv, int cvb, int rows) {
    __shared__ D4 s0[256];
    __shared__ D4 s1[256];
    const int tid = threadIdx.x;
    const int tv = tid % cvb, prow = tid / cvb;
    const int vec = blockIdx.y * cvb + tv;
    const int g = blockIdx.z;
    const bool active = prow < rows && vec < cv;
    D4 a0 = {0., 0., 0., 0.}, a1 = {0., 0., 0., 0.};
    if (active) {
        const int c = vec * 4;
        const long long pix_base = per_frame ? (long long)g * hw : 0;
        for (long long q = (long long)blockIdx.x * rows + prow; q < pix_total; q += (long long)gridDim.x * rows) {
            const float4 v = ldg4(x + (pix_base + q) * ldx + c);
            d4acc(a0, v);
            d4acc2(a1, v);
        }
    }
    s0[tid] = a0;
    s1[tid] = a1;
    __syncthreads();
    int span = 1;
    while (span < rows) span <<= 1;
    for (int s = span >> 1; s > 0; s >>= 1) {
        if (active && prow < s && prow + s < rows) {
            d4add(s0[tid], s0[tid + s * cvb]);
            d4add(s1[tid], s1[tid + s * cvb]);
        }
        __syncthreads();
    }
    if (active && prow == 0) {
        double* o = sums + (long long)g * 2 * Cp + vec * 4;
        const D4 r0 = s0[tid], r1 = s1[tid];
        atomicAdd(o + 0, r0.x); atomicAdd(o + 1, r0.y); atomicAdd(o + 2, r0.z); atomicAdd(o + 3, r0.w);
        o += Cp;
        atomicAdd(o + 0, r1.x); atomicAdd(o + 1, r1.y); atomicAdd(o + 2, r1.z); atomicAdd(o + 3, r1.w);
    }
}

template <int MODE>
static int launch_colstats(StatP& p, cudaStream_t st, const char* what) {
    const int cv = p.Cp / 4;
    const int cvb = cv < 64 ? cv : 64;
    const int rows = 256 / cvb;
    const int groups = p.per_frame ? p.N : 1;
    cudaError_t e = cudaMemsetAsync(p.sums, 0, sizeof(float) * 2 * (size_t)p.Cp * groups, st);
    if (e != cudaSuccess) { mk_set_error("%s memset: %s", what, cudaGetErrorString(e)); return (int)e; }
    const long long pix = p.per_frame ? p.hw : (long long)p.N * p.hw;
    if (pix == 0) return 0;
    const int ychunks = (int)mk_cdiv(cv, cvb);
    long long nblk = mk_cdiv(pix, (long long)rows * 8);
    long long cap = mk_cdiv(4LL * mk_num_sms(), (long long)ychunks * groups);
    if (nblk > cap) nblk = cap;
    if (nblk < 1) nblk = 1;
    dim3 grid((unsigned)nblk, (unsigned)ychunks, (unsigned)groups);
    k_colstats<MODE><<<grid, 256, 0, st>>>(p, cv, cvb, rows);
    return mk_check_launch(what);
}

MK_EXPORT int mk_colstats(const float* x, int ld, int N, long long hw, int Cp, int per_frame, float* sums,
                          void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0, "mk_colstats: channels must be x4");
    MK_REQUIRE(!per_frame || N <= 65535, "mk_colstats: too many groups");
    StatP p = {};
    p.x = x; p.ldx = ld; p.N = N; p.hw = hw; p.Cp = Cp; p.per_frame = per_frame; p.sums = sums; p.slope = -1.f;
    return launch_colstats<0>(p, (cudaStream_t)stream, "mk_colstats");
}

MK_EXPORT int mk_colstats_f64(const float* x, int ld, int N, long long hw, int Cp, int per_frame, double* sums,
                              void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0, "mk_colstats_f64: channels must be x4");
    MK_REQUIRE(!per_frame || N <= 65535, "mk_colstats_f64: too many groups");
    cudaStream_t st = (cudaStream_t)stream;
    const int cv = Cp / 4;
    const int cvb = cv < 64 ? cv : 64;
    const int rows = 256 / cvb;
    const int groups = per_frame ? N : 1;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)Cp * groups, st);
    if (e != cudaSuccess) { mk_set_error("mk_colstats_f64 memset: %s", cudaGetErrorString(e)); return (int)e; }
    const long long pix = per_frame ? hw : (long long)N * hw;
    if (pix == 0) return 0;
    const int ychunks = (int)mk_cdiv(cv, cvb);
    long long nblk = mk_cdiv(pix, (long long)rows * 8);
    long long cap = mk_cdiv(4LL * mk_num_sms(), (long long)ychunks * groups);
    if (nblk > cap) nblk = cap;
    if (nblk < 1) nblk = 1;
    dim3 grid((unsigned)nblk, (unsigned)ychunks, (unsigned)groups);
    k_colstats_f64<<<grid, 256, 0, st>>>(x, ld, pix, hw, Cp, per_frame, sums, cv, cvb, rows);
    return mk_check_launch("mk_colstats_f64");
}

MK_EXPORT int mk_norm_bwd_reduce(const float* x, int ldx, const float* dout, int ldd, int N, int H, int W, int Cp,
                                 const float* params, int per_frame, float slope, int pool, float* sums,
                                 void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ldx % 4 == 0 && ldd % 4 == 0, "mk_norm_bwd_reduce: channels must be x4");
    MK_REQUIRE(!per_frame || N <= 65535, "mk_norm_bwd_reduce: too many groups");
    StatP p = {};
    p.x = x; p.ldx = ldx; p.N = N; p.H = H; p.W = W; p.hw = (long long)H * W; p.Cp = Cp;
    p.dout = dout; p.ldd = ldd; p.Hp = pool ? H >> 1 : H; p.Wp = pool ? W >> 1 : W;
    p.params = params; p.per_frame = per_frame; p.slope = slope; p.pool = pool; p.sums = sums;
    return launch_colstats<1>(p, (cudaStream_t)stream, "mk_norm_bwd_reduce");
}

// ------------------------------------------------------------------------------------------------ finalize
__global__ void k_norm_finalize(const double* __restrict__ sums, int groups, int C, int Cp, double count,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                float* running_mean, float* running_var, float momentum, long long* nbt,
                                float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * Cp) return;
    int g = i / Cp, c = i % Cp;
    float* o = out + (long long)g * 4 * Cp;
    if (c >= C) {
        o[c] = 0.f; o[Cp + c] = 0.f; o[2 * Cp + c] = 0.f; o[3 * Cp + c] = 0.f;
        return;
    }
    double s = sums[(long long)g * 2 * Cp + c], ss = sums[(long long)g * 2 * Cp + Cp + c];
    double mean = s / count;
    double var = ss / count - mean * mean;
    if (var < 0.0) var = 0.0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float scale = ga * (float)invstd;
    o[c] = (float)mean;
    o[Cp + c] = (float)invstd;
    o[2 * Cp + c] = scale;
    o[3 * Cp + c] = be - (float)mean * scale;
    if (running_mean && g == 0) {
        double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        if (c == 0 && nbt) *nbt += 1;
    }
}

MK_EXPORT int mk_norm_finalize(const double* sums, int groups, int C, int Cp, double count, const float* gamma,
                               const float* beta, float eps, float* running_mean, float* running_var,
                               float momentum, long long* num_batches_tracked, float* out, void* stream) {
    int total = groups * Cp;
    if (total == 0) return 0;
    k_norm_finalize<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        sums, groups, C, Cp, count, gamma, beta, eps, running_mean, running_var, momentum, num_batches_tracked, out);
    return mk_check_launch("mk_norm_finalize");
}

__global__ void k_norm_eval_params(const float* __restrict__ rm, const float* __restrict__ rv,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int C, int Cp,
                                   float eps, float* __restrict__ o) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c >= C) {
        o[c] = 0.f; o[Cp + c] = 0.f; o[2 * Cp + c] = 0.f; o[3 * Cp + c] = 0.f;
        return;
    }
    float invstd = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
    float scale = (gamma ? gamma[c] : 1.f) * invstd;
    o[c] = rm[c];
    o[Cp + c] = invstd;
    o[2 * Cp + c] = scale;
    o[3 * Cp + c] = (beta ? beta[c] : 0.f) - rm[c] * scale;
}

MK_EXPORT int mk_norm_eval_params(const float* running_mean, const float* running_var, const float* gamma,
                                  const float* beta, int C, int Cp, float eps, float* out, void* stream) {
    k_norm_eval_params<<<(Cp + 127) / 128, 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, gamma, beta, C,
                                                                           Cp, eps, out);
    return mk_check_launch("mk_norm_eval_params");
}

// ------------------------------------------------------------------------------------------------ apply (+act, +pool)
__device__ __forceinline__ float4 affine_act(float4 v, float4 sc, float4 sh, float slope) {
    float4 z = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (slope >= 0.f) {
        z.x = z.x > 0.f ? z.x : z.x * slope; z.y = z.y > 0.f ? z.y : z.y * slope;
        z.z = z.z > 0.f ? z.z : z.z * slope; z.w = z.w > 0.f ? z.w : z.w * slope;
    }
    return z;
}

__global__ void __launch_bounds__(256) k_norm_apply(const float* __restrict__ x, int ldx, int N, int H, int W, int Cp,
                                                    const float* __restrict__ params, int per_frame, float slope,
                                                    int pool, float* __restrict__ out, int ldo, int Ho, int Wo,
                                                    long long total, const FastDiv fcv, const FastDiv fwo,
                                                    const FastDiv fho) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;  // one (output pixel, float4) per thread, 32-bit index math
    if (i >= (unsigned)total) return;
    unsigned cq, wo, ho;
    const unsigned op = fd_divmod(i, fcv, cq);
    const unsigned t = fd_divmod(op, fwo, wo);
    const unsigned n = fd_divmod(t, fho, ho);
    const int c = (int)cq * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = f4zero();
    if (params) {
        const float* pr = params + (long long)(per_frame ? n : 0) * 4 * Cp;
        sc = ldg4(pr + 2 * Cp + c);
        sh = ldg4(pr + 3 * Cp + c);
    }
    float4 r;
    if (pool) {
        const float* b = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * ldx + c;
        const float4 v0 = ldg4(b), v1 = ldg4(b + ldx), v2 = ldg4(b + (long long)W * ldx),
                     v3 = ldg4(b + (long long)W * ldx + ldx);
        r = affine_act(v0, sc, sh, slope) + affine_act(v1, sc, sh, slope) + affine_act(v2, sc, sh, slope) +
            affine_act(v3, sc, sh, slope);
        r = r * 0.25f;
    } else {
        r = affine_act(ldg4(x + (long long)op * ldx + c), sc, sh, slope);
    }
    st4(out + (long long)op * ldo + c, r);
}

MK_EXPORT int mk_norm_apply(const float* x, int ldx, int N, int H, int W, int Cp, const float* params, int per_frame,
                            float slope, int pool, float* out, int ldo, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "mk_norm_apply: channels must be x4");
    const int Ho = pool ? H >> 1 : H, Wo = pool ? W >> 1 : W;
    const long long total = (long long)N * Ho * Wo * (Cp / 4);
    if (total == 0) return 0;
    MK_REQUIRE(total < (1LL << 31), "mk_norm_apply: more than 2^31 work items");
    k_norm_apply<<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, N, H, W, Cp, params, per_frame, slope, pool, out, ldo, Ho, Wo, total, make_fastdiv(Cp / 4),
        make_fastdiv(Wo), make_fastdiv(Ho));
    return mk_check_launch("mk_norm_apply");
}

// ------------------------------------------------------------------------------------------------ backward apply
__global__ void __launch_bounds__(256) k_norm_bwd_apply(const float* __restrict__ x, int ldx,
                                                        const float* __restrict__ dout, int ldd, int N, int H, int W,
                                                        int Cp, const float* __restrict__ params,
                                                        const float* __restrict__ sums, float inv_count,
                                                        int per_frame, int normed, float slope, int pool,
                                                        float* __restrict__ dx, int lddx, int Hp, int Wp,
                                                        long long total, const FastDiv fcv, const FastDiv fw,
                                                        const FastDiv fh) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= (unsigned)total) return;
    {
        unsigned cq, wq, hq;
        const unsigned pix = fd_divmod(i, fcv, cq);
        const unsigned t = fd_divmod(pix, fw, wq);
        const unsigned n = fd_divmod(t, fh, hq);
        const int c = (int)cq * 4, w = (int)wq, h = (int)hq;
        const long long g = per_frame ? n : 0;
        float4 mean = f4zero(), invstd = f4zero(), sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = f4zero();
        if (params) {
            const float* pr = params + g * 4 * Cp;
            mean = ldg4(pr + c); invstd = ldg4(pr + Cp + c); sc = ldg4(pr + 2 * Cp + c); sh = ldg4(pr + 3 * Cp + c);
        }
        const float4 v = ldg4(x + (long long)pix * ldx + c);
        float4 d;
        if (pool) {
            const int hp = h >> 1, wp = w >> 1;
            d = (hp < Hp && wp < Wp) ? ldg4(dout + (((long long)n * Hp + hp) * Wp + wp) * ldd + c) * 0.25f : f4zero();
        } else {
            d = ldg4(dout + (long long)pix * ldd + c);
        }
        if (slope >= 0.f) {
            float4 z = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z),
                                   fmaf(v.w, sc.w, sh.w));
            d.x = z.x > 0.f ? d.x : d.x * slope; d.y = z.y > 0.f ? d.y : d.y * slope;
            d.z = z.z > 0.f ? d.z : d.z * slope; d.w = z.w > 0.f ? d.w : d.w * slope;
        }
        float4 r;
        if (normed) {
            const float* sp = sums + g * 2 * Cp;
            const float4 s1 = ldg4(sp + c) * inv_count, s2 = ldg4(sp + Cp + c) * inv_count;
            r.x = sc.x * (d.x - s1.x - (v.x - mean.x) * invstd.x * s2.x);
            r.y = sc.y * (d.y - s1.y - (v.y - mean.y) * invstd.y * s2.y);
            r.z = sc.z * (d.z - s1.z - (v.z - mean.z) * invstd.z * s2.z);
            r.w = sc.w * (d.w - s1.w - (v.w - mean.w) * invstd.w * s2.w);
        } else {
            r = d * sc;
        }
        st4(dx + (long long)pix * lddx + c, r);
    }
}

MK_EXPORT int mk_norm_bwd_apply(const float* x, int ldx, const float* dout, int ldd, int N, int H, int W, int Cp,
                                const float* params, const float* sums, double count, int per_frame, int normed,
                                float slope, int pool, float* dx, int lddx, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ldx % 4 == 0 && ldd % 4 == 0 && lddx % 4 == 0, "mk_norm_bwd_apply: channels must be x4");
    MK_REQUIRE(!normed || (params && sums), "mk_norm_bwd_apply: normed needs params and sums");
    const long long total = (long long)N * H * W * (Cp / 4);
    if (total == 0) return 0;
    MK_REQUIRE(total < (1LL << 31), "mk_norm_bwd_apply: more than 2^31 work items");
    k_norm_bwd_apply<<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, dout, ldd, N, H, W, Cp, params, sums, (float)(1.0 / count), per_frame, normed, slope, pool, dx, lddx,
        pool ? H >> 1 : H, pool ? W >> 1 : W, total, make_fastdiv(Cp / 4), make_fastdiv(W), make_fastdiv(H));
    return mk_check_launch("mk_norm_bwd_apply");
}
