// Convolution, fp32 exact-parity path: NHWC implicit GEMM on the FFMA pipe.
//   rows   M = output pixels (N*Ho*Wo, optionally ordered so a thread owns all four pixels of a 2x2 pool window)
//   cols     = output channels
//   depth  K = R*S*Cin_p, flattened (tap-major, channel-minor) so a float4 never straddles taps
// Operands are staged through shared memory (double buffered, one __syncthreads per K tile); the im2col gather,
// the nearest-x2 upsample (util.py:84) and the zero padding are address arithmetic in the loader, the bias /
// folded-BN affine, residual add, activation and 2x2 pool are the epilogue.  The tensor-core (tcgen05) path for
// the same contract lives in conv_tc.cu; this file is the bit-faithful fp32 reference implementation on device
// that the parity tests pin to the oracle at 1e-5.
#include "common.cuh"
#include <cuda_bf16.h>

struct ConvP {
    const float* x; int N, Hin, Win, Hl, Wl, Cin_p, ldx, ups;
    const float* w; int R, S, pad, Ktot;
    int Ho, Wo, Hp, Wp;  // Hp/Wp = pooled dims (== Ho/Wo when pool == 0)
    const float* scale; const float* shift; const float* resid; int ldr, act; float slope;
    float* y; int Cout_p, ldy, pool;
    long long M;  // GEMM rows
};

constexpr int BK = 16;
constexpr int APAD = 4;

template <int BM, int BN, int TM, bool POOL>
__global__ void __launch_bounds__(256) k_conv_ffma(const ConvP p) {
    constexpr int TN = 4;
    constexpr int NTX = BN / TN;
    constexpr int NTY = BM / TM;
    static_assert(NTX * NTY == 256, "tile/thread mismatch");
    constexpr int A_PER = BM * (BK / 4) / 256;           // float4 per thread per A tile
    constexpr int B_VECS = BK * BN / 4;                  // float4 per B tile
    constexpr int B_PER = (B_VECS + 255) / 256;

    __shared__ __align__(16) float As[2][BM][BK + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- per-thread A-row bookkeeping (rows this thread LOADS, not the rows it computes)
    const float* a_base[A_PER];
    int a_hi0[A_PER], a_wi0[A_PER];
    bool a_ok[A_PER];
    const int a_kv = tid & 3;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int row = (tid >> 2) + i * 64;
        long long m;
        int ho, wo;
        long long n;
        if (POOL) {
            int q = row / (BM / 4), pl = row % (BM / 4);
            long long pp = (long long)blockIdx.x * (BM / 4) + pl;
            m = pp;  // validity in pooled-pixel units
            a_ok[i] = pp < p.M / 4;
            int wp = (int)(pp % p.Wp);
            long long t = pp / p.Wp;
            int hp = (int)(t % p.Hp);
            n = t / p.Hp;
            ho = 2 * hp + (q >> 1);
            wo = 2 * wp + (q & 1);
        } else {
            m = m0 + row;
            a_ok[i] = m < p.M;
            wo = (int)(m % p.Wo);
            long long t = m / p.Wo;
            ho = (int)(t % p.Ho);
            n = t / p.Ho;
        }
        a_base[i] = p.x + n * (long long)p.Hin * p.Win * p.ldx;
        a_hi0[i] = ho - p.pad;
        a_wi0[i] = wo - p.pad;
    }

    float4 a_reg[A_PER];
    float4 b_reg[B_PER];

    auto load_tiles = [&](int kt) {
        const int kg = kt * BK + a_kv * 4;
        int tap = kg / p.Cin_p;
        int ci = kg - tap * p.Cin_p;
        int r = tap / p.S;
        int s = tap - r * p.S;
        const bool kok = kg < p.Ktot;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int hi = a_hi0[i] + r, wi = a_wi0[i] + s;
            bool ok = kok && a_ok[i] && hi >= 0 && hi < p.Hl && wi >= 0 && wi < p.Wl;
            a_reg[i] = ok ? ldg4(a_base[i] + ((long long)(hi >> p.ups) * p.Win + (wi >> p.ups)) * p.ldx + ci)
                          : f4zero();
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int l = tid + i * 256;
            if (B_VECS >= 256 || l < B_VECS) {
                int kr = l / (BN / 4), nv = l % (BN / 4);
                int k = kt * BK + kr, n = n0 + nv * 4;
                b_reg[i] = (k < p.Ktot && n < p.Cout_p) ? ldg4(p.w + (long long)k * p.Cout_p + n) : f4zero();
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) st4(&As[buf][(tid >> 2) + i * 64][a_kv * 4], a_reg[i]);
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int l = tid + i * 256;
            if (B_VECS >= 256 || l < B_VECS) st4(&Bs[buf][l / (BN / 4)][(l % (BN / 4)) * 4], b_reg[i]);
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nkt = (p.Ktot + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = ld4(&As[buf][ty + i * NTY][kk]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float4 b = ld4(&Bs[buf][kk + j4][tx * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    float av = j4 == 0 ? a[i].x : j4 == 1 ? a[i].y : j4 == 2 ? a[i].z : a[i].w;
                    acc[i][0] = fmaf(av, b.x, acc[i][0]);
                    acc[i][1] = fmaf(av, b.y, acc[i][1]);
                    acc[i][2] = fmaf(av, b.z, acc[i][2]);
                    acc[i][3] = fmaf(av, b.w, acc[i][3]);
                }
            }
        }
        if (kt + 1 < nkt) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const int n = n0 + tx * 4;
    if (n >= p.Cout_p) return;
    float4 sc = p.scale ? ldg4(p.scale + n) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 sh = p.shift ? ldg4(p.shift + n) : f4zero();
    auto finish = [&](float* a4, long long pix) -> float4 {
        float4 v = make_float4(fmaf(a4[0], sc.x, sh.x), fmaf(a4[1], sc.y, sh.y), fmaf(a4[2], sc.z, sh.z),
                               fmaf(a4[3], sc.w, sh.w));
        if (p.resid) v = v + ldg4(p.resid + pix * p.ldr + n);
        if (p.act == 1) {
            v.x = v.x > 0.f ? v.x : v.x * p.slope; v.y = v.y > 0.f ? v.y : v.y * p.slope;
            v.z = v.z > 0.f ? v.z : v.z * p.slope; v.w = v.w > 0.f ? v.w : v.w * p.slope;
        } else if (p.act == 2) {
            v.x = 1.f / (1.f + expf(-v.x)); v.y = 1.f / (1.f + expf(-v.y));
            v.z = 1.f / (1.f + expf(-v.z)); v.w = 1.f / (1.f + expf(-v.w));
        }
        return v;
    };
    if (POOL) {
        // rows owned: ty + i*NTY  ->  q = row / (BM/4), pl = row % (BM/4)
        constexpr int PER = TM / 4;  // pooled pixels per thread
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            int pl = ty + j * NTY;
            long long pp = (long long)blockIdx.x * (BM / 4) + pl;
            if (pp >= p.M / 4) continue;
            int wp = (int)(pp % p.Wp);
            long long t = pp / p.Wp;
            int hp = (int)(t % p.Hp);
            long long nn = t / p.Hp;
            float4 sum = f4zero();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int i = q * PER + j;  // row = q*(BM/4) + pl = ty + i*NTY  with NTY*PER == BM/4
                long long pix = (nn * p.Ho + (2 * hp + (q >> 1))) * p.Wo + (2 * wp + (q & 1));
                sum = sum + finish(acc[i], pix);
            }
            if (p.pool == 1) sum = sum * 0.25f;
            st4(p.y + ((nn * p.Hp + hp) * p.Wp + wp) * p.ldy + n, sum);
        }
    } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            long long m = m0 + ty + i * NTY;
            if (m >= p.M) continue;
            st4(p.y + m * p.ldy + n, finish(acc[i], m));
        }
    }
}

template <int BM, int BN, int TM>
static void launch_conv(const ConvP& p, cudaStream_t st) {
    if (p.pool) {
        dim3 grid((unsigned)mk_cdiv(p.M / 4, BM / 4), (unsigned)mk_cdiv(p.Cout_p, BN));
        k_conv_ffma<BM, BN, TM, true><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid((unsigned)mk_cdiv(p.M, BM), (unsigned)mk_cdiv(p.Cout_p, BN));
        k_conv_ffma<BM, BN, TM, false><<<grid, 256, 0, st>>>(p);
    }
}

MK_EXPORT int mk_conv2d(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups, const float* wpack,
                        int R, int S, int pad, const float* scale, const float* shift, const float* resid, int ldr,
                        int act, float slope, float* y, int Cout_p, int ldy, int pool, void* stream) {
    MK_REQUIRE(Cin_p % 4 == 0 && Cout_p % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "mk_conv2d: channels must be x4");
    MK_REQUIRE(!resid || ldr % 4 == 0, "mk_conv2d: resid ld must be x4");
    ConvP p;
    p.x = x; p.N = N; p.Hin = Hin; p.Win = Win; p.ups = ups ? 1 : 0;
    p.Hl = Hin << p.ups; p.Wl = Win << p.ups; p.Cin_p = Cin_p; p.ldx = ldx;
    p.w = wpack; p.R = R; p.S = S; p.pad = pad; p.Ktot = R * S * Cin_p;
    p.Ho = p.Hl + 2 * pad - R + 1; p.Wo = p.Wl + 2 * pad - S + 1;
    MK_REQUIRE(p.Ho > 0 && p.Wo > 0, "mk_conv2d: empty output");
    p.scale = scale; p.shift = shift; p.resid = resid; p.ldr = ldr; p.act = act; p.slope = slope;
    p.y = y; p.Cout_p = Cout_p; p.ldy = ldy; p.pool = pool;
    if (pool) {
        p.Hp = p.Ho >> 1; p.Wp = p.Wo >> 1;
        p.M = (long long)N * p.Hp * p.Wp * 4;
    } else {
        p.Hp = p.Ho; p.Wp = p.Wo;
        p.M = (long long)N * p.Ho * p.Wo;
    }
    if (p.M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout_p <= 16) launch_conv<256, 16, 4>(p, st);
    else if (Cout_p <= 32) launch_conv<128, 32, 4>(p, st);
    else if (p.M <= 64LL * mk_num_sms()) launch_conv<64, 64, 4>(p, st);
    else launch_conv<128, 64, 8>(p, st);
    return mk_check_launch("mk_conv2d");
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[k][co] = sum_pixels A[pix][k] * dY[pix][co],  A = im2col(x) gathered on the fly.  Split over pixel ranges
// (grid.z) with fp32 atomics into the zero-filled packed gradient.
struct WgP {
    const float* x; int N, Hin, Win, Hl, Wl, Cin_p, ldx, ups;
    const float* dy; int Cout_p, ldy, R, S, pad, Ktot, Ho, Wo;
    float* dw; long long M, chunk;
};

template <int BKK, int BN, int TM>
__global__ void __launch_bounds__(256) k_conv_wgrad(const WgP p) {
    constexpr int BP = 16, TN = 4;
    constexpr int NTX = BN / TN, NTY = BKK / TM;
    static_assert(NTX * NTY == 256, "tile/thread mismatch");
    constexpr int A_PER = BP * BKK / 4 / 256;
    constexpr int B_VECS = BP * BN / 4;
    __shared__ __align__(16) float As[2][BP][BKK];
    __shared__ __align__(16) float Bs[2][BP][BN];
    const int tid = threadIdx.x, tx = tid % NTX, ty = tid / NTX;
    const int k0 = blockIdx.x * BKK, n0 = blockIdx.y * BN;
    const long long p_begin = (long long)blockIdx.z * p.chunk;
    long long p_end = p_begin + p.chunk;
    if (p_end > p.M) p_end = p.M;

    // A loads: vector index l = tid + i*256 -> pixel l / (BKK/4), kvec l % (BKK/4); k fixed per (thread,i)
    int a_r[A_PER], a_s[A_PER], a_ci[A_PER], a_p[A_PER], a_k[A_PER];
    bool a_kok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int l = tid + i * 256;
        a_p[i] = l / (BKK / 4);
        a_k[i] = (l % (BKK / 4)) * 4;
        int kg = k0 + a_k[i];
        a_kok[i] = kg < p.Ktot;
        int tap = kg / p.Cin_p;
        a_ci[i] = kg - tap * p.Cin_p;
        a_r[i] = tap / p.S;
        a_s[i] = tap - a_r[i] * p.S;
    }
    float4 a_reg[A_PER], b_reg;
    auto load_tiles = [&](long long pb) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            long long m = pb + a_p[i];
            bool ok = a_kok[i] && m < p_end;
            float4 v = f4zero();
            if (ok) {
                int wo = (int)(m % p.Wo);
                long long t = m / p.Wo;
                int ho = (int)(t % p.Ho);
                long long n = t / p.Ho;
                int hi = ho - p.pad + a_r[i], wi = wo - p.pad + a_s[i];
                if (hi >= 0 && hi < p.Hl && wi >= 0 && wi < p.Wl)
                    v = ldg4(p.x + ((n * p.Hin + (hi >> p.ups)) * p.Win + (wi >> p.ups)) * p.ldx + a_ci[i]);
            }
            a_reg[i] = v;
        }
        if (tid < B_VECS) {
            int pp = tid / (BN / 4), nv = tid % (BN / 4);
            long long m = pb + pp;
            int n = n0 + nv * 4;
            b_reg = (m < p_end && n < p.Cout_p) ? ldg4(p.dy + m * p.ldy + n) : f4zero();
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) st4(&As[buf][a_p[i]][a_k[i]], a_reg[i]);
        if (tid < B_VECS) st4(&Bs[buf][tid / (BN / 4)][(tid % (BN / 4)) * 4], b_reg);
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (p_begin < p_end) {
        load_tiles(p_begin);
        store_tiles(0);
        __syncthreads();
        int it = 0;
        for (long long pb = p_begin; pb < p_end; pb += BP, ++it) {
            const int buf = it & 1;
            const bool more = pb + BP < p_end;
            if (more) load_tiles(pb + BP);
#pragma unroll
            for (int pp = 0; pp < BP; ++pp) {
                float a[TM];
#pragma unroll
                for (int i4 = 0; i4 < TM / 4; ++i4) {
                    float4 t = ld4(&As[buf][pp][ty * TM + i4 * 4]);
                    a[i4 * 4 + 0] = t.x; a[i4 * 4 + 1] = t.y; a[i4 * 4 + 2] = t.z; a[i4 * 4 + 3] = t.w;
                }
                float4 b = ld4(&Bs[buf][pp][tx * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                    acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                    acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
                }
            }
            if (more) store_tiles(buf ^ 1);
            __syncthreads();
        }
    }
    const int n = n0 + tx * 4;
    if (n >= p.Cout_p) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int k = k0 + ty * TM + i;
        if (k >= p.Ktot) continue;
        float* dst = p.dw + (long long)k * p.Cout_p + n;
        if (gridDim.z == 1) {
            st4(dst, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        } else {
            atomicAdd(dst + 0, acc[i][0]); atomicAdd(dst + 1, acc[i][1]);
            atomicAdd(dst + 2, acc[i][2]); atomicAdd(dst + 3, acc[i][3]);
        }
    }
}

MK_EXPORT int mk_conv2d_wgrad(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups, const float* dy,
                              int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    MK_REQUIRE(Cin_p % 4 == 0 && Cout_p % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "mk_conv2d_wgrad: channels x4");
    WgP p;
    p.x = x; p.N = N; p.Hin = Hin; p.Win = Win; p.ups = ups ? 1 : 0; p.Hl = Hin << p.ups; p.Wl = Win << p.ups;
    p.Cin_p = Cin_p; p.ldx = ldx; p.dy = dy; p.Cout_p = Cout_p; p.ldy = ldy; p.R = R; p.S = S; p.pad = pad;
    p.Ktot = R * S * Cin_p; p.Ho = p.Hl + 2 * pad - R + 1; p.Wo = p.Wl + 2 * pad - S + 1;
    p.dw = dwpack; p.M = (long long)N * p.Ho * p.Wo;
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = Cout_p <= 16;
    const int BKK = narrow ? 256 : 128, BN = narrow ? 16 : 64;
    long long tiles = mk_cdiv(p.Ktot, BKK) * mk_cdiv(Cout_p, BN);
    long long splits = mk_cdiv(2LL * mk_num_sms(), tiles);
    long long max_splits = mk_cdiv(p.M, 64);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    p.chunk = mk_cdiv(mk_cdiv(p.M, splits), 16) * 16;
    splits = mk_cdiv(p.M, p.chunk);
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dwpack, 0, sizeof(float) * (size_t)p.Ktot * Cout_p, st);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)mk_cdiv(p.Ktot, BKK), (unsigned)mk_cdiv(Cout_p, BN), (unsigned)splits);
    if (narrow) k_conv_wgrad<256, 16, 4><<<grid, 256, 0, st>>>(p);
    else k_conv_wgrad<128, 64, 8><<<grid, 256, 0, st>>>(p);
    return mk_check_launch("mk_conv2d_wgrad");
}

// ------------------------------------------------------------------------------------------------ weight (un)packing
// Parameter layout (Co, Ci/groups, 1, R, S) <-> GEMM layout [tap][Kin_p][Kout_p]; see include/monkey_b200.h.
// cross operand of one weight (row = tap * Kout + ko of the K-major pack, ki = input channel, v = hi + remainder)
__device__ __forceinline__ void pack_cross(__nv_bfloat16* cross, long long row, int ki, int Kin, float v, float hi) {
    const int Kin8 = (Kin + 7) & ~7;
    __nv_bfloat16* c = cross + (row * Kin8 + (ki & ~7)) * 2;
    const int j = ki & 7;
    c[j] = __float2bfloat16_rn(v);
    c[8 + j] = __float2bfloat16_rn(v - hi);
    if (ki + 4 >= Kin8) return;
    if (ki + 4 >= Kin) {   // Kin = 8 q + 4: the last group holds 4 real channels, zero the other 4
        c[j + 4] = __float2bfloat16_rn(0.f);
        c[12 + j] = __float2bfloat16_rn(0.f);
    }
}

__global__ void k_pack_weight(const float* __restrict__ w, int Co, int Cig, int R, int S, int groups,
                              const int* __restrict__ cin_map, int Cin_p, int Cout_p, int mode,
                              float* __restrict__ wp, long long total, const float* __restrict__ bias,
                              float* __restrict__ bias_p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bias_p && i < Cout_p) bias_p[i] = (bias && i < Co) ? bias[i] : 0.f;
    if (i >= total) return;
    // (a shared-memory-tiled variant with coalesced reads AND writes measured no faster in the step - 46.86 vs 46.74 ms,
    //  the packs are latency-, not bandwidth-bound - and was dropped)
    // modes 2/3 = modes 0/1 in the tensor-core layout [tap][Kout][Kin] (K-major rows), values rounded to TF32;
    // bit 3 (mode | 8): reference-precision pack - behind the hi half (offset `total`) follows the CROSS operand of the
    // BF16 correction MMA (conv_halo.cu): [tap][Kout][Kin rounded up to 8] 4-byte slots, every group of 8 input
    // channels stored as 16 bf16 = [bf16(v) x8 | bf16(v - hi) x8]
    const int x3 = (mode >> 3) & 1;
    mode &= 7;
    const int tc = mode >> 1;
    mode &= 1;
    const int Kin = mode == 0 ? Cin_p : Cout_p, Kout = mode == 0 ? Cout_p : Cin_p;
    int ko, ki, tap;
    if (tc) {
        ki = (int)(i % Kin);
        long long t = i / Kin;
        ko = (int)(t % Kout);
        tap = (int)(t / Kout);
    } else {
        ko = (int)(i % Kout);
        long long t = i / Kout;
        ki = (int)(t % Kin);
        tap = (int)(t / Kin);
    }
    int r = tap / S, s = tap - r * S;
    int ci_p = mode == 0 ? ki : ko, co = mode == 0 ? ko : ki;
    if (mode == 1) { r = R - 1 - r; s = S - 1 - s; }
    float v = 0.f;
    int ci = cin_map ? cin_map[ci_p] : ci_p;
    if (co < Co && ci >= 0) {
        int cog = Co / groups;
        int g = co / cog;
        int cil = ci - g * Cig;
        if (cil >= 0 && cil < Cig) v = w[(((long long)co * Cig + cil) * R + r) * S + s];
    }
    if (tc) {
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float hi = __uint_as_float(u);
        if (x3) pack_cross(reinterpret_cast<__nv_bfloat16*>(wp + total), (long long)tap * Kout + ko, ki, Kin, v, hi);
        v = hi;
    }
    wp[i] = v;
}

// mode 4: sub-pixel pack of a 3x3 kernel applied after a nearest x2 upsample (util.py:84-85).  For output parity
// (py,px) the conv is a 2x2 conv of the low-resolution input whose taps are sums of the 3x3 taps:
//   py = 0: rows {0} | {1,2}      py = 1: rows {0,1} | {2}        (same for columns)
// Layout [parity = py*2+px][tap = r2*2+s2][Cout_p][Cin_p], rounded to TF32 after the sum.
__global__ void k_pack_weight_ups(const float* __restrict__ w, int Co, int Ci, const int* __restrict__ cin_map,
                                  int Cin_p, int Cout_p, float* __restrict__ wp, long long total, int x3) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci_p = (int)(i % Cin_p);
    long long t = i / Cin_p;
    const int co = (int)(t % Cout_p);
    t /= Cout_p;
    const int tap = (int)(t & 3), par = (int)(t >> 2);
    const int py = par >> 1, px = par & 1, r2 = tap >> 1, s2 = tap & 1;
    const int ci = cin_map ? cin_map[ci_p] : ci_p;
    float v = 0.f;
    if (co < Co && ci >= 0 && ci < Ci) {
        const int r_lo = py == 0 ? (r2 == 0 ? 0 : 1) : (r2 == 0 ? 0 : 2), r_hi = py == 0 ? (r2 == 0 ? 0 : 2) : (r2 == 0 ? 1 : 2);
        const int s_lo = px == 0 ? (s2 == 0 ? 0 : 1) : (s2 == 0 ? 0 : 2), s_hi = px == 0 ? (s2 == 0 ? 0 : 2) : (s2 == 0 ? 1 : 2);
        const float* wk = w + ((long long)co * Ci + ci) * 9;
        for (int r = r_lo; r <= r_hi; ++r)
            for (int sx = s_lo; sx <= s_hi; ++sx) v += wk[r * 3 + sx];
    }
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    const float hi = __uint_as_float(u);
    wp[i] = hi;
    if (x3) pack_cross(reinterpret_cast<__nv_bfloat16*>(wp + total), t * Cout_p + co, ci_p, Cin_p, v, hi);
}

MK_EXPORT int mk_pack_weight(const float* w, int Co, int Cig, int R, int S, int groups, const int* cin_map,
                             int Cin_p, int Cout_p, int mode, float* wpack, const float* bias, float* bias_p,
                             void* stream) {
    long long total = (long long)R * S * Cin_p * Cout_p;
    if (total == 0) return 0;
    MK_REQUIRE(total >= Cout_p, "mk_pack_weight: degenerate shape");
    if ((mode & 7) == 4) {
        MK_REQUIRE(R == 3 && S == 3 && groups == 1, "mk_pack_weight mode 4: 3x3 ungrouped kernels only");
        total = 16LL * Cin_p * Cout_p;
        k_pack_weight_ups<<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, Co, Cig, cin_map, Cin_p,
                                                                                           Cout_p, wpack, total, mode >> 3);
        if (bias_p)  // zero-padded bias copy: the regular kernel with an empty weight range (total = 0)
            k_pack_weight<<<(unsigned)mk_cdiv(Cout_p, 256), 256, 0, (cudaStream_t)stream>>>(
                w, Co, Cig, R, S, groups, cin_map, Cin_p, Cout_p, 0, wpack, 0, bias, bias_p);
        return mk_check_launch("mk_pack_weight(ups)");
    }
    k_pack_weight<<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        w, Co, Cig, R, S, groups, cin_map, Cin_p, Cout_p, mode, wpack, total, bias, bias_p);
    return mk_check_launch("mk_pack_weight");
}

template <bool ACC>
__global__ void k_unpack_wgrad(const float* __restrict__ dwp, int Co, int Cig, int R, int S, int groups,
                               const int* __restrict__ cin_inv, int Cin_p, int Cout_p, float* __restrict__ dw,
                               long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int s = (int)(i % S);
    long long t = i / S;
    int r = (int)(t % R);
    t /= R;
    int cil = (int)(t % Cig);
    int co = (int)(t / Cig);
    int cog = Co / groups;
    int ci = (co / cog) * Cig + cil;        // logical input channel
    int ci_p = cin_inv ? cin_inv[ci] : ci;  // physical position
    const float v = dwp[((long long)(r * S + s) * Cin_p + ci_p) * Cout_p + co];
    if (ACC) dw[i] += v; else dw[i] = v;
}

// cin_map here is the INVERSE map (logical -> physical), length = logical Cin.
MK_EXPORT int mk_unpack_wgrad(const float* dwpack, int Co, int Cig, int R, int S, int groups, const int* cin_inv,
                              int Cin_p, int Cout_p, float* dw, void* stream) {
    long long total = (long long)Co * Cig * R * S;
    if (total == 0) return 0;
    k_unpack_wgrad<false><<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        dwpack, Co, Cig, R, S, groups, cin_inv, Cin_p, Cout_p, dw, total);
    return mk_check_launch("mk_unpack_wgrad");
}

// Same, ACCUMULATING into `grad` (the parameter's .grad tensor, parameter layout): the weight gradient goes from the packed
// GEMM result straight into the optimiser's flat gradient buffer - no temporary and no separate add kernel per parameter.
MK_EXPORT int mk_unpack_wgrad_acc(const float* dwpack, int Co, int Cig, int R, int S, int groups, const int* cin_inv,
                                  int Cin_p, int Cout_p, float* grad, void* stream) {
    long long total = (long long)Co * Cig * R * S;
    if (total == 0) return 0;
    k_unpack_wgrad<true><<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        dwpack, Co, Cig, R, S, groups, cin_inv, Cin_p, Cout_p, grad, total);
    return mk_check_launch("mk_unpack_wgrad_acc");
}

// Adjoint of the sub-pixel weight pack (mode 4): folds the gradient of the 16 (parity, 2x2 tap) kernels,
// dwpack_ups [16][Cin_p][Cout_p] from mk_conv2d_wgrad_halo_ups, back onto the 3x3 taps of the parameter.  Tap row r of
// the 3x3 kernel belongs to sub-row r2 = (r >= 1) for py = 0 ({0} | {1,2}) and r2 = (r >= 2) for py = 1 ({0,1} | {2}).
template <bool ACC>
__global__ void k_unpack_wgrad_ups(const float* __restrict__ dwp, int Co, int Ci, const int* __restrict__ cin_inv,
                                   int Cin_p, int Cout_p, float* __restrict__ dw, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = (int)(i % 3);
    long long t = i / 3;
    const int r = (int)(t % 3);
    t /= 3;
    const int ci = (int)(t % Ci), co = (int)(t / Ci);
    const int ci_p = cin_inv ? cin_inv[ci] : ci;
    float v = 0.f;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        const int r2 = py == 0 ? (r >= 1) : (r >= 2);
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const int s2 = px == 0 ? (s >= 1) : (s >= 2);
            const int slot = (py * 2 + px) * 4 + r2 * 2 + s2;
            v += dwp[((long long)slot * Cin_p + ci_p) * Cout_p + co];
        }
    }
    if (ACC) dw[i] += v; else dw[i] = v;
}

MK_EXPORT int mk_unpack_wgrad_ups(const float* dwpack_ups, int Co, int Ci, const int* cin_inv, int Cin_p, int Cout_p,
                                  float* dw, int accumulate, void* stream) {
    long long total = (long long)Co * Ci * 9;
    if (total == 0) return 0;
    if (accumulate)
        k_unpack_wgrad_ups<true><<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
            dwpack_ups, Co, Ci, cin_inv, Cin_p, Cout_p, dw, total);
    else
        k_unpack_wgrad_ups<false><<<(unsigned)mk_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
            dwpack_ups, Co, Ci, cin_inv, Cin_p, Cout_p, dw, total);
    return mk_check_launch("mk_unpack_wgrad_ups");
}
