// Halo-window tensor-core weight gradient for sm_100a:   dW[r,s][ci][co] = sum_pixels dY[pix][co] * X[pix + (r,s)][ci]
//
// k_wgrad_tc (wgrad_tc.cu) re-fetches the shifted X tile once per filter tap and runs the GEMM as D[co][ci] - with the
// small channel counts of the full-resolution layers (co = 48, 32, 12 ...) most of the 128 MMA rows are padding
// (48->48 @256x256 x 8: 233 us, 93 TFLOP/s; 128->32: 1.0 ms).  This kernel
//   * loads the X HALO of a pixel tile ONCE per 32-channel chunk (one 4-D TMA box {32 ch, 16 w, TR + 4 h, 1 n},
//     128B swizzle with 32-byte atoms = the MN-major UMMA layout) and reads every filter tap as a row-shifted window
//     of it: pixel m = 16*row + col of the tile pairs with halo row m + 16*r + s;
//   * puts the TAP ROWS on the M dimension: the A operand of one MMA is the halo window of column tap s, whose four
//     32-row M blocks are strided by ONE IMAGE ROW of the halo (descriptor leading-dimension byte offset = 16 pixels x
//     128 B): block j is the same 32 input channels seen through tap row r = j.  D[32 j + c][co] is therefore
//     dW[(j, s)][ci = c][co]: a 3x3 layer fills 96 of the 128 MMA rows whatever its channel counts are, N = co is
//     as narrow as the layer (N = 48 retires in 24 clk), and a (chunk, s) pair costs one MMA per 8 pixels;
//   * the dY tile is loaded once per pixel tile (box {32 ch, 16 w, TR h}); its junk columns (col >= 16 - (S-1), whose
//     shifted partners wrap into the next image row) are zeroed in shared memory by the helper warps.  In the
//     reference-precision mode (mk_conv2d_wgrad_halo_x3; scheme and error analysis in conv_halo.cu) they also round
//     both operands to TF32 in place and write the CROSS operands of the BF16 correction MMA: per pixel (128 B, the
//     same offset as in the fp32 tile) two 64-byte K rows of 32 channels - X: [bf16(x - hi) | bf16(x)], dY:
//     [bf16(dy) | bf16(dy - hi)] - an MN-major SWIZZLE_64B operand with the fp32 tile's LBO / SBO / K-step / tap-shift
//     byte offsets (both swizzles XOR with address bits 7-8 = pixel & 3).  Per 8 pixels: one kind::tf32 MMA (K = 8)
//     + one kind::f16 MMA (K = 16 rows = 8 pixels x {lo, top}) instead of the three of 3xTF32;
//   * accumulates in TMEM across ALL pixel tiles of the CTA's range ((chunks x S) accumulators of co columns), one
//     epilogue at the end: lane quarter q of the accumulator = tap row q, written (pixel splits: fp32 atomics) to the
//     packed gradient [tap][Cin_p][Cout_p].
//   * UPSAMPLED CONV (mk_conv2d_wgrad_halo_ups): the weight gradient of conv3x3(nearest_x2(x)) is taken on the LOW-resolution
//     grid as four sub-pixel passes (output parity py, px): dWsub[parity][2x2 tap] = sum dY[2h+py][2w+px] X[h + r2 - (1-py)]
//     [w + s2 - (1-px)] - the X halo is the low-resolution tensor itself (no upsampled copy), dY is read through a 5-D
//     TMA map (c, px, w, py, n*H + h), and mk_unpack_wgrad_ups folds the 16 sub-kernels back onto the 3x3 taps (the
//     adjoint of the sub-pixel weight pack).  16 tap-pixel products per low-res pixel instead of 36.
// Grid = (co tiles, ci-chunk groups, pixel splits).  Envelope: stride 1, R = S in {3, 4} (2: sub-pixel passes), image at
// least one tile.
#include "common.cuh"
#include "../../include/monkey_b200.h"
#include "tc_common.cuh"

namespace {
using namespace mk_tc;

constexpr int WH_THREADS = 256;   // warps 0-3 helpers (zero / split) then epilogue, 4 TMEM alloc, 5 TMA producer, 7 MMA issuer
constexpr int WH_SMEM_MAX = 227 * 1024;
constexpr int WH_MAXST = 4;

struct WHP {
    int N, Ho, Wo, Cout_p, Cin_p, R, S, pad_h, pad_w;
    int ups, py, px;             // sub-pixel pass: dY parity (py, px), X = the low-resolution input
    int TR, TWv, tilesW, tilesH, ntiles, tiles_per_split;
    int xa_half, dy_half;        // bytes of one ci-chunk halo / one 32-channel dY box (hi halves)
    int nci, nco, co_pad;        // ci chunks per CTA, dY boxes per CTA, UMMA N
    int x_region, dy_region;     // bytes of all X chunks (hi) / all dY boxes (hi) of one stage
    int stage_bytes, stages, x3, tmem_cols, total_chunks;
    float* dw;
};

// MN-major operand, SWIZZLE_128B with 32-byte atoms: [pixel rows][32 channels = 128 B]; the 4-pixel groups of the K
// direction are 512 B apart (SBO), the 32-channel blocks of the M / N direction `lbo` bytes apart.
__device__ __forceinline__ uint64_t desc_mn(const void* smem, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// layout-type field of the descriptor: SWIZZLE_128B_BASE32B (1) of the fp32 tiles -> SWIZZLE_64B (4) of the bf16 cross tiles
constexpr uint64_t CROSS_FLIP = ((uint64_t)1 ^ (uint64_t)4) << 61;

// all MMAs of one ci chunk for one pixel tile: S column taps x KS steps of 8 pixels, unrolled
template <int S, int KS, bool X3>
__device__ __forceinline__ void issue_chunk(uint32_t d0, int co_pad, uint64_t xa, uint64_t dy, uint64_t x_lo16,
                                            uint64_t dy_lo16, uint32_t idesc, uint32_t idesc_c, uint32_t acc_first) {
    // K-step outer, column tap inner: consecutive MMAs go to DIFFERENT accumulators.  Back-to-back MMAs into the same
    // TMEM tile are dependent (accumulate) and, at N <= 128, each costs a fixed ~50 clk on top of its N/2 clk of math
    // (80 clk measured for N = 48 whatever the kernel); interleaving the S independent accumulators lets them overlap.
#pragma unroll
    for (int k = 0; k < KS; ++k) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const uint32_t d = d0 + (uint32_t)(s * co_pad);
            const uint64_t a = xa + (uint64_t)(s * 8 + k * 64);   // + s pixel rows (128 B each), + 8 pixel rows per K step
            const uint64_t b = dy + (uint64_t)(k * 64);
            const uint32_t acc = k ? 1u : acc_first;
            if (X3) {
                umma_tf32(d, a, b, idesc, acc);                                               // x_hi * dy_hi
                umma_bf16(d, (a + x_lo16) ^ CROSS_FLIP, (b + dy_lo16) ^ CROSS_FLIP, idesc_c, 1u);   // x_lo * dy + x * dy_lo
            } else {
                umma_tf32(d, a, b, idesc, acc);
            }
        }
    }
}

template <int S, int TR, bool X3>
__global__ void __launch_bounds__(WH_THREADS, 1)
k_wgrad_halo(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, const WHP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* ready = full + WH_MAXST;
    uint64_t* empty = ready + WH_MAXST;
    uint64_t* tmem_full = empty + WH_MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * 128;
    const int ci_chunk0 = blockIdx.y * p.nci;
    const int nci = min(p.nci, p.total_chunks - ci_chunk0);
    const int n_this = min(128, p.Cout_p - co0);
    const int nco = (n_this + 31) >> 5;
    const int t0 = blockIdx.z * p.tiles_per_split;
    const int t1 = min(p.ntiles, t0 + p.tiles_per_split);
    const int tiles_per_img = p.tilesW * p.tilesH;
    constexpr int KS = TR * 2;

    if (warp == 5 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDy) : "memory");
    }
    if (warp == 7 && lane == 0) {
        for (int i = 0; i < WH_MAXST; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 4); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        // ===================================================================== TMA producer
        if (elect_one()) {
            int st = 0;
            uint32_t ph = 0;
            for (int tile = t0; tile < t1; ++tile) {
                const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
                const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
                const int w0 = tw * p.TWv, h0 = th * TR;
                mbar_wait(&empty[st], ph ^ 1);
                uint8_t* sb = smem + st * p.stage_bytes;
                mbar_expect_tx(&full[st], nci * p.xa_half + nco * p.dy_half);
                for (int c = 0; c < nci; ++c)
                    tma_load_4d(sb + c * p.xa_half, &tmX, &full[st], (ci_chunk0 + c) * 32, w0 - p.pad_w, h0 - p.pad_h, n);
                uint8_t* dyb = sb + (p.x_region << (X3 ? 1 : 0));
                for (int j = 0; j < nco; ++j)
                    if (p.ups) tma_load_5d(dyb + j * p.dy_half, &tmDy, &full[st], co0 + j * 32, p.px, w0, p.py, n * p.Ho + h0);
                    else tma_load_4d(dyb + j * p.dy_half, &tmDy, &full[st], co0 + j * 32, w0, h0, n);
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        }
    }
    if (warp < 4 && t1 > t0) {
        // ===================================================================== helpers = the four epilogue warps, idle
        // until the last tile: zero the junk columns of dY (their shifted X partners wrap into the next image row) and,
        // in 3xTF32 mode, split both operands hi / lo in place
        const int tid = threadIdx.x;   // 0..127
        int st = 0;
        uint32_t ph = 0;
        for (int tile = t0; tile < t1; ++tile) {
            mbar_wait(&full[st], ph);
            uint8_t* sb = smem + st * p.stage_bytes;
            uint8_t* dyb = sb + (p.x_region << (X3 ? 1 : 0));
            if (X3) {
                // one 16-byte chunk (4 channels) per lane and piece, consecutive lanes on consecutive chunks; lanes l, l ^ 1
                // hold the 8 channels of one 32-byte swizzle unit (the 32B-atom swizzle never splits it).  The even
                // lane assembles the first 64-byte K row of the pixel, the odd lane the second one.
                float4* hi = reinterpret_cast<float4*>(sb);
                uint4* cr = reinterpret_cast<uint4*>(sb + p.x_region);
                const int n4 = (nci * p.xa_half) >> 4;                        // multiple of 128
                const int odd = tid & 1;
                for (int i0 = tid; i0 < n4; i0 += 512) {
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + 128 * j < n4) v[j] = hi[i0 + 128 * j];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + 128 * j < n4) {
                            const int i = i0 + 128 * j;
                            float4 h;
                            uint2 lo, top;
                            split_cross(v[j], h, lo.x, lo.y, top.x, top.y);
                            uint2 send = odd ? lo : top, recv;
                            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                            hi[i] = h;
                            // X: K row 0 = bf16(x - hi), K row 1 = bf16(x); 16-byte slot = the 32-byte unit's index
                            cr[(i & ~7) + 4 * odd + ((i & 7) >> 1)] =
                                odd ? make_uint4(recv.x, recv.y, top.x, top.y) : make_uint4(lo.x, lo.y, recv.x, recv.y);
                        }
                }
                float4* dhi = reinterpret_cast<float4*>(dyb);
                uint4* dcr = reinterpret_cast<uint4*>(dyb + p.dy_region);
                const int m4 = (nco * p.dy_half) >> 4;
                for (int i0 = tid; i0 < m4; i0 += 512) {
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + 128 * j < m4) v[j] = dhi[i0 + 128 * j];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + 128 * j < m4) {
                            const int i = i0 + 128 * j;
                            const int pix = (i >> 3) & (TR * 16 - 1);       // 8 float4 per 128-byte pixel row
                            if ((pix & 15) >= p.TWv) v[j] = f4zero();
                            float4 h;
                            uint2 lo, top;
                            split_cross(v[j], h, lo.x, lo.y, top.x, top.y);
                            uint2 send = odd ? top : lo, recv;
                            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                            dhi[i] = h;
                            // dY: K row 0 = bf16(dy), K row 1 = bf16(dy - hi) (pairs with X's rows)
                            dcr[(i & ~7) + 4 * odd + ((i & 7) >> 1)] =
                                odd ? make_uint4(recv.x, recv.y, lo.x, lo.y) : make_uint4(top.x, top.y, recv.x, recv.y);
                        }
                }
            } else {
                // junk pixel rows only: (16 - TWv) of every 16 rows, 8 float4 each
                const int junk = 16 - p.TWv;
                const int per_box = TR * junk * 8;
                for (int i = tid; i < nco * per_box; i += 128) {
                    const int box = i / per_box, rem = i - box * per_box;
                    const int prow = rem >> 3, q4 = rem & 7;
                    const int row = prow / junk, col = p.TWv + (prow - row * junk);
                    reinterpret_cast<float4*>(dyb + box * p.dy_half + (row * 16 + col) * 128)[q4] = f4zero();
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[st]);
            if (++st == p.stages) { st = 0; ph ^= 1; }
        }
    }
    if (warp == 7) {
        // ===================================================================== MMA issuer
        if (elect_one() && t1 > t0) {
            // M = 128 = 4 tap rows x 32 channels (a_major = b_major = MN), N = co_pad
            const uint32_t idesc = umma_idesc_tf32(128, p.co_pad) | (1u << 15) | (1u << 16);
            const uint32_t idesc_c = umma_idesc_bf16(128, p.co_pad) | (1u << 15) | (1u << 16);
            const uint64_t x_lo16 = (uint64_t)(p.x_region >> 4), dy_lo16 = (uint64_t)(p.dy_region >> 4);
            const uint64_t xa16 = (uint64_t)(p.xa_half >> 4);
            int st = 0;
            uint32_t ph = 0;
#pragma unroll 1
            for (int tile = t0; tile < t1; ++tile) {
                mbar_wait(&ready[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint8_t* sb = smem + st * p.stage_bytes;
                uint64_t xa = desc_mn(sb, 16 * 128);                                   // M blocks: one image row apart
                const uint64_t dy = desc_mn(sb + (p.x_region << (X3 ? 1 : 0)), (uint32_t)p.dy_half);  // N blocks: next box
                const uint32_t acc_first = tile > t0 ? 1u : 0u;
                uint32_t d = tmem_base;
#pragma unroll 1
                for (int c = 0; c < nci; ++c, xa += xa16, d += (uint32_t)(S * p.co_pad))
                    issue_chunk<S, KS, X3>(d, p.co_pad, xa, dy, x_lo16, dy_lo16, idesc, idesc_c, acc_first);
                umma_commit(&empty[st]);
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
            umma_commit(tmem_full);
        }
    } else if (warp < 4 && t1 > t0) {
        // ===================================================================== epilogue: lane quarter = tap row
        const int r = warp;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool atomic = gridDim.z > 1;
        for (int c = 0; c < nci; ++c) {
            const int ci = (ci_chunk0 + c) * 32 + lane;
            const bool valid = r < p.R && ci < p.Cin_p;
            for (int s = 0; s < S; ++s) {
                const uint32_t tacc = tmem_base + ((uint32_t)(r * 32) << 16) + (uint32_t)((c * S + s) * p.co_pad);
                float* dst = p.dw + ((long long)(r * S + s) * p.Cin_p + ci) * p.Cout_p + co0;
                for (int col = 0; col < n_this; col += 16) {
                    float v[16];
                    tmem_ld16(tacc + (uint32_t)col, v);
                    if (!valid) continue;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (col + j >= n_this) break;
                        const float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (atomic) atomicAdd(reinterpret_cast<float4*>(dst + col + j), o);
                        else st4(dst + col + j, o);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

}  // namespace

static thread_local int t_whx3 = 0;
static thread_local int* t_whplan = nullptr;

static thread_local int t_whups = -1;   // >= 0: sub-pixel pass (parity py * 2 + px) of the upsampled conv's weight gradient

// Returns 0 on success, -2 outside the envelope (callers use mk_conv2d_wgrad_tc).
static int wgrad_halo_impl(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                           int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    const int ups_par = t_whups;
    const int Ho = ups_par >= 0 ? Hin : Hin + 2 * pad - R + 1, Wo = ups_par >= 0 ? Win : Win + 2 * pad - S + 1;
    if (Cin_p % 4 || Cout_p % 4 || ldx % 4 || ldy % 4 || R != S || (R != 3 && R != 4 && !(R == 2 && ups_par >= 0)) ||
        (ups_par >= 0 && (R != 2 || Hin % 8)) || Ho < 1 || Wo < 1) {
        mk_set_error("mk_conv2d_wgrad_halo: outside the halo kernel's envelope");
        return -2;
    }
    WHP p;
    p.x3 = t_whx3;
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout_p = Cout_p; p.Cin_p = Cin_p; p.R = R; p.S = S; p.dw = dwpack;
    p.ups = ups_par >= 0 ? 1 : 0;
    p.py = p.ups ? (ups_par >> 1) : 0;
    p.px = p.ups ? (ups_par & 1) : 0;
    p.pad_h = p.ups ? 1 - p.py : pad;
    p.pad_w = p.ups ? 1 - p.px : pad;
    p.TWv = 16 - (S - 1);
    p.tilesW = (Wo + p.TWv - 1) / p.TWv;
    p.total_chunks = (Cin_p + 31) / 32;
    const int co_tiles = (Cout_p + 127) / 128;
    const int n_tile = Cout_p < 128 ? Cout_p : 128;
    p.co_pad = (n_tile + 15) & ~15;
    p.nco = (n_tile + 31) / 32;
    const int sms = mk_num_sms();
    // plan: TR in {8, 4} rows per tile and ci chunks per CTA so that two stages fit in shared memory
    int best_tr = 0, best_nci = 0, best_groups = 0;
    for (int tr = 8; tr >= 4; tr >>= 1) {
        int nci = 512 / (S * p.co_pad);
        if (nci > p.total_chunks) nci = p.total_chunks;
        for (; nci >= 1; --nci) {
            const int stage = ((nci * (tr + 4) * 16 * 128) + (p.nco * tr * 16 * 128)) << p.x3;
            if (2 * stage + 2048 <= WH_SMEM_MAX) break;
        }
        if (nci < 1) continue;
        const int groups = (p.total_chunks + nci - 1) / nci;   // every chunk group re-reads the dY tile
        if (!best_tr || groups < best_groups) { best_tr = tr; best_nci = nci; best_groups = groups; }
    }
    if (!best_tr) {
        mk_set_error("mk_conv2d_wgrad_halo: no shared-memory plan");
        return -2;
    }
    // balance the chunk groups (5 chunks, 3 fit -> 3 + 2)
    const int ci_groups = (p.total_chunks + best_nci - 1) / best_nci;
    p.nci = (p.total_chunks + ci_groups - 1) / ci_groups;
    p.TR = best_tr;
    p.tilesH = (Ho + p.TR - 1) / p.TR;
    p.ntiles = p.tilesW * p.tilesH * N;
    const double useful = (double)Ho * Wo / ((double)p.tilesH * p.TR * p.tilesW * p.TWv);
    if (useful < 0.7 || p.ntiles < 8) {
        mk_set_error("mk_conv2d_wgrad_halo: %d tiles, %.0f %% useful: left to mk_conv2d_wgrad_tc", p.ntiles, 100.0 * useful);
        return -2;
    }
    p.xa_half = (p.TR + 4) * 16 * 128;
    p.dy_half = p.TR * 16 * 128;
    p.x_region = p.nci * p.xa_half;
    p.dy_region = p.nco * p.dy_half;
    p.stage_bytes = (p.x_region + p.dy_region) << p.x3;
    p.stages = (WH_SMEM_MAX - 2048) / p.stage_bytes;
    if (p.stages > WH_MAXST) p.stages = WH_MAXST;
    // the M = 128 window reads 4 image-row blocks past the last K step and the N blocks of absent dY boxes: keep every
    // operand address inside the allocation (one spare stage-sized margin is already there when stages >= 2)
    const int cols = p.nci * S * p.co_pad;
    p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    long long groups = (long long)co_tiles * ci_groups;
    // pixel splits: one wave of CTAs (1 CTA per SM); the epilogue of a CTA (atomics of nci x S x co x 96..128 rows)
    // costs about as much as 6 pixel tiles
    long long splits = pick_splits(groups, p.ntiles, sms, 6.0, 65535);
    p.tiles_per_split = (int)mk_cdiv(p.ntiles, splits);
    splits = mk_cdiv(p.ntiles, p.tiles_per_split);
    if (p.stages > p.tiles_per_split) p.stages = p.tiles_per_split < 2 ? 2 : p.tiles_per_split;
    int smem_bytes = p.stages * p.stage_bytes + 2048;
    MK_REQUIRE(p.stages >= 2 && smem_bytes <= WH_SMEM_MAX, "mk_conv2d_wgrad_halo: ring plan failed (%d)", smem_bytes);
    if (t_whplan) {
        int* o = t_whplan;
        o[0] = co_tiles; o[1] = ci_groups; o[2] = (int)splits; o[3] = smem_bytes; o[4] = p.stages; o[5] = p.TR;
        o[6] = p.nci; o[7] = p.nco; o[8] = p.tmem_cols; o[9] = p.ntiles; o[10] = p.tiles_per_split; o[11] = p.co_pad;
        o[12] = p.stage_bytes; o[13] = p.x3; o[14] = p.TWv; o[15] = p.tilesH;
        return 0;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_wgrad_halo: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tmX, tmDy;
    cuuint32_t es[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_p, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)Win * ldx * 4, (cuuint64_t)Hin * Win * ldx * 4};
        cuuint32_t box[4] = {32, 16, (cuuint32_t)(p.TR + 4), 1};
        CUresult rc = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(rc == CUDA_SUCCESS, "mk_conv2d_wgrad_halo: x tensor map rejected (%d)", (int)rc);
    }
    if (p.ups) {
        // dY[n][2h + py][2w + px][c] as (c, px, w, py, n * H + h): full-resolution strides, low-resolution tile grid
        cuuint64_t dims[5] = {(cuuint64_t)Cout_p, 2, (cuuint64_t)Wo, 2, (cuuint64_t)N * Ho};
        cuuint64_t strides[4] = {(cuuint64_t)ldy * 4, (cuuint64_t)2 * ldy * 4, (cuuint64_t)2 * Wo * ldy * 4,
                                 (cuuint64_t)4 * Wo * ldy * 4};
        cuuint32_t box[5] = {32, 1, 16, 1, (cuuint32_t)p.TR};
        cuuint32_t es5[5] = {1, 1, 1, 1, 1};
        CUresult rc = encode(&tmDy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(dy), dims, strides, box, es5,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(rc == CUDA_SUCCESS, "mk_conv2d_wgrad_halo: sub-pixel dy tensor map rejected (%d)", (int)rc);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)Cout_p, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldy * 4, (cuuint64_t)Wo * ldy * 4, (cuuint64_t)Ho * Wo * ldy * 4};
        cuuint32_t box[4] = {32, 16, (cuuint32_t)p.TR, 1};
        CUresult rc = encode(&tmDy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(rc == CUDA_SUCCESS, "mk_conv2d_wgrad_halo: dy tensor map rejected (%d)", (int)rc);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dwpack, 0, sizeof(float) * (size_t)R * S * Cin_p * Cout_p, st);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_halo memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)co_tiles, (unsigned)ci_groups, (unsigned)splits);
    cudaError_t le = cudaSuccess;
#define WH_LAUNCH(SS, TT, XX)                                                                                          \
    do {                                                                                                               \
        static unsigned long long attr_done = 0;                                                                       \
        if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {                                           \
            le = cudaFuncSetAttribute(k_wgrad_halo<SS, TT, XX>, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM_MAX); \
            if (le == cudaSuccess) attr_done |= attr_bit;                                                              \
        }                                                                                                              \
        if (le == cudaSuccess) k_wgrad_halo<SS, TT, XX><<<grid, WH_THREADS, smem_bytes, st>>>(tmX, tmDy, p);           \
    } while (0)
    if (S == 3) {
        if (p.TR == 8) { if (p.x3) WH_LAUNCH(3, 8, true); else WH_LAUNCH(3, 8, false); }
        else { if (p.x3) WH_LAUNCH(3, 4, true); else WH_LAUNCH(3, 4, false); }
    } else if (S == 2) {
        if (p.TR == 8) { if (p.x3) WH_LAUNCH(2, 8, true); else WH_LAUNCH(2, 8, false); }
        else { if (p.x3) WH_LAUNCH(2, 4, true); else WH_LAUNCH(2, 4, false); }
    } else {
        if (p.TR == 8) { if (p.x3) WH_LAUNCH(4, 8, true); else WH_LAUNCH(4, 8, false); }
        else { if (p.x3) WH_LAUNCH(4, 4, true); else WH_LAUNCH(4, 4, false); }
    }
#undef WH_LAUNCH
    if (le != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_halo: smem attribute: %s", cudaGetErrorString(le)); return (int)le; }
    return mk_check_launch("mk_conv2d_wgrad_halo");
}

MK_EXPORT int mk_conv2d_wgrad_halo(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                   int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    t_whups = -1;
    return wgrad_halo_impl(x, N, Hin, Win, Cin_p, ldx, dy, Cout_p, ldy, R, S, pad, dwpack, stream);
}

// Weight gradient of conv3x3(nearest_x2(x)), pad 1, on the low-resolution grid: x [N][Hin][Win][ldx] (NOT upsampled),
// dy [N][2 Hin][2 Win][ldy]; dwpack_ups [16 = parity x 2x2 tap][Cin_p][Cout_p] (the gradient of the mode-4 pack; fold it
// onto the 3x3 taps with mk_unpack_wgrad_ups).  -2 (nothing launched) outside the envelope.
MK_EXPORT int mk_conv2d_wgrad_halo_ups(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                       int Cout_p, int ldy, float* dwpack_ups, void* stream) {
    int rc = 0;
    for (int par = 0; par < 4 && rc == 0; ++par) {
        t_whups = par;
        rc = wgrad_halo_impl(x, N, Hin, Win, Cin_p, ldx, dy, Cout_p, ldy, 2, 2, 0,
                             dwpack_ups + (size_t)par * 4 * Cin_p * Cout_p, stream);
        if (rc == -2 && par > 0) {
            mk_set_error("mk_conv2d_wgrad_halo_ups: parity %d refused after parity 0 was launched", par);
            rc = -1;
        }
    }
    t_whups = -1;
    return rc;
}

MK_EXPORT int mk_conv2d_wgrad_halo_ups_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                          int Cout_p, int ldy, float* dwpack_ups, void* stream) {
    t_whx3 = 1;
    const int rc = mk_conv2d_wgrad_halo_ups(x, N, Hin, Win, Cin_p, ldx, dy, Cout_p, ldy, dwpack_ups, stream);
    t_whx3 = 0;
    return rc;
}

MK_EXPORT int mk_conv2d_wgrad_halo_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                      int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    t_whx3 = 1;
    const int rc = mk_conv2d_wgrad_halo(x, N, Hin, Win, Cin_p, ldx, dy, Cout_p, ldy, R, S, pad, dwpack, stream);
    t_whx3 = 0;
    return rc;
}

// Dry run of the planner: out[16] = co tiles, ci groups, pixel splits, smem bytes, stages, TR, ci chunks per CTA, dY
// boxes, TMEM columns, tiles, tiles per split, UMMA N, stage bytes, x3, valid tile width, tile rows of the image.
MK_EXPORT int mk_conv2d_wgrad_halo_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int R, int S, int pad, int x3,
                                        int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_wgrad_halo_plan: out is NULL");
    t_whplan = out;
    t_whx3 = x3 ? 1 : 0;
    const int rc = mk_conv2d_wgrad_halo(nullptr, N, Hin, Win, Cin_p, Cin_p, nullptr, Cout_p, Cout_p, R, S, pad, nullptr,
                                        nullptr);
    t_whplan = nullptr;
    t_whx3 = 0;
    return rc;
}

// Dry run of the four sub-pixel passes of the upsampled conv's weight gradient: same out[16] as above.
MK_EXPORT int mk_conv2d_wgrad_halo_ups_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int x3, int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_wgrad_halo_ups_plan: out is NULL");
    t_whplan = out;
    t_whx3 = x3 ? 1 : 0;
    const int rc = mk_conv2d_wgrad_halo_ups(nullptr, N, Hin, Win, Cin_p, Cin_p, nullptr, Cout_p, Cout_p, nullptr, nullptr);
    t_whplan = nullptr;
    t_whx3 = 0;
    return rc;
}
