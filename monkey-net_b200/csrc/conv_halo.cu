// Halo-window tensor-core convolution for sm_100a: persistent CTAs, resident weights, TMA-store epilogue.
//
// k_conv_tc (conv_tc.cu) fetches a shifted 128-pixel A tile once PER FILTER TAP and the weights of that tap once per
// tile: ncu on 48->48 3x3 @256x256 showed 1.66 GB crossing L2->SM for a 100 MB input (lts throughput 59 %, tensor pipe
// 14 %) - the many-tile, small-channel layers that dominate the 256x256 configurations are L2->SMEM bound there (the
// L2 hands every SM ~42 B/clk; an M128 x N48 x K8 TF32 MMA retires in 24 clk).  This kernel removes both re-fetches:
//
//   * HALO WINDOWS.  A super-tile is RB row-blocks of 8 rows x TWv columns, TWv = 16 - (S-1).  ONE 4-D TMA box
//     {32 ch, 16 w, 8*RB + R h, 1 n} per 32-channel chunk lands its halo as 128-byte pixel rows, 16 pixels per image
//     row, 128B-swizzled (out-of-bounds = zero fill = conv padding).  GEMM row m = 16*row + col of row-block rb and tap
//     (r, s) is shared-memory row m + 16*(8*rb + r) + s of that buffer, i.e. every tap of every row-block is the SAME
//     buffer read through a UMMA descriptor whose start address is advanced by whole 128-byte rows.  The 128B swizzle
//     XORs the 16-byte-chunk bits with ABSOLUTE address bits 7-9 (validated on a B200, tools/halo_probe.py: all 9 / 16
//     shifted windows agree with the exact kernel at TF32 rounding with descriptor base offset 0), so a row-shifted
//     window of a 1024-byte-aligned buffer is a legal canonical K-major operand.  Columns >= TWv of each 16-pixel row
//     are junk GEMM rows the epilogue never stores (87.5 % of the MMA rows useful for 3x3).
//   * RESIDENT WEIGHTS.  CTAs are persistent (grid = #SMs x cout tiles, static tile striding).  When the packed weights
//     of the CTA's cout tile fit ([tap][chunk][b_rows x 128 B] <= ~150 KB) they are loaded ONCE per CTA; otherwise they
//     stream through a ring and the planner picks RB = 2 or 4 so one weight fetch feeds 2-4 accumulators.
//   * REFERENCE PRECISION (mk_conv2d_tc_halo_x3): TF32 main term + BF16 cross terms, 2 MMAs per K step.  With
//     v = hi + lo (hi = rna_tf32(v)), a*b = a_hi*b_hi + (a_lo*b + a*b_lo) + O(2^-22).  The main term is one kind::tf32
//     MMA (K = 8 channels).  The two cross terms are 2^-11 of the product, so 8 mantissa bits are enough for them
//     (error 2^-20, the level of fp32 accumulation itself): ONE kind::f16 BF16 MMA (K = 16) whose K dimension is the
//     concatenation  A' = [bf16(a_lo) x8 | bf16(a) x8],  B' = [bf16(b) x8 | bf16(b_lo) x8]  of the same 8 channels -
//     32 bytes per row like the TF32 operand, same swizzled geometry, same descriptors.  The weights' cross operand is
//     written by mk_pack_weight (mode | 16) behind the hi half; the halo's is produced in shared memory by four
//     splitter warps.  Emulated error vs fp64 on random data: 6.6e-7 rms (fp32 matmul 2.7e-7, 3xTF32 7.8e-8,
//     1xTF32 2.9e-4); cost 2 operand passes instead of the 3 of 3xTF32.
//   * DOUBLE-BUFFERED TMEM.  2 x RB accumulators: the epilogue of tile t overlaps the MMAs of tile t+1.
//   * TMA-STORE EPILOGUE.  tcgen05.ld (thread = pixel) -> bias / affine / residual / activation -> 128B-swizzled
//     staging rows in shared memory -> cp.async.bulk.tensor store of {32 ch, TWv, 8} boxes: full-line writes, image /
//     channel edges clipped by the TMA unit.  The residual tile is prefetched by its own TMA producer warp.
//
//   * COLUMN TAPS ON N (template CT, layers with S * Cout_p <= 256, Cout_p % 16 == 0).  Measured on B200, an SS-mode
//     UTCHMMA costs ~18 clk + 0.35 clk per operand ROW it reads from shared memory (M + N rows of 32 bytes: 80 clk for
//     M128 x N48, 108 clk for N128) - the narrow layers are bound by operand fetch, not by the tensor pipe, and the
//     A window (128 rows) is re-read for every tap.  CT reads it once per tap ROW: the B operand of (chunk, r) is the
//     S column taps' weights stacked on N (rows s * Cout_p + co: ONE 3-D TMA box {32 ci, Cout_p, S} of the
//     [tap][Cout_p][Cin_p] pack), the MMA computes E[m][s, co] = sum_ci X[m + 16 r][ci] W[r, s][co][ci] for the
//     UNSHIFTED column of every pixel, and the epilogue adds the S partial columns of neighbouring pixels:
//     y[m][co] = sum_s E[m + s][s, co] - pixel m + s is lane + s of the same warp (16-pixel rows, the junk columns
//     col >= TWv are exactly the lanes that would wrap), i.e. S - 1 shuffles per output value.  3x3, Cout 48:
//     3 x (128 + 144) operand rows per K step instead of 9 x (128 + 48).
//
// Warp roles (384 threads): 0-3 = epilogue (TMEM lane quarter = warp), 4 = TMEM allocator + residual producer,
// 5 = TMA producer (halo, resident weights), 11 = TMA producer of streamed weights, 6-7 and 9-10 = cross-operand
// splitters, 8 = MMA issuer.  The issuer has the highest warp id of
// its scheduler (the arbiter is highest-wid-first) and its loop is fully unrolled over the filter taps (kernel
// template <R, S, X3, RESIDENT>): with N = 48 an MMA retires in 24 clk, so the single issuing thread can afford only
// a handful of instructions per MMA - the first version spent ~45 (runtime tap decode, 64-bit descriptor math, role
// flags read from constant memory) and sat at 15 % tensor-pipe utilisation with every barrier idle.  Envelope: stride 1, no upsample, R, S <= 4, enough tiles to fill the
// machine (the few-tile deep layers keep k_conv_tc's split-K).  Same contract and epilogue as mk_conv2d_tc.
#include "common.cuh"
#include "../../include/monkey_b200.h"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace mk_tc;

constexpr int HK = 32;                 // fp32 channels per chunk = 128 bytes
constexpr int MAXA = 4, MAXB = 40;     // ring depth bounds (MAXB also bounds the resident slots: 9 taps x 4 chunks = 36)
constexpr int H_SMEM_MAX = 227 * 1024;
constexpr int H_THREADS = 384;
constexpr int H_FIXED = 3072;   // barriers (<= 848 B) + the CTA's scale / shift vectors (1 KB) + 1023 B alignment slack

struct HP {
    int N, Ho, Wo, Cout_p, Cin_p, R, S, pad_h, pad_w;
    int ups, py, px;       // sub-pixel pass of conv3x3(nearest_x2(x)): 2x2 conv on the low-res grid, output parity (py, px)
    int TWv, RB, tilesW, tilesH, ntiles;
    int halo_rows, a_half, a_stage, a_stages;
    int b_rows, b_half, b_tx, b_slot, b_slots, resident;   // b_tx = bytes one weight TMA box delivers
    int nchunks, ngroups, npad, acc_cols, tmem_cols;   // acc_cols = TMEM columns per accumulator
    int x3, ct;            // ct: column taps stacked on N (b_rows = round16(S * Cout_p), one weight slot per tap ROW)
    int stg_bytes, nstg;   // nstg = 1..3 staging buffers (and as many residual buffers)
    const float* scale; const float* shift; int has_resid, act; float slope;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// One chunk's MMAs for one row-block with RESIDENT weights, fully unrolled over taps and K steps (NK compile-time):
// per MMA two 64-bit uniform adds + one UTCHMMA, no predicates, no loop-carried vector registers.
template <int R, int S, bool X3, int NK, int RBT, bool CT>
__device__ __forceinline__ void issue_taps_resident(uint32_t d, int acc_cols, uint64_t ad, uint64_t bd, uint64_t a_half16,
                                                    uint64_t b_half16, uint64_t b_slot16, uint32_t idesc,
                                                    uint32_t idesc2, uint32_t acc_first) {
    // tap -> K step -> row-block: consecutive MMAs alternate between the RBT independent accumulators (back-to-back
    // MMAs into the same TMEM tile are dependent and pay a fixed ~50 clk each on top of their N/2 clk of math)
#pragma unroll
    for (int tap = 0; tap < (CT ? R : R * S); ++tap) {
        // row-shifted window, 16-byte units (CT: one slot per tap ROW, the column taps ride on N)
        const uint64_t a0 = ad + (uint64_t)((CT ? tap * 16 : (tap / S) * 16 + (tap % S)) * 8);
        const uint64_t b = bd + (uint64_t)tap * b_slot16;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const uint32_t acc = (tap | k) ? 1u : acc_first;
#pragma unroll
            for (int rb = 0; rb < RBT; ++rb) {
                const uint64_t a = a0 + (uint64_t)(rb * 1024);               // next row-block: + 8 image rows = 16 KB
                const uint32_t dd = d + (uint32_t)(rb * acc_cols);
                if (X3) {
                    umma_tf32(dd, a + 2 * k, b + 2 * k, idesc, acc);                            // a_hi * b_hi
                    umma_bf16(dd, a + a_half16 + 2 * k, b + b_half16 + 2 * k, idesc2, 1u);      // a_lo * b + a * b_lo
                } else {
                    umma_tf32(dd, a + 2 * k, b + 2 * k, idesc, acc);
                }
            }
        }
    }
}

// Streaming weights: tap-major (one weight slot feeds every row-block before it is released).
template <int R, int S, bool X3, int NK, bool CT>
__device__ __forceinline__ void issue_taps_streaming(uint32_t dbase, int RB, int npad, uint64_t a_desc, uint64_t b_desc0,
                                                     uint64_t a_half16, uint64_t b_half16, uint64_t b_slot16,
                                                     uint32_t idesc, uint32_t idesc2, uint32_t acc_first,
                                                     uint64_t* b_full, uint64_t* b_empty, int& bs, uint32_t& bphase,
                                                     int b_slots) {
#pragma unroll
    for (int tap = 0; tap < (CT ? R : R * S); ++tap) {
        mbar_wait(&b_full[bs], bphase);
        const uint64_t b = b_desc0 + (uint64_t)bs * b_slot16;
        const uint64_t a0 = a_desc + (uint64_t)((CT ? tap * 16 : (tap / S) * 16 + (tap % S)) * 8);
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const uint32_t acc = (tap | k) ? 1u : acc_first;
            uint64_t a = a0;
            uint32_t d = dbase;
#pragma unroll 1
            for (int rb = 0; rb < RB; ++rb, a += 1024, d += (uint32_t)npad) {   // alternate the independent accumulators
                if (X3) {
                    umma_tf32(d, a + 2 * k, b + 2 * k, idesc, acc);
                    umma_bf16(d, a + a_half16 + 2 * k, b + b_half16 + 2 * k, idesc2, 1u);
                } else {
                    umma_tf32(d, a + 2 * k, b + 2 * k, idesc, acc);
                }
            }
        }
        umma_commit(&b_empty[bs]);
        if (++bs == b_slots) { bs = 0; bphase ^= 1; }
    }
}

template <int R, int S, bool X3, bool RES, bool CT>
__global__ void __launch_bounds__(H_THREADS, 1)
k_conv_halo(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmY,
            const __grid_constant__ CUtensorMap tmR, const HP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + p.a_stages * p.a_stage;
    uint8_t* stg = b_ring + p.b_slots * p.b_slot;             // nstg output staging buffers
    uint8_t* rbuf = stg + p.nstg * p.stg_bytes;               // nstg residual buffers (has_resid only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(rbuf + (p.has_resid ? p.nstg * p.stg_bytes : 0));
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + MAXA;
    uint64_t* a_split = a_empty + MAXA;
    uint64_t* b_full = a_split + MAXA;
    uint64_t* b_empty = b_full + MAXB;
    uint64_t* tmem_full = b_empty + MAXB;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* r_full = tmem_empty + 2;
    uint64_t* r_empty = r_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + 4);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);   // this CTA's 128 output channels: epilogue affine
    float* s_shift = s_scale + 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cout0 = blockIdx.y * 128;
    const int n_this = min(128, p.Cout_p - cout0);
    const int ngroups = (n_this + 31) >> 5;
    const int ntaps = CT ? p.R : p.R * p.S;   // weight slots per channel chunk
    const int tiles_per_img = p.tilesW * p.tilesH;

    if (warp == 5 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
        if (p.has_resid) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    }
    if (warp == 8 && lane == 0) {
        for (int i = 0; i < MAXA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_split[i], 4); }
        for (int i = 0; i < MAXB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        for (int i = 0; i < 4; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 128) {   // the epilogue reads these 8 x float4 per group from shared memory (broadcast), not from L1/L2
        const int co = cout0 + (int)threadIdx.x;
        s_scale[threadIdx.x] = (p.scale && co < p.Cout_p) ? p.scale[co] : 1.f;
        s_shift[threadIdx.x] = (p.shift && co < p.Cout_p) ? p.shift[co] : 0.f;
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
        // ===================================================================== TMA producer: halo (+ resident weights)
        // Streamed weights have their own producer (warp 11): one thread serving both rings in order issued the next
        // chunk's halo only after the last weight slot of the current chunk had found a free ring entry, i.e. a few
        // taps before the MMAs needed it - ncu showed the issuer waiting 30 % of its time for the split halo.
        if (elect_one()) {
            int ai = 0;
            for (int tile = blockIdx.x, lt = 0; tile < p.ntiles; tile += gridDim.x, ++lt) {
                const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
                const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
                const int w0 = tw * p.TWv, h0 = th * 8 * p.RB;
                for (int ch = 0; ch < p.nchunks; ++ch, ++ai) {
                    const int as = ai % p.a_stages;
                    mbar_wait(&a_empty[as], ((ai / p.a_stages) & 1) ^ 1);
                    mbar_expect_tx(&a_full[as], p.a_half);
                    tma_load_4d(a_ring + as * p.a_stage, &tmA, &a_full[as], ch * HK, w0 - p.pad_w, h0 - p.pad_h, n);
                    if (!p.resident || lt > 0) continue;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const int bs = ch * ntaps + tap;
                        uint8_t* b = b_ring + bs * p.b_slot;
                        mbar_expect_tx(&b_full[bs], p.b_tx << p.x3);
                        tma_load_3d(b, &tmB, &b_full[bs], ch * HK, cout0, CT ? tap * p.S : tap);
                        if (p.x3) tma_load_3d(b + p.b_half, &tmB2, &b_full[bs], ch * HK, cout0, CT ? tap * p.S : tap);
                    }
                }
            }
        }
    } else if (warp == 11) {
        // ===================================================================== TMA producer: streamed weights
        if (!p.resident && elect_one()) {
            int bi = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                for (int ch = 0; ch < p.nchunks; ++ch)
                    for (int tap = 0; tap < ntaps; ++tap, ++bi) {
                        const int bs = bi % p.b_slots;
                        mbar_wait(&b_empty[bs], ((bi / p.b_slots) & 1) ^ 1);
                        uint8_t* b = b_ring + bs * p.b_slot;
                        mbar_expect_tx(&b_full[bs], p.b_tx << p.x3);
                        tma_load_3d(b, &tmB, &b_full[bs], ch * HK, cout0, CT ? tap * p.S : tap);
                        if (p.x3) tma_load_3d(b + p.b_half, &tmB2, &b_full[bs], ch * HK, cout0, CT ? tap * p.S : tap);
                    }
        }
    } else if (warp == 8) {
        // ===================================================================== MMA issuer (one elected thread)
        if (elect_one()) {
            // Descriptors are kept as 64-bit values and advanced with plain adds (the 14-bit start-address field never
            // carries: shared memory is < 256 KB): the unrolled body is two 64-bit uniform adds + one UTCHMMA per MMA.
            const int n_mma = CT ? p.b_rows : ((n_this + 15) & ~15);
            const uint32_t idesc = umma_idesc_tf32(128, n_mma);
            const uint32_t idesc2 = umma_idesc_bf16(128, n_mma);   // cross terms: kind::f16, BF16 x BF16, K = 16
            const uint64_t a_desc0 = umma_desc(a_ring), b_desc0 = umma_desc(b_ring);
            const int RB = p.RB, nchunks = p.nchunks, a_stages = p.a_stages, b_slots = p.b_slots, npad = p.acc_cols;
            const uint64_t a_stage16 = (uint64_t)(p.a_stage >> 4), b_slot16 = (uint64_t)(p.b_slot >> 4);
            const uint64_t a_half16 = (uint64_t)(p.a_half >> 4), b_half16 = (uint64_t)(p.b_half >> 4);
            const int nk_last = ((p.Cin_p - (nchunks - 1) * HK) + 7) >> 3;
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
#pragma unroll 1
            for (int tile = blockIdx.x, lt = 0; tile < p.ntiles; tile += gridDim.x, ++lt) {
                const int buf = lt & 1;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dbase = tmem_base + (uint32_t)(buf * RB * npad);
#pragma unroll 1
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(X3 ? &a_split[as] : &a_full[as], aphase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int nk = ch == nchunks - 1 ? nk_last : 4;
                    const uint64_t a_desc = a_desc0 + (uint64_t)as * a_stage16;
                    const uint32_t acc_first = ch ? 1u : 0u;
                    if (RES) {
                        if (lt == 0) {   // resident weights: this chunk's taps land once per CTA
#pragma unroll 1
                            for (int tap = 0; tap < (CT ? R : R * S); ++tap) mbar_wait(&b_full[ch * (CT ? R : R * S) + tap], 0);
                        }
                        const uint64_t bd = b_desc0 + (uint64_t)(ch * (CT ? R : R * S)) * b_slot16;
#define HALO_ISSUE_RES(NKK)                                                                                               \
    do {                                                                                                                  \
        if (CT || RB == 1) issue_taps_resident<R, S, X3, NKK, 1, CT>(dbase, npad, a_desc, bd, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first); \
        else if (RB == 2) issue_taps_resident<R, S, X3, NKK, 2, false>(dbase, npad, a_desc, bd, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first); \
        else issue_taps_resident<R, S, X3, NKK, 4, false>(dbase, npad, a_desc, bd, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first); \
    } while (0)
                        if (nk == 4) HALO_ISSUE_RES(4);
                        else if (nk == 2) HALO_ISSUE_RES(2);
                        else if (nk == 1) HALO_ISSUE_RES(1);
                        else HALO_ISSUE_RES(3);
#undef HALO_ISSUE_RES
                    } else {
                        if (nk == 4) issue_taps_streaming<R, S, X3, 4, CT>(dbase, RB, npad, a_desc, b_desc0, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first, b_full, b_empty, bs, bphase, b_slots);
                        else if (nk == 2) issue_taps_streaming<R, S, X3, 2, CT>(dbase, RB, npad, a_desc, b_desc0, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first, b_full, b_empty, bs, bphase, b_slots);
                        else if (nk == 1) issue_taps_streaming<R, S, X3, 1, CT>(dbase, RB, npad, a_desc, b_desc0, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first, b_full, b_empty, bs, bphase, b_slots);
                        else issue_taps_streaming<R, S, X3, 3, CT>(dbase, RB, npad, a_desc, b_desc0, a_half16, b_half16, b_slot16, idesc, idesc2, acc_first, b_full, b_empty, bs, bphase, b_slots);
                    }
                    umma_commit(&a_empty[as]);
                    if (++as == a_stages) { as = 0; aphase ^= 1; }
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp == 4) {
        // ===================================================================== residual producer
        if (p.has_resid && elect_one()) {
            int gi = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
                const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
                const int w0 = tw * p.TWv, h0 = th * 8 * p.RB;
                for (int rb = 0; rb < p.RB; ++rb)
                    for (int g = 0; g < ngroups; ++g, ++gi) {
                        const int sb = gi % p.nstg;
                        mbar_wait(&r_empty[sb], ((gi / p.nstg) & 1) ^ 1);
                        mbar_expect_tx(&r_full[sb], p.stg_bytes);
                        tma_load_4d(rbuf + sb * p.stg_bytes, &tmR, &r_full[sb], cout0 + g * 32, w0, h0 + 8 * rb, n);
                    }
            }
        }
    } else if (warp == 6 || warp == 7 || warp == 9 || warp == 10) {
        // ===================================================================== cross-operand splitters (4 warps)
        // One 16-byte chunk (4 channels of one pixel row) per lane and piece, consecutive lanes on consecutive chunks
        // (conflict-free).  The 8 channels of a K step are two chunks that the 128B swizzle keeps adjacent
        // (chunk ^ (row & 7) flips only bit 0 inside the pair; odd rows hold them swapped) = lanes l and l ^ 1.
        // In place: hi = rna_tf32(v).  Cross tile, same swizzled position: logical chunk 0 of the pair = bf16(v - hi)
        // of the 8 channels, logical chunk 1 = bf16(v) of the 8 channels; the lane pair trades halves by shuffle.
        if (X3) {
            const int tid = (warp >= 9 ? warp - 7 : warp - 6) * 32 + lane;      // 0..127
            const int n4 = p.a_half >> 4;                                        // multiple of 128 (16 chunks x 8 rows)
            int ai = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                for (int ch = 0; ch < p.nchunks; ++ch, ++ai) {
                    const int as = ai % p.a_stages;
                    mbar_wait(&a_full[as], (ai / p.a_stages) & 1);
                    float4* hi = reinterpret_cast<float4*>(a_ring + as * p.a_stage);
                    uint4* cr = reinterpret_cast<uint4*>(a_ring + as * p.a_stage + p.a_half);
                    // four independent pieces per thread and pass (the in-place store may alias the next load as far
                    // as the compiler knows: one piece per pass would run at one shared-memory round trip each)
                    for (int i0 = tid; i0 < n4; i0 += 512) {
                        float4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (i0 + 128 * j < n4) v[j] = hi[i0 + 128 * j];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = i0 + 128 * j;
                            if (i < n4) {                                         // warp-uniform (n4 % 128 == 0)
                                float4 h;
                                uint2 lo, top;
                                split_cross(v[j], h, lo.x, lo.y, top.x, top.y);
                                // this lane holds channels 0-3 of the K step iff chunk parity == row parity
                                const bool first = ((i ^ (i >> 3)) & 1) == 0;
                                uint2 send = first ? top : lo, recv;
                                recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                                recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                                hi[i] = h;
                                cr[i] = first ? make_uint4(lo.x, lo.y, recv.x, recv.y)      // bf16(v - hi), channels 0-7
                                              : make_uint4(recv.x, recv.y, top.x, top.y);   // bf16(v), channels 0-7
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_split[as]);
                }
        }
    } else if (warp < 4) {
        // ===================================================================== epilogue (warps 0-3)
        const int q = warp;
        const int m = q * 32 + lane;              // TMEM lane = GEMM row = 16 * tile row + tile column
        const int col = m & 15, row = m >> 4;
        const bool valid = col < p.TWv;
        const int srow = row * p.TWv + col;       // row of the dense {32 ch, TWv, 8} staging box
        const int sw = srow & 7;                  // 128B swizzle phase of that row (buffers are 1024-byte aligned)
        const bool leader = threadIdx.x == 0;
        int gi = 0;
        for (int tile = blockIdx.x, lt = 0; tile < p.ntiles; tile += gridDim.x, ++lt) {
            const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
            const int th = trem / p.tilesW, tw = trem - th * p.tilesW;
            const int w0 = tw * p.TWv, h0 = th * 8 * p.RB;
            const int buf = lt & 1;
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int rb = 0; rb < p.RB; ++rb) {
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.RB + rb) * p.acc_cols);
                for (int g = 0; g < ngroups; ++g, ++gi) {
                    const int sb = gi % p.nstg;
                    uint8_t* sbuf = stg + sb * p.stg_bytes;
                    if (p.nstg == 1) {
                        // single staging buffer: the store issued one group ago must have finished reading it
                        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        epi_bar();
                    }
                    // (two buffers: sbuf was handed to the TMA store two groups ago, and the leader waited for that
                    //  read to complete BEFORE the barrier of the previous group - one barrier per group suffices)
                    if (p.has_resid) mbar_wait(&r_full[sb], (gi / p.nstg) & 1);
                    const int cbase = g * 32;
                    const int cn = min(32, n_this - cbase);           // multiple of 4
                    float v[32];
                    if (CT) {
                        // E[m][s, co] -> y[m][co] = sum_s E[m + s][s, co]: pixel m + s is lane + s of this warp (two
                        // 16-pixel rows per warp; the lanes that would cross a row are the junk columns col >= TWv)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (hh == 1 && cn <= 16) break;
                            tmem_ld16(tacc + (uint32_t)(cbase + 16 * hh), v + 16 * hh);
#pragma unroll
                            for (int sx = 1; sx < S; ++sx) {
                                float u[16];
                                tmem_ld16(tacc + (uint32_t)(sx * p.Cout_p + cbase + 16 * hh), u);
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[16 * hh + j] += __shfl_down_sync(0xffffffffu, u[j], sx);
                            }
                        }
                    } else if (cn > 16) tmem_ld32(tacc + (uint32_t)cbase, v);
                    else tmem_ld16(tacc + (uint32_t)cbase, v);
                    // One uniform branch per PASS over the 32 channels, not per float4: an epilogue warp runs alone on its
                    // scheduler, so its time is its dependent-issue latency - the former loop (5 uniform branches in each of
                    // 8 iterations, ~500 instructions per group) took ~2200 clk per group and bounded the N >= 128 layers.
                    {
                        const float4* ssc = reinterpret_cast<const float4*>(s_scale + cbase);
                        const float4* ssh = reinterpret_cast<const float4*>(s_shift + cbase);
                        if (p.scale) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t = ssc[j];
                                v[4 * j] *= t.x; v[4 * j + 1] *= t.y; v[4 * j + 2] *= t.z; v[4 * j + 3] *= t.w;
                            }
                        }
                        if (p.shift) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t = ssh[j];
                                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                            }
                        }
                        if (p.has_resid && valid) {
                            const float4* rrow = reinterpret_cast<const float4*>(rbuf + sb * p.stg_bytes + srow * 128);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (4 * j < cn) {
                                    const float4 t = rrow[j ^ sw];
                                    v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                                }
                            }
                        }
                        if (p.act == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                        } else if (p.act == 2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
                        }
                        if (valid) {
                            float4* orow = reinterpret_cast<float4*>(sbuf + srow * 128);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (4 * j < cn) orow[j ^ sw] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> TMA store reads
                    __syncwarp();
                    if (p.has_resid && lane == 0) mbar_arrive(&r_empty[sb]);       // residual buffer consumed
                    // the buffer written by the NEXT group must have been read by its store (nstg groups ago): with
                    // two buffers that is the store issued one group ago, with three the one before it
                    if (leader && p.nstg == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    if (leader && p.nstg == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    epi_bar();
                    if (leader) {
                        // sub-pixel pass: the output is seen as (c, px, w, py, n * H + h) with the full-resolution strides
                        if (p.ups) tma_store_5d(&tmY, sbuf, cout0 + cbase, p.px, w0, p.py, n * p.Ho + h0 + 8 * rb);
                        else tma_store_4d(&tmY, sbuf, cout0 + cbase, w0, h0 + 8 * rb, n);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            // every accumulator of this tile has been read: hand the TMEM buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

}  // namespace

static thread_local int t_hx3 = 0;
static thread_local int t_hct = 0;   // column-taps-on-N requested for this call (set by the exported entry point)
static thread_local int* t_hplan = nullptr;

// Column taps on N pay when the MMA phase of the plain schedule outweighs the epilogue: CT divides the operand rows
// read per K step by ~2 (3x3) but reads S accumulator column groups per output group and forces RB = 1.  Per-tile
// model with constants measured on B200 (profiles/r2_conv_micro_v8..v11): UTCHMMA = 18 + 0.35 * (M + N) clk, epilogue
// ~1500 clk per 32-channel group of a 128-pixel tile (+300 per extra column tap in CT mode); the slower of the two
// phases bounds a tile.  MONKEY_B200_HALO_CT = 0 never, 2 whenever eligible (experiments).
static bool halo_wants_ct(int R, int S, int Cin_p, int Cout_p, int x3) {
    static int ct_env = -1;
    if (ct_env < 0) {
        const char* e = getenv("MONKEY_B200_HALO_CT");
        ct_env = e ? atoi(e) : 1;
    }
    if (!ct_env || S <= 1 || Cout_p > 128 || Cout_p % 16 || S * Cout_p > 256) return false;
    if (ct_env == 2) return true;
    const int ksteps = (Cin_p + 7) / 8, groups = (Cout_p + 31) / 32;
    const double passes = x3 ? 2.0 : 1.0;
    const double mma_plain = passes * R * S * ksteps * (18.0 + 0.35 * (128 + Cout_p));
    const double mma_ct = passes * R * ksteps * (18.0 + 0.35 * (128 + S * Cout_p));
    const double epi_plain = 1500.0 * groups, epi_ct = (1500.0 + 300.0 * (S - 1)) * groups;
    const double t_plain = mma_plain > epi_plain ? mma_plain : epi_plain;
    const double t_ct = mma_ct > epi_ct ? mma_ct : epi_ct;
    return t_ct < 0.95 * t_plain;
}

static thread_local int t_hups = -1;   // >= 0: sub-pixel pass (parity py * 2 + px) of the upsampled conv, set by mk_conv2d_tc_halo_ups

static int halo_conv_impl(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc,
                          int R, int S, int pad, const float* scale, const float* shift, const float* resid,
                          int ldr, int act, float slope, float* y, int Cout_p, int ldy, void* stream) {
    const int ups_par = t_hups;
    // sub-pixel pass: a 2x2 conv on the LOW-resolution grid (output tile grid = input grid), rows {h-1, h} for py = 0 and
    // {h, h+1} for py = 1 (same for columns): an asymmetric "padding" of 1 - py / 1 - px before the first tap
    const int Ho = ups_par >= 0 ? Hin : Hin + 2 * pad - R + 1, Wo = ups_par >= 0 ? Win : Win + 2 * pad - S + 1;
    if (Cin_p % 4 || ldx % 4 || Cout_p % 4 || ldy % 4 || (resid && ldr % 4) || R != S ||
        (R != 1 && R != 3 && R != 4 && !(R == 2 && ups_par >= 0)) || (ups_par >= 0 && (R != 2 || resid || Hin % 8)) ||
        Ho < 1 || Wo < 1) {
        mk_set_error("mk_conv2d_tc_halo: outside the halo kernel's envelope");
        return -2;
    }
    HP p;
    p.x3 = t_hx3;
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout_p = Cout_p; p.Cin_p = Cin_p; p.R = R; p.S = S;
    p.ups = ups_par >= 0 ? 1 : 0;
    p.py = p.ups ? (ups_par >> 1) : 0;
    p.px = p.ups ? (ups_par & 1) : 0;
    p.pad_h = p.ups ? 1 - p.py : pad;
    p.pad_w = p.ups ? 1 - p.px : pad;
    p.TWv = 16 - (S - 1);
    p.tilesW = (Wo + p.TWv - 1) / p.TWv;
    p.nchunks = (Cin_p + HK - 1) / HK;
    const int n_tile = Cout_p < 128 ? Cout_p : 128;
    p.ct = t_hct;   // column taps stacked on N (see the header); eligibility checked by halo_wants_ct
    p.b_rows = p.ct ? S * Cout_p : ((n_tile + 15) & ~15);
    p.b_half = p.b_rows * 128;
    p.b_tx = p.b_half;
    p.b_slot = p.b_half << p.x3;
    p.ngroups = (n_tile + 31) / 32;
    p.npad = p.b_rows;
    p.stg_bytes = p.TWv * 8 * 128;
    p.has_resid = resid ? 1 : 0;
    p.scale = scale; p.shift = shift; p.act = act; p.slope = slope;
    const int cout_tiles = (Cout_p + 127) / 128;
    const int sms = mk_num_sms();
    const int nslot_chunk = p.ct ? R : R * S;              // weight slots per channel chunk
    const int nB = nslot_chunk * p.nchunks;
    // K steps of one tap summed over the chunks, and the weight bytes one pass over all (tap, chunk) pairs fetches
    int ksteps = 0;
    for (int ch = 0; ch < p.nchunks; ++ch) ksteps += ((Cin_p - ch * HK > HK ? HK : Cin_p - ch * HK) + 7) >> 3;
    const double w_bytes = (double)nslot_chunk * Cin_p * p.b_rows * 4 * (p.x3 ? 2 : 1);   // hi + cross operand
    // Planner: for RB in {1, 2, 4} and {2, 1} staging buffers find whether the weights can stay resident, and model the
    // time per 8-row block as max(MMA floor, L2->SM bytes / 40 B per clk) x the wave-quantisation loss of the static
    // tile striding.  Smallest modelled time wins; ties go to the smaller RB.
    double best_score = 0.0;
    int best_rb = 0, best_nstg = 0, best_res = 0;
    static int force_rb = -1;   // experiments: MONKEY_B200_HALO_RB = 1 | 2 | 4 pins the row-block count
    if (force_rb < 0) {
        const char* e = getenv("MONKEY_B200_HALO_RB");
        force_rb = e ? atoi(e) : 0;
    }
    for (int rb = 1; rb <= (p.ct ? 1 : 4); rb <<= 1) {
        if (!p.ct && force_rb && rb != force_rb && !(rb == 1 && 2 * force_rb * p.npad > 512)) continue;
        if (2 * rb * p.npad > 512) break;                        // double-buffered accumulators in TMEM
        if (rb > 1 && 8 * (rb / 2) >= Ho) break;                           // the extra row-blocks would all be empty
        if (p.ups && Ho % (8 * rb)) break;                                  // (n, h) are one merged store dimension: no partial tiles
        const int halo_rows = 8 * rb + R;                                   // (R-1) halo rows + 1 overrun row
        const int a_stage = (halo_rows * 16 * 128) << p.x3;
        int nstg = 0, res = 0, bslots = 0;
        static int max_nstg = -1;   // experiments: MONKEY_B200_HALO_NSTG caps the staging buffers (default 3)
        if (max_nstg < 0) {
            const char* e = getenv("MONKEY_B200_HALO_NSTG");
            max_nstg = e ? atoi(e) : 3;
            if (max_nstg < 1 || max_nstg > 3) max_nstg = 3;
        }
        // preference: resident weights with 3 or 2 staging buffers (3: every TMA store has two group times to drain),
        // a weight ring with 2 (shared memory buys ring depth before a third staging buffer), then either with one
        auto try_resident = [&](int ns) {
            const int fixed = ns * (1 + p.has_resid) * p.stg_bytes + H_FIXED;
            if (!nstg && nB <= MAXB && nB * p.b_slot + 2 * a_stage + fixed <= H_SMEM_MAX) { nstg = ns; res = 1; bslots = nB; }
        };
        auto try_ring = [&](int ns) {
            const int fixed = ns * (1 + p.has_resid) * p.stg_bytes + H_FIXED;
            if (!nstg && (ns == 1 ? 2 : 3) * p.b_slot + 2 * a_stage + fixed <= H_SMEM_MAX) {
                nstg = ns; res = 0;
                bslots = (H_SMEM_MAX - fixed - 2 * a_stage) / p.b_slot;
                if (bslots > 8) bslots = 8;
            }
        };
        for (int ns = max_nstg; ns >= 2; --ns) try_resident(ns);
        if (max_nstg >= 2) try_ring(2);
        try_resident(1);
        try_ring(1);
        if (!nstg) break;
        const double mma_clk = (double)rb * nslot_chunk * ksteps * (128.0 * p.b_rows / 256.0) * (p.x3 ? 2 : 1);
        const double a_bytes = (double)halo_rows * 16 * Cin_p * 4;
        const double l2 = a_bytes + (res ? 0.0 : w_bytes) + (double)rb * (1 + p.has_resid) * p.TWv * 8 * n_tile * 4;
        const long long tiles = (long long)p.tilesW * ((Ho + 8 * rb - 1) / (8 * rb)) * N;
        if (rb > 1 && tiles * cout_tiles < sms) break;                      // would not even fill one wave
        int gx = sms / cout_tiles < 1 ? 1 : sms / cout_tiles;
        const double waves = (double)tiles / gx;
        const double quant = (double)((tiles + gx - 1) / gx) / (waves > 1e-9 ? waves : 1e-9);
        const double rows_eff = (double)Ho / ((Ho + 8 * rb - 1) / (8 * rb) * 8.0 * rb);
        // a streamed weight slot is consumed in (rb * K steps) MMAs but takes a full L2 round trip (~2800 clk) to
        // refill: with `bslots` slots in flight the ring sustains one slot per 2800 / bslots clk
        const double slot_clk = mma_clk / nB;
        const double ring_clk = res ? 0.0 : (1500.0 / bslots) * nB;
        double t = mma_clk > l2 / 40.0 ? mma_clk : l2 / 40.0;
        if (ring_clk > t) t = ring_clk;
        if (p.x3) t += 3000.0;   // 3xTF32: fixed cost of a super-tile (halo load -> split -> MMA chain, two stages deep);
                                 // calibrated on the 4->64 and 48->48 layers, where RB = 2 measured 1.8x / 1.09x faster.
                                 // In 1xTF32 RB = 1 with resident weights measured best on every layer tried.
        (void)slot_clk;
        const double score = t / rb * quant / rows_eff;
        if (!best_rb || score < best_score * 0.97) { best_rb = rb; best_score = score; best_nstg = nstg; best_res = res; }
    }
    if (!best_rb) {
        mk_set_error("mk_conv2d_tc_halo: no shared-memory plan for this layer");
        return -2;
    }
    (void)best_score;
    p.RB = best_rb;
    p.nstg = best_nstg;
    p.resident = best_res;
    p.halo_rows = 8 * p.RB + R;
    p.a_half = p.halo_rows * 16 * 128;
    p.a_stage = p.a_half << p.x3;
    p.tilesH = (Ho + 8 * p.RB - 1) / (8 * p.RB);
    p.ntiles = p.tilesW * p.tilesH * N;
    // envelope of the persistent launch: at least one tile per CTA, and tiles that are mostly inside the image
    // (small / deep levels: 16-wide rows of a 16x16 image would be 57 % junk; mk_conv2d_tc's split-K serves them)
    const double useful = (double)Ho * Wo / ((double)p.tilesH * 8 * p.RB * p.tilesW * p.TWv);
    if ((long long)p.ntiles * cout_tiles < sms || useful < 0.7) {
        mk_set_error("mk_conv2d_tc_halo: %d tiles, %.0f %% useful: left to mk_conv2d_tc", p.ntiles, 100.0 * useful);
        return -2;
    }
    const int fixed = p.nstg * (1 + p.has_resid) * p.stg_bytes + H_FIXED;
    int budget = H_SMEM_MAX - fixed;
    if (p.resident) {
        p.b_slots = nB;
        p.a_stages = (budget - nB * p.b_slot) / p.a_stage;
    } else {
        // streaming: two halo stages, the rest of the budget holds weight slots (at most 8: more buy nothing, and a
        // third halo stage is worth more)
        p.a_stages = 2;
        p.b_slots = (budget - 2 * p.a_stage) / p.b_slot;
        if (p.b_slots > 8) {
            p.b_slots = 8;
            p.a_stages = (budget - 8 * p.b_slot) / p.a_stage;
        }
    }
    if (p.a_stages > MAXA) p.a_stages = MAXA;
    MK_REQUIRE(p.a_stages >= 2 && (p.resident || p.b_slots >= 2), "mk_conv2d_tc_halo: ring plan failed");
    p.acc_cols = p.npad;
    const int cols = 2 * p.RB * p.acc_cols;
    p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    const int smem_bytes = p.a_stages * p.a_stage + p.b_slots * p.b_slot + fixed;
    MK_REQUIRE(smem_bytes <= H_SMEM_MAX, "mk_conv2d_tc_halo: shared memory plan exceeds 227 KB (%d)", smem_bytes);
    int grid_x = sms / cout_tiles;
    if (grid_x < 1) grid_x = 1;
    if (grid_x > p.ntiles) grid_x = p.ntiles;
    if (t_hplan) {
        int* o = t_hplan;
        o[0] = grid_x; o[1] = cout_tiles; o[2] = p.RB; o[3] = smem_bytes; o[4] = p.a_stages; o[5] = p.b_slots;
        o[6] = p.resident; o[7] = p.tmem_cols; o[8] = p.ntiles; o[9] = p.halo_rows; o[10] = p.a_stage; o[11] = p.b_slot;
        o[12] = p.TWv; o[13] = p.ngroups; o[14] = p.npad; o[15] = p.x3 | (p.nstg << 1) | (p.ct << 3);
        return 0;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_tc_halo: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tmA, tmB, tmB2, tmY, tmR;
    cuuint32_t es4[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_p, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)Win * ldx * 4, (cuuint64_t)Hin * Win * ldx * 4};
        cuuint32_t box[4] = {(cuuint32_t)HK, 16, (cuuint32_t)p.halo_rows, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es4,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: activation tensor map rejected (%d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_p, (cuuint64_t)Cout_p, (cuuint64_t)(R * S)};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_p * 4, (cuuint64_t)Cin_p * Cout_p * 4};
        // CT: one box = the S column taps of a tap row, rows s * Cout_p + co
        cuuint32_t box[3] = {(cuuint32_t)HK, (cuuint32_t)(p.ct ? Cout_p : p.b_rows), (cuuint32_t)(p.ct ? S : 1)};
        cuuint32_t es[3] = {1, 1, 1};
        // sub-pixel pass: the mode-4 pack is [parity][2x2 tap][Cout_p][Cin_p] (16 taps) + its cross operand behind it
        const size_t par_taps = p.ups ? (size_t)(p.py * 2 + p.px) * 4 : 0, all_taps = p.ups ? 16 : (size_t)R * S;
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                            const_cast<float*>(wpack_tc) + par_taps * Cout_p * Cin_p, dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: weight tensor map rejected (%d)", (int)r);
        tmB2 = tmB;
        if (p.x3) {
            // cross operand of the weights (mk_pack_weight mode | 16): [tap][Cout_p][Cin_p rounded up to 8] 4-byte slots
            // (two bf16 each) behind the R*S*Cout_p*Cin_p floats of the hi half; moved as opaque 32-bit words
            const cuuint64_t cin8 = (cuuint64_t)((Cin_p + 7) & ~7);
            cuuint64_t dims2[3] = {cin8, (cuuint64_t)Cout_p, (cuuint64_t)(R * S)};
            cuuint64_t strides2[2] = {cin8 * 4, cin8 * Cout_p * 4};
            r = encode(&tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                       const_cast<float*>(wpack_tc) + all_taps * Cout_p * Cin_p + par_taps * Cout_p * cin8, dims2, strides2, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: cross-operand tensor map rejected (%d)", (int)r);
        }
    }
    if (p.ups) {
        // y[n][2h + py][2w + px][c] as (c, px, w, py, n * H + h): the frame stride (2H)(2W) ld equals H times the stride of h
        cuuint64_t dims[5] = {(cuuint64_t)Cout_p, 2, (cuuint64_t)Wo, 2, (cuuint64_t)N * Ho};
        cuuint64_t strides[4] = {(cuuint64_t)ldy * 4, (cuuint64_t)2 * ldy * 4, (cuuint64_t)2 * Wo * ldy * 4,
                                 (cuuint64_t)4 * Wo * ldy * 4};
        cuuint32_t box[5] = {32, 1, (cuuint32_t)p.TWv, 1, 8};
        cuuint32_t es5[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, y, dims, strides, box, es5,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: sub-pixel output tensor map rejected (%d)", (int)r);
        tmR = tmY;
    } else
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && !resid) { tmR = tmY; break; }
        const float* base = which ? resid : y;
        const int ld = which ? ldr : ldy;
        cuuint64_t dims[4] = {(cuuint64_t)Cout_p, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)Wo * ld * 4, (cuuint64_t)Ho * Wo * ld * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.TWv, 8, 1};
        CUresult r = encode(which ? &tmR : &tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims,
                            strides, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: output tensor map rejected (%d)", (int)r);
    }
    dim3 grid((unsigned)grid_x, (unsigned)cout_tiles, 1);
    cudaError_t le = cudaSuccess;
#define HALO_LAUNCH(RR, SS, XX, RE, CC)                                                                              \
    do {                                                                                                             \
        static unsigned long long attr_done = 0;                                                                     \
        if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {                                         \
            le = cudaFuncSetAttribute(k_conv_halo<RR, SS, XX, RE, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      H_SMEM_MAX);                                                                   \
            if (le == cudaSuccess) attr_done |= attr_bit;                                                            \
        }                                                                                                            \
        if (le == cudaSuccess)                                                                                       \
            k_conv_halo<RR, SS, XX, RE, CC><<<grid, H_THREADS, smem_bytes, (cudaStream_t)stream>>>(tmA, tmB, tmB2, tmY, tmR, p); \
    } while (0)
#define HALO_RS(XX, RE)                                                                     \
    do {                                                                                    \
        if (R == 3 && S == 3) { if (p.ct) HALO_LAUNCH(3, 3, XX, RE, true); else HALO_LAUNCH(3, 3, XX, RE, false); } \
        else if (R == 4 && S == 4) { if (p.ct) HALO_LAUNCH(4, 4, XX, RE, true); else HALO_LAUNCH(4, 4, XX, RE, false); } \
        else if (R == 2 && S == 2) { if (p.ct) HALO_LAUNCH(2, 2, XX, RE, true); else HALO_LAUNCH(2, 2, XX, RE, false); } \
        else HALO_LAUNCH(1, 1, XX, RE, false);                                              \
    } while (0)
    if (p.x3) { if (p.resident) HALO_RS(true, true); else HALO_RS(true, false); }
    else { if (p.resident) HALO_RS(false, true); else HALO_RS(false, false); }
#undef HALO_RS
#undef HALO_LAUNCH
    if (le != cudaSuccess) { mk_set_error("mk_conv2d_tc_halo: smem attribute: %s", cudaGetErrorString(le)); return (int)le; }
    return mk_check_launch("mk_conv2d_tc_halo");
}

MK_EXPORT int mk_conv2d_tc_halo(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc,
                                int R, int S, int pad, const float* scale, const float* shift, const float* resid,
                                int ldr, int act, float slope, float* y, int Cout_p, int ldy, void* stream) {
    t_hct = (R == S && halo_wants_ct(R, S, Cin_p, Cout_p, t_hx3)) ? 1 : 0;
    int rc = halo_conv_impl(x, N, Hin, Win, Cin_p, ldx, wpack_tc, R, S, pad, scale, shift, resid, ldr, act, slope, y,
                            Cout_p, ldy, stream);
    if (rc == -2 && t_hct) {   // no shared-memory plan with the S-fold weight slots: the plain schedule
        t_hct = 0;
        rc = halo_conv_impl(x, N, Hin, Win, Cin_p, ldx, wpack_tc, R, S, pad, scale, shift, resid, ldr, act, slope, y,
                            Cout_p, ldy, stream);
    }
    t_hct = 0;
    return rc;
}

// conv3x3(nearest_x2(x)), pad 1 (modules/util.py:84-85), as four sub-pixel 2x2 halo-window passes on the low-resolution
// grid: `wpack_ups` is the mode-4 pack of mk_pack_weight ([parity][2x2 tap][Cout_p][Cin_p], | 8 for reference precision),
// y is [N][2 Hin][2 Win][ldy].  Returns -2 (before launching anything) outside the envelope (Hin % 8, few tiles ...).
MK_EXPORT int mk_conv2d_tc_halo_ups(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_ups,
                                    const float* scale, const float* shift, int act, float slope, float* y, int Cout_p,
                                    int ldy, void* stream) {
    int rc = 0;
    for (int par = 0; par < 4 && rc == 0; ++par) {
        t_hups = par;
        t_hct = halo_wants_ct(2, 2, Cin_p, Cout_p, t_hx3) ? 1 : 0;
        rc = halo_conv_impl(x, N, Hin, Win, Cin_p, ldx, wpack_ups, 2, 2, 0, scale, shift, nullptr, 0, act, slope, y, Cout_p,
                            ldy, stream);
        if (rc == -2 && t_hct) {
            t_hct = 0;
            rc = halo_conv_impl(x, N, Hin, Win, Cin_p, ldx, wpack_ups, 2, 2, 0, scale, shift, nullptr, 0, act, slope, y,
                                Cout_p, ldy, stream);
        }
        if (rc == -2 && par > 0) {   // all four passes share one plan: a later refusal would leave y half written
            mk_set_error("mk_conv2d_tc_halo_ups: parity %d refused after parity 0 was launched", par);
            rc = -1;
        }
    }
    t_hups = -1;
    t_hct = 0;
    return rc;
}

MK_EXPORT int mk_conv2d_tc_halo_ups_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_ups,
                                       const float* scale, const float* shift, int act, float slope, float* y,
                                       int Cout_p, int ldy, void* stream) {
    t_hx3 = 1;
    const int rc = mk_conv2d_tc_halo_ups(x, N, Hin, Win, Cin_p, ldx, wpack_ups, scale, shift, act, slope, y, Cout_p, ldy,
                                         stream);
    t_hx3 = 0;
    return rc;
}

MK_EXPORT int mk_conv2d_tc_halo_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc,
                                   int R, int S, int pad, const float* scale, const float* shift, const float* resid,
                                   int ldr, int act, float slope, float* y, int Cout_p, int ldy, void* stream) {
    t_hx3 = 1;
    const int rc = mk_conv2d_tc_halo(x, N, Hin, Win, Cin_p, ldx, wpack_tc, R, S, pad, scale, shift, resid, ldr, act, slope,
                                     y, Cout_p, ldy, stream);
    t_hx3 = 0;
    return rc;
}

// Dry run of the planner (no device state touched): out[16] = grid.x, grid.y (cout tiles), RB, dynamic smem bytes,
// halo stages, weight slots, weights resident?, TMEM columns, tiles, halo rows, halo stage bytes, weight slot bytes,
// valid tile width, 32-channel output groups, accumulator columns, x3.  Returns -2 outside the envelope.
MK_EXPORT int mk_conv2d_tc_halo_plan(int N, int Hin, int Win, int Cin_p, int R, int S, int pad, int Cout_p, int has_resid,
                                     int x3, int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_tc_halo_plan: out is NULL");
    t_hplan = out;
    t_hx3 = x3 ? 1 : 0;
    static float dummy;
    const int rc = mk_conv2d_tc_halo(nullptr, N, Hin, Win, Cin_p, Cin_p, nullptr, R, S, pad, nullptr, nullptr,
                                     has_resid ? &dummy : nullptr, Cout_p, 0, 0.f, nullptr, Cout_p, Cout_p, nullptr);
    t_hplan = nullptr;
    t_hx3 = 0;
    return rc;
}

// Dry run of the four sub-pixel passes of the upsampled conv (they share one plan): same out[16] as above.
MK_EXPORT int mk_conv2d_tc_halo_ups_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int x3, int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_tc_halo_ups_plan: out is NULL");
    t_hplan = out;
    t_hx3 = x3 ? 1 : 0;
    const int rc = mk_conv2d_tc_halo_ups(nullptr, N, Hin, Win, Cin_p, Cin_p, nullptr, nullptr, nullptr, 0, 0.f, nullptr,
                                         Cout_p, Cout_p, nullptr);
    t_hplan = nullptr;
    t_hx3 = 0;
    return rc;
}
