// Tensor-core weight gradient for sm_100a:   dW[tap][ci][co] = sum_pixels dY[pix][co] * X[pix + tap][ci]
//
// GEMM view: D[M = co][N = ci] accumulated over K = pixels.  In NHWC both operands have their M/N index (channels)
// contiguous and the contraction index (pixels) strided, i.e. they are "MN-major" UMMA operands - tcgen05 takes them
// directly (a_major = b_major = 1 in the instruction descriptor), so no transposed copies of the activations exist:
//   * operand tile = 4-D TMA boxes {32 ch, TW, TH, TN} of PC = 64 pixels; each box is [64 pixels][128 B] in the 128B
//     swizzle with 32-byte atoms (TMA SWIZZLE_128B_ATOM_32B == UMMA SWIZZLE_128B_BASE32B, the only MN-major layout
//     32-bit operands may use) = one MN block of the canonical layout ((8,n),(4,k)) with LBO = box size, SBO = 512 B
//     (4 pixel rows).  Out-of-bounds pixels are zero-filled by the TMA unit = the convolution's zero padding.
//   * ONE CTA owns a pixel range and ALL filter taps that fit in TMEM (taps x round16(ci) <= 512 accumulator
//     columns): the dY tile of a pixel chunk is loaded once (A ring) and multiplied with the R*S shifted X tiles
//     (B ring), each tap accumulating into its own TMEM column range.  The previous version ran one tap per CTA:
//     dY was fetched R*S times, 4 dY boxes were staged whatever Cout was, and the full-resolution small-channel
//     layers (24 -> 24 @ 64x64 x 32 frames) took 150 us; now they are a single pass over X and dY.
//   * only the 32-channel boxes that exist are staged (ceil(co/32) A boxes, ceil(ci/32) B boxes).  The MMA still runs
//     M = 128: accumulator rows beyond the staged boxes are products of stale shared memory and are never read.
//   * reference precision (mk_conv2d_wgrad_tc_x3; scheme in conv_halo.cu / wgrad_halo.cu): both operands are
//     activations, so both rings carry a cross half behind the hi half of every slot; the four epilogue warps round
//     each landed slot to TF32 in place and write per pixel two 64-byte bf16 K rows ([lo | top] for dY, [top | lo] for
//     X: an MN-major SWIZZLE_64B operand with the fp32 box's offsets), published through a_split / b_split mbarriers;
//     the issuer runs one kind::tf32 MMA + one kind::f16 (BF16, K = 16 rows = 8 pixels) MMA per 8 pixels.
//   * grid = (co tiles, tap groups x ci tiles, pixel splits); splits combine with fp32 atomics (red) into the packed
//     gradient, whose layout [tap][Cin_p][Cout_p] makes the epilogue's per-column writes coalesced across lanes.
#include "tc_common.cuh"
#include "../../include/monkey_b200.h"

namespace {
using namespace mk_tc;

constexpr int PC = 64;                 // pixels per chunk (GEMM K per pipeline step)
constexpr int BOX_BYTES = PC * 128;    // one 32-channel box
constexpr int MAX_A = 4, MAX_B = 12;   // ring depths (upper bounds)
constexpr int WSMEM_MAX = 227 * 1024;

struct WgTcP {
    int N, Ho, Wo, Cout_p, Cin_p, R, S, pad;
    int TW, TH, TN, tilesW, tilesH, nchunks, chunks_per_split, n_ci_tiles;
    int taps_per_cta, n_tap_groups, npad;   // npad = accumulator columns per tap (round16 of the ci tile)
    int na_max, nb_max, a_slots, b_slots, tmem_cols;
    int x3;   // reference precision: slot = [hi boxes | cross boxes]
    float* dw;
};

__device__ __forceinline__ uint64_t umma_desc_mn(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)(BOX_BYTES >> 4) << 16;  // LBO: next 32-channel MN block
    d |= (uint64_t)(512 >> 4) << 32;        // SBO: next group of 4 pixel rows (the 32B-atom swizzle repeats every 4)
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B - the only MN-major layout tf32 operands may use
    return d;
}

__global__ void __launch_bounds__(256) k_wgrad_tc(const __grid_constant__ CUtensorMap tmDy,
                                                  const __grid_constant__ CUtensorMap tmX, const WgTcP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int A_HALF = p.na_max * BOX_BYTES, B_HALF = p.nb_max * BOX_BYTES;
    const int A_SLOT = A_HALF << p.x3, B_SLOT = B_HALF << p.x3;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + p.a_slots * A_SLOT;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + p.b_slots * B_SLOT);
    uint64_t* a_empty = a_full + MAX_A;
    uint64_t* b_full = a_empty + MAX_A;
    uint64_t* b_empty = b_full + MAX_B;
    uint64_t* a_split = b_empty + MAX_B;
    uint64_t* b_split = a_split + MAX_A;
    uint64_t* tmem_full = b_split + MAX_B;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * 128;
    const int tg = blockIdx.y / p.n_ci_tiles, ci0 = (blockIdx.y % p.n_ci_tiles) * 128;
    const int tap_begin = tg * p.taps_per_cta;
    const int ntaps = min(p.taps_per_cta, p.R * p.S - tap_begin);
    const int m_this = min(128, p.Cout_p - co0);
    const int n_this = min(128, p.Cin_p - ci0);          // multiple of 4; the MMA runs on the next multiple of 16
    const int na = (m_this + 31) >> 5, nb = (n_this + 31) >> 5;
    const int q0 = blockIdx.z * p.chunks_per_split;
    const int q1 = min(p.nchunks, q0 + p.chunks_per_split);
    const int nq = q1 - q0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_A; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_split[i], 4); }
        for (int i = 0; i < MAX_B; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&b_split[i], 4); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (elect_one()) {
            int bi = 0;
            for (int qi = 0; qi < nq; ++qi) {
                int q = q0 + qi;
                const int tw = q % p.tilesW; q /= p.tilesW;
                const int th = q % p.tilesH; q /= p.tilesH;
                const int w0 = tw * p.TW, h0 = th * p.TH, n0 = q * p.TN;
                const int as = qi % p.a_slots;
                mbar_wait(&a_empty[as], ((qi / p.a_slots) & 1) ^ 1);
                uint8_t* a = a_ring + as * A_SLOT;
                mbar_expect_tx(&a_full[as], na * BOX_BYTES);
                for (int j = 0; j < na; ++j) tma_load_4d(a + j * BOX_BYTES, &tmDy, &a_full[as], co0 + j * 32, w0, h0, n0);
                for (int t = 0; t < ntaps; ++t, ++bi) {
                    const int tap = tap_begin + t;
                    const int r = tap / p.S, s = tap - r * p.S;
                    const int bs = bi % p.b_slots;
                    mbar_wait(&b_empty[bs], ((bi / p.b_slots) & 1) ^ 1);
                    uint8_t* b = b_ring + bs * B_SLOT;
                    mbar_expect_tx(&b_full[bs], nb * BOX_BYTES);
                    for (int j = 0; j < nb; ++j)
                        tma_load_4d(b + j * BOX_BYTES, &tmX, &b_full[bs], ci0 + j * 32, w0 + s - p.pad, h0 + r - p.pad,
                                    n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // M = 128 (co), N = round16(n_this) (ci), both operands MN-major; tap t accumulates at column t * npad
        const uint32_t idesc = umma_idesc_tf32(128, (n_this + 15) & ~15) | (1u << 15) | (1u << 16);
        const uint32_t idesc_c = umma_idesc_bf16(128, (n_this + 15) & ~15) | (1u << 15) | (1u << 16);
        int bi = 0;
        for (int qi = 0; qi < nq; ++qi) {
            const int as = qi % p.a_slots;
            mbar_wait(p.x3 ? &a_split[as] : &a_full[as], (qi / p.a_slots) & 1);
            for (int t = 0; t < ntaps; ++t, ++bi) {
                const int bs = bi % p.b_slots;
                mbar_wait(p.x3 ? &b_split[bs] : &b_full[bs], (bi / p.b_slots) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_mn(a_ring + as * A_SLOT);
                    const uint64_t bdesc = umma_desc_mn(b_ring + bs * B_SLOT);
                    const uint32_t dcol = tmem_base + (uint32_t)(t * p.npad);
                    if (p.x3) {
                        // cross operands: same offsets, layout type SWIZZLE_128B_BASE32B (1) -> SWIZZLE_64B (4)
                        constexpr uint64_t FLIP = ((uint64_t)1 ^ (uint64_t)4) << 61;
                        const uint64_t acdesc = umma_desc_mn(a_ring + as * A_SLOT + A_HALF) ^ FLIP;
                        const uint64_t bcdesc = umma_desc_mn(b_ring + bs * B_SLOT + B_HALF) ^ FLIP;
#pragma unroll
                        for (int k = 0; k < PC / 8; ++k) {
                            umma_tf32(dcol, adesc + 64 * k, bdesc + 64 * k, idesc, (qi | k) ? 1u : 0u);
                            umma_bf16(dcol, acdesc + 64 * k, bcdesc + 64 * k, idesc_c, 1u);
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < PC / 8; ++k)  // 8 pixel rows = 1024 B = 64 sixteen-byte units
                        umma_tf32(dcol, adesc + 64 * k, bdesc + 64 * k, idesc, (qi | k) ? 1u : 0u);
                    umma_commit(&b_empty[bs]);
                    if (t == ntaps - 1) {
                        umma_commit(&a_empty[as]);
                        if (qi == nq - 1) umma_commit(tmem_full);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && nq > 0) {
        // ===================================================================== operand split (x3), then epilogue
        if (p.x3) {
            const int tid = threadIdx.x - 128;
            // lanes l, l ^ 1 hold the 8 channels of one 32-byte swizzle unit; the even lane assembles the pixel's first
            // 64-byte K row, the odd lane the second one (A side: [lo | top], B side: [top | lo])
            auto split_slot = [&](uint8_t* slot, int nboxes, int half, bool a_side) {
                float4* hi = reinterpret_cast<float4*>(slot);
                uint4* cr = reinterpret_cast<uint4*>(slot + half);
                const int n4 = nboxes * (BOX_BYTES / 16);       // multiple of 128
                const int odd = tid & 1;
                const bool lo_row = a_side ? !odd : odd;        // this lane assembles the bf16(v - hi) row
                for (int i = tid; i < n4; i += 128) {
                    float4 v = hi[i], h;
                    uint2 lo, top;
                    split_cross(v, h, lo.x, lo.y, top.x, top.y);
                    uint2 send = lo_row ? top : lo, mine = lo_row ? lo : top, recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    hi[i] = h;
                    cr[(i & ~7) + 4 * odd + ((i & 7) >> 1)] =
                        odd ? make_uint4(recv.x, recv.y, mine.x, mine.y) : make_uint4(mine.x, mine.y, recv.x, recv.y);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            };
            int bi = 0;
            for (int qi = 0; qi < nq; ++qi) {
                const int as = qi % p.a_slots;
                mbar_wait(&a_full[as], (qi / p.a_slots) & 1);
                split_slot(a_ring + as * A_SLOT, na, A_HALF, true);
                if (lane == 0) mbar_arrive(&a_split[as]);
                for (int t = 0; t < ntaps; ++t, ++bi) {
                    const int bs = bi % p.b_slots;
                    mbar_wait(&b_full[bs], (bi / p.b_slots) & 1);
                    split_slot(b_ring + bs * B_SLOT, nb, B_HALF, false);
                    if (lane == 0) mbar_arrive(&b_split[bs]);
                }
            }
        }
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        const bool valid = co < p.Cout_p;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int t = 0; t < ntaps; ++t) {
            float* base = p.dw + ((long long)(tap_begin + t) * p.Cin_p + ci0) * p.Cout_p + co;
            for (int c = 0; c < n_this; c += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.npad + c), v);
                if (!valid) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (c + j >= n_this) break;  // ragged Cin_p: the extra columns are products with TMA zero fill
                    float* dst = base + (long long)(c + j) * p.Cout_p;
                    if (gridDim.z == 1) *dst = v[j];
                    else atomicAdd(dst, v[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

}  // namespace

static thread_local int* t_wplan_out = nullptr;  // set by mk_conv2d_wgrad_tc_plan for a dry run
static thread_local int t_wx3 = 0;               // set by mk_conv2d_wgrad_tc_x3

MK_EXPORT int mk_conv2d_wgrad_tc(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                 int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    if (Cin_p % 4 || Cout_p % 4 || ldx % 4 || ldy % 4) {
        mk_set_error("mk_conv2d_wgrad_tc: unsupported channel configuration");
        return -2;
    }
    WgTcP p;
    p.N = N; p.Ho = Hin + 2 * pad - R + 1; p.Wo = Win + 2 * pad - S + 1;
    MK_REQUIRE(p.Ho > 0 && p.Wo > 0, "mk_conv2d_wgrad_tc: empty output");
    p.Cout_p = Cout_p; p.Cin_p = Cin_p; p.R = R; p.S = S; p.pad = pad; p.dw = dwpack;
    p.TW = pow2_ceil(p.Wo) < 16 ? pow2_ceil(p.Wo) : 16;
    p.TH = pow2_ceil(p.Ho) < PC / p.TW ? pow2_ceil(p.Ho) : PC / p.TW;
    p.TN = PC / (p.TW * p.TH);
    p.tilesW = (p.Wo + p.TW - 1) / p.TW; p.tilesH = (p.Ho + p.TH - 1) / p.TH;
    p.nchunks = p.tilesW * p.tilesH * ((N + p.TN - 1) / p.TN);
    p.n_ci_tiles = (Cin_p + 127) / 128;
    const int co_tiles = (Cout_p + 127) / 128;
    // accumulator columns per tap, taps per CTA (TMEM has 512 columns)
    const int n_tile = Cin_p < 128 ? Cin_p : 128;
    p.npad = (n_tile + 15) & ~15;
    p.taps_per_cta = 512 / p.npad < R * S ? 512 / p.npad : R * S;
    p.n_tap_groups = (R * S + p.taps_per_cta - 1) / p.taps_per_cta;
    p.taps_per_cta = (R * S + p.n_tap_groups - 1) / p.n_tap_groups;  // balance the groups (9 taps, 4 fit -> 3+3+3)
    p.n_tap_groups = (R * S + p.taps_per_cta - 1) / p.taps_per_cta;
    const int cols = p.taps_per_cta * p.npad;
    p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    p.na_max = ((Cout_p < 128 ? Cout_p : 128) + 31) / 32;
    p.nb_max = (n_tile + 31) / 32;
    // pixel splits: fill the machine (about one CTA per SM - the 512-column case allows no second resident CTA)
    const long long tiles = (long long)co_tiles * p.n_ci_tiles * p.n_tap_groups;
    const int sms = mk_num_sms();
    p.x3 = t_wx3;
    // (wave-aware: tc_common.cuh:pick_splits; two CTAs share an SM only with <= 256 TMEM columns and 1x slots)
    long long splits = pick_splits(tiles, p.nchunks, ((p.tmem_cols <= 256 && !p.x3) ? 2LL : 1LL) * sms, 4.0, 65535);
    p.chunks_per_split = (int)mk_cdiv(p.nchunks, splits);
    splits = mk_cdiv(p.nchunks, p.chunks_per_split);
    // rings: never deeper than the loops; within ~100 KB when two CTAs can share an SM, ~200 KB otherwise
    const int budget = (p.tmem_cols <= 256 && tiles * splits > sms && !p.x3) ? 100 * 1024 : 200 * 1024;
    const int a_slot = (p.na_max * BOX_BYTES) << p.x3, b_slot = (p.nb_max * BOX_BYTES) << p.x3;
    p.a_slots = p.chunks_per_split < 2 ? 1 : 2;
    if (p.x3 && p.a_slots * a_slot + 2 * b_slot + 1536 > WSMEM_MAX) p.a_slots = 1;   // 3xTF32 slots are twice as big
    int bs = (budget - p.a_slots * a_slot) / b_slot;
    const long long b_loads = (long long)p.chunks_per_split * p.taps_per_cta;
    if (bs > MAX_B) bs = MAX_B;
    if (bs > b_loads) bs = (int)b_loads;
    if (bs < 2) bs = 2;
    p.b_slots = bs;
    // the M = 128 MMA reads four 32-channel A boxes whatever Cout is: the rows beyond the staged boxes are never
    // used, but the addresses must lie inside this CTA's shared-memory allocation
    int ring_bytes = p.a_slots * a_slot + p.b_slots * b_slot + 512 /*barriers*/;
    const int a_reach = (p.a_slots - 1) * a_slot + (p.x3 ? p.na_max * BOX_BYTES : 0) + 4 * BOX_BYTES;
    if (ring_bytes < a_reach) ring_bytes = a_reach;
    const int smem_bytes = ring_bytes + 1024 /*align*/;
    MK_REQUIRE(smem_bytes <= WSMEM_MAX, "mk_conv2d_wgrad_tc: shared memory plan exceeds 227 KB (%d)", smem_bytes);
    if (t_wplan_out) {  // dry run (mk_conv2d_wgrad_tc_plan)
        int* o = t_wplan_out;
        o[0] = co_tiles; o[1] = p.n_tap_groups * p.n_ci_tiles; o[2] = (int)splits; o[3] = smem_bytes;
        o[4] = p.a_slots; o[5] = p.b_slots; o[6] = p.taps_per_cta; o[7] = p.npad; o[8] = p.tmem_cols;
        o[9] = p.TW; o[10] = p.TH; o[11] = p.TN; o[12] = p.nchunks; o[13] = p.chunks_per_split; o[14] = p.na_max;
        o[15] = p.nb_max;
        return 0;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_wgrad_tc: cuTensorMapEncodeTiled unavailable");

    CUtensorMap tmDy, tmX;
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
    cuuint32_t es[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cout_p, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldy * 4, (cuuint64_t)p.Wo * ldy * 4, (cuuint64_t)p.Ho * p.Wo * ldy * 4};
        CUresult rc = encode(&tmDy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(rc == CUDA_SUCCESS, "mk_conv2d_wgrad_tc: dy tensor map rejected (%d)", (int)rc);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_p, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)Win * ldx * 4, (cuuint64_t)Hin * Win * ldx * 4};
        CUresult rc = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(rc == CUDA_SUCCESS, "mk_conv2d_wgrad_tc: x tensor map rejected (%d)", (int)rc);
    }
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long attr_done = 0;
    if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WSMEM_MAX);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done |= attr_bit;
    }
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dwpack, 0, sizeof(float) * (size_t)R * S * Cin_p * Cout_p, st);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_tc memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)co_tiles, (unsigned)(p.n_tap_groups * p.n_ci_tiles), (unsigned)splits);
    k_wgrad_tc<<<grid, 256, smem_bytes, st>>>(tmDy, tmX, p);
    return mk_check_launch("mk_conv2d_wgrad_tc");
}

// 3xTF32 variant: fp32-accurate weight gradient on the tensor cores (same contract).
MK_EXPORT int mk_conv2d_wgrad_tc_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                    int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    t_wx3 = 1;
    const int rc = mk_conv2d_wgrad_tc(x, N, Hin, Win, Cin_p, ldx, dy, Cout_p, ldy, R, S, pad, dwpack, stream);
    t_wx3 = 0;
    return rc;
}

// Dry run of mk_conv2d_wgrad_tc's host-side planning (see mk_conv2d_tc_plan): out[16] = grid.x (co tiles), grid.y
// (tap groups x ci tiles), grid.z (pixel splits), dynamic smem bytes, A ring slots, B ring slots, taps per CTA,
// accumulator columns per tap, TMEM columns, TMA box TW, TH, TN, pixel chunks, chunks per split, A boxes, B boxes.
MK_EXPORT int mk_conv2d_wgrad_tc_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int R, int S, int pad, int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_wgrad_tc_plan: out is NULL");
    t_wx3 = pad >> 8;  // bits 8+ of `pad` select the 3xTF32 plan
    pad &= 255;
    struct Reset { ~Reset() { t_wx3 = 0; } } reset;
    t_wplan_out = out;
    const int rc = mk_conv2d_wgrad_tc(nullptr, N, Hin, Win, Cin_p, Cin_p, nullptr, Cout_p, Cout_p, R, S, pad, nullptr,
                                      nullptr);
    t_wplan_out = nullptr;
    return rc;
}
