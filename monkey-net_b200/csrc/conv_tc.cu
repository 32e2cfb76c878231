// Tensor-core convolution for sm_100a: TMA-tiled implicit GEMM on tcgen05 (kind::tf32) with TMEM accumulators.
//
//   D[128 pixels][BN couts] += A[128 pixels][32 ch] * B[BN couts][32 ch]^T      per (filter tap, 32-channel chunk)
//
// * A tile = ONE 4-D TMA box {32 ch, TW, TH, TN} of the NHWC activation, shifted by the tap offset; out-of-bounds
//   pixels / channels are zero-filled by the TMA unit, which IS the convolution's zero padding (and the K padding
//   of ragged channel counts).  The box lands in shared memory as 128 rows x 128 B with the 128B swizzle - exactly
//   the canonical K-major UMMA operand layout, so no thread ever touches the operands.
// * B tile = 3-D TMA box {32 ch, BN, 1} of the weights packed [tap][Cout_p][Cin_p] (K-major, pre-rounded to TF32).
// * warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane, tcgen05.mma, commit -> frees the
//   smem stage), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b -> registers -> affine, residual,
//   activation -> 128-bit stores).  smem ring of STAGES stages with full/empty mbarriers; accumulator handed over
//   through a tmem_full mbarrier.  The ring takes ~200 KB (6-8 stages): at these sizes the K loop is bound by the
//   TMA round trip, not by the MMA issue rate, so depth beats a second resident CTA.
// * Occupancy instead of one fat CTA per SM: the ring is sized to the K loop actually run (2-8 stages) inside a
//   budget that lets 2-3 CTAs share an SM (TMEM columns are allocated to the tile's N, 32..128), so one CTA's
//   prologue / epilogue overlaps its neighbours' main loops.
// * split-K (grid.z): layers whose tile count cannot fill 148 SMs (deep, low-resolution levels: 128 pixels x 1152 K)
//   split the (tap, channel-chunk) loop over several CTAs that combine with red.global.add.v4.f32 into the
//   zero-initialised output; split 0 carries the bias / residual.  Only for the linear epilogue (act == 0).
// * Reference precision (mk_conv2d_tc_x3): TF32 main term + BF16 cross terms, two MMAs per K step (scheme and error
//   analysis in conv_halo.cu).  The weight pack carries its cross operand behind the hi half (mk_pack_weight mode | 8);
//   the ACTIVATION tile is split in shared memory by the four epilogue warps, idle during the main loop: hi =
//   rna_tf32(v) in place, [bf16(v - hi) x8 | bf16(v) x8] per K step into a second tile of the same swizzled layout,
//   published to the tensor core's async proxy with fence.proxy.async + a per-stage mbarrier.
// Same contract as mk_conv2d (conv.cu) for stride-1 convs without the pool option.
#include "common.cuh"
#include "../../include/monkey_b200.h"
#include "tc_common.cuh"

namespace {
using namespace mk_tc;

constexpr int BM = 128;        // output pixels per CTA (UMMA M)
constexpr int BN_MAX = 128;    // output channels per CTA (UMMA N <= 128)
constexpr int KC = 32;         // fp32 channels per stage = 128 bytes = one swizzle row
constexpr int MAX_STAGES = 8;   // the ring is as deep as ~200 KB of shared memory allows: the K loop of these convs is
                                // TMA-latency bound (~1 us round trip vs ~0.13 us of MMA per stage), not MMA bound
constexpr int A_BYTES = BM * KC * 4;
constexpr int SMEM_MAX = 227 * 1024;

struct TcP {
    int N, Ho, Wo, Cout_p, ldy, Cin_p, R, S, pad;
    int ups;  // 1: nearest-x2-upsampled 3x3 conv as four 2x2 sub-pixel convs (grid.z = output parity)
    int TW, TH, TN, tilesW, tilesH;
    int nstages, stage_bytes;  // smem ring: stage = A tile (16 KB) + B tile (b_rows x 128 B)
    int ksplit, iters_per_split, tmem_cols;
    int x3, b_bytes;  // reference precision: stage = A_hi | B_hi | A_cross | B_cross
    const float* scale; const float* shift; const float* resid; int ldr, act; float slope;
    float* y;
};

__global__ void __launch_bounds__(256) k_conv_tc(const __grid_constant__ CUtensorMap tmA,
                                                 const __grid_constant__ CUtensorMap tmB,
                                                 const __grid_constant__ CUtensorMap tmB2, const TcP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.nstages, STAGE_BYTES = p.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* splitb = empty + MAX_STAGES;
    uint64_t* tmem_full = splitb + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile coordinates
    int t = blockIdx.x;
    const int tw = t % p.tilesW; t /= p.tilesW;
    const int th = t % p.tilesH; t /= p.tilesH;
    const int w0 = tw * p.TW, h0 = th * p.TH, n0 = t * p.TN;
    const int cout0 = blockIdx.y * BN_MAX;
    const int n_this = min(BN_MAX, p.Cout_p - cout0);
    const int nchunks = (p.Cin_p + KC - 1) / KC;
    const int split = (int)blockIdx.z % p.ksplit, zpar = (int)blockIdx.z / p.ksplit;
    const int it0 = split * p.iters_per_split;
    const int it1 = min(p.R * p.S * nchunks, it0 + p.iters_per_split);
    const int niter = it1 - it0;  // >= 1 by construction of ksplit
    // sub-pixel decomposition of conv3x3(upsample2x(x)): output parity (py,px) is a 2x2 conv of x whose taps are
    // sums of the 3x3 taps (pre-summed by mk_pack_weight mode 4) with row offsets {-1,0} (py=0) or {0,+1} (py=1)
    const int py = p.ups ? (zpar >> 1) : 0, px = p.ups ? (zpar & 1) : 0;
    const int pad_h = p.ups ? 1 - py : p.pad, pad_w = p.ups ? 1 - px : p.pad;
    const int tap0 = p.ups ? zpar * 4 : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&splitb[s], 4); }
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
            for (int li = 0; li < niter; ++li) {
                const int stage = li % STAGES;
                const uint32_t phase = (li / STAGES) & 1;
                const int it = it0 + li;
                const int tap = it / nchunks, ch = it - tap * nchunks;
                const int r = tap / p.S, s = tap - r * p.S;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a = smem + stage * STAGE_BYTES;
                mbar_expect_tx(&full[stage], A_BYTES + (p.x3 ? 2 : 1) * p.b_bytes);
                tma_load_4d(a, &tmA, &full[stage], ch * KC, w0 + s - pad_w, h0 + r - pad_h, n0);
                tma_load_3d(a + A_BYTES, &tmB, &full[stage], ch * KC, cout0, tap0 + tap);
                if (p.x3) tma_load_3d(a + 2 * A_BYTES + p.b_bytes, &tmB2, &full[stage], ch * KC, cout0, tap0 + tap);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        const uint32_t idesc = umma_idesc_tf32(BM, (n_this + 15) & ~15);  // rows beyond Cout_p are TMA zero fill
        const uint32_t idesc_c = umma_idesc_bf16(BM, (n_this + 15) & ~15);
        for (int li = 0; li < niter; ++li) {
            const int stage = li % STAGES;
            const uint32_t phase = (li / STAGES) & 1;
            const int ch = (it0 + li) % nchunks;
            mbar_wait(p.x3 ? &splitb[stage] : &full[stage], phase);  // x3: the split tile implies the landed one
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint8_t* a = smem + stage * STAGE_BYTES;
                const uint64_t adesc = umma_desc(a), bdesc = umma_desc(a + A_BYTES);
                int kleft = p.Cin_p - ch * KC;
                if (kleft > KC) kleft = KC;
                const int nk = (kleft + 7) >> 3;  // UMMA K = 8 tf32 (32 bytes); the TMA zero-fills the ragged tail
                if (p.x3) {
                    const uint64_t acdesc = umma_desc(a + A_BYTES + p.b_bytes);
                    const uint64_t bcdesc = umma_desc(a + 2 * A_BYTES + p.b_bytes);
                    for (int k = 0; k < nk; ++k) {   // a_hi * b_hi (TF32), then a_lo * b + a * b_lo (BF16, K = 16)
                        umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (li | k) ? 1u : 0u);
                        umma_bf16(tmem_base, acdesc + 2 * k, bcdesc + 2 * k, idesc_c, 1u);
                    }
                } else
                for (int k = 0; k < nk; ++k)      // advancing 32 B inside the 128 B swizzle row = +2 in 16 B units
                    umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (li | k) ? 1u : 0u);
                umma_commit(&empty[stage]);       // frees the smem stage when these MMAs retire
                if (li == niter - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================================================================== operand split (x3), then epilogue
        if (p.x3) {
            const int tid = threadIdx.x - 128;
            for (int li = 0; li < niter; ++li) {
                const int stage = li % STAGES;
                mbar_wait(&full[stage], (li / STAGES) & 1);
                float4* hi = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES);
                uint4* cr = reinterpret_cast<uint4*>(smem + stage * STAGE_BYTES + A_BYTES + p.b_bytes);
                // one 16-byte chunk (4 channels) per lane and piece; the other 4 channels of the K step sit in the
                // neighbouring chunk (lane ^ 1; the 128B swizzle swaps the pair in odd rows) - see conv_halo.cu
#pragma unroll
                for (int k = 0; k < A_BYTES / 16 / 128; ++k) {
                    const int i = tid + 128 * k;
                    const float4 v = hi[i];
                    float4 h;
                    uint2 lo, top;
                    split_cross(v, h, lo.x, lo.y, top.x, top.y);
                    const bool first = ((i ^ (i >> 3)) & 1) == 0;   // this lane holds channels 0-3 of its K step
                    uint2 send = first ? top : lo, recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    hi[i] = h;
                    cr[i] = first ? make_uint4(lo.x, lo.y, recv.x, recv.y) : make_uint4(recv.x, recv.y, top.x, top.y);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&splitb[stage]);
            }
        }
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;        // GEMM row = pixel of the tile, TMA box order (w fastest, then h, n)
        const int iw = row % p.TW;
        const int ih = (row / p.TW) % p.TH;
        const int in_ = row / (p.TW * p.TH);
        const int n = n0 + in_, h = h0 + ih, w = w0 + iw;
        const bool valid = n < p.N && h < p.Ho && w < p.Wo;
        const long long pix = p.ups ? ((long long)n * (2 * p.Ho) + (2 * h + py)) * (2 * p.Wo) + (2 * w + px)
                                    : ((long long)n * p.Ho + h) * p.Wo + w;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < n_this; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (!valid) continue;
            const int co = cout0 + c;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                if (c + j >= n_this) break;  // ragged Cout_p (multiple of 4, not of 16)
                float4 sc = p.scale ? ldg4(p.scale + co + j) : make_float4(1.f, 1.f, 1.f, 1.f);
                if (p.ksplit > 1) {  // partial sum of a split-K tile: linear epilogue, split 0 adds bias + residual
                    float4 o = make_float4(v[j] * sc.x, v[j + 1] * sc.y, v[j + 2] * sc.z, v[j + 3] * sc.w);
                    if (split == 0) {
                        if (p.shift) o = o + ldg4(p.shift + co + j);
                        if (p.resid) o = o + ldg4(p.resid + pix * p.ldr + co + j);
                    }
                    atomicAdd(reinterpret_cast<float4*>(p.y + pix * p.ldy + co + j), o);
                    continue;
                }
                float4 sh = p.shift ? ldg4(p.shift + co + j) : f4zero();
                float4 o = make_float4(fmaf(v[j], sc.x, sh.x), fmaf(v[j + 1], sc.y, sh.y), fmaf(v[j + 2], sc.z, sh.z),
                                       fmaf(v[j + 3], sc.w, sh.w));
                if (p.resid) o = o + ldg4(p.resid + pix * p.ldr + co + j);
                if (p.act == 1) {
                    o.x = o.x > 0.f ? o.x : o.x * p.slope; o.y = o.y > 0.f ? o.y : o.y * p.slope;
                    o.z = o.z > 0.f ? o.z : o.z * p.slope; o.w = o.w > 0.f ? o.w : o.w * p.slope;
                } else if (p.act == 2) {
                    o.x = 1.f / (1.f + expf(-o.x)); o.y = 1.f / (1.f + expf(-o.y));
                    o.z = 1.f / (1.f + expf(-o.z)); o.w = 1.f / (1.f + expf(-o.w));
                }
                st4(p.y + pix * p.ldy + co + j, o);
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

static thread_local int* t_plan_out = nullptr;  // set by mk_conv2d_tc_plan for a dry run
static thread_local int t_x3 = 0;               // set by mk_conv2d_tc_x3 around its call of mk_conv2d_tc

// Returns 0 on success, -2 if the shape is outside this kernel's envelope (caller uses mk_conv2d instead).
MK_EXPORT int mk_conv2d_tc(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups,
                           const float* wpack_tc, int R, int S, int pad, const float* scale, const float* shift, const float* resid, int ldr, int act,
                           float slope, float* y, int Cout_p, int ldy, void* stream) {
    if (Cin_p % 4 || ldx % 4 || Cout_p % 4 || ldy % 4 || (resid && ldr % 4)) {
        mk_set_error("mk_conv2d_tc: unsupported channel configuration");
        return -2;
    }
    TcP p;
    p.ups = ups ? 1 : 0;
    if (p.ups) {
        MK_REQUIRE(R == 3 && S == 3 && pad == 1, "mk_conv2d_tc: the upsampled path is the 3x3 / pad 1 conv of util.py:71-88");
        R = S = 2;  // per-parity 2x2 kernels; the GEMM tile walks the LOW-resolution grid
        pad = 0;
        p.N = N; p.Ho = Hin; p.Wo = Win;
    } else {
        p.N = N; p.Ho = Hin + 2 * pad - R + 1; p.Wo = Win + 2 * pad - S + 1;
    }
    MK_REQUIRE(p.Ho > 0 && p.Wo > 0, "mk_conv2d_tc: empty output");
    p.Cout_p = Cout_p; p.ldy = ldy; p.Cin_p = Cin_p; p.R = R; p.S = S; p.pad = pad;
    p.TW = pow2_ceil(p.Wo) < 16 ? pow2_ceil(p.Wo) : 16;
    p.TH = pow2_ceil(p.Ho) < BM / p.TW ? pow2_ceil(p.Ho) : BM / p.TW;
    p.TN = BM / (p.TW * p.TH);
    p.tilesW = (p.Wo + p.TW - 1) / p.TW; p.tilesH = (p.Ho + p.TH - 1) / p.TH;
    const int tilesN = (N + p.TN - 1) / p.TN;
    p.scale = scale; p.shift = shift; p.resid = resid; p.ldr = ldr; p.act = act; p.slope = slope; p.y = y;

    const int b_rows = Cout_p < BN_MAX ? (Cout_p + 15) & ~15 : BN_MAX;  // weight rows per stage = UMMA N
    p.x3 = t_x3;
    p.b_bytes = b_rows * KC * 4;
    p.stage_bytes = (p.x3 ? 2 : 1) * (A_BYTES + p.b_bytes);
    p.tmem_cols = b_rows <= 32 ? 32 : (b_rows <= 64 ? 64 : 128);
    const int grid_y = (Cout_p + BN_MAX - 1) / BN_MAX;
    const long long tiles = (long long)p.tilesW * p.tilesH * tilesN * grid_y * (p.ups ? 4 : 1);
    const int niter_total = R * S * ((Cin_p + KC - 1) / KC);
    const int sms = mk_num_sms();
    // split-K only for the linear epilogue into a dense output, when the tiles alone leave most SMs idle
    p.ksplit = 1;
    if (act == 0 && ldy == Cout_p && tiles * 2 <= sms && niter_total >= 4) {
        long long want = (2LL * sms + tiles - 1) / tiles;
        if (want > niter_total / 2) want = niter_total / 2;  // at least 2 K iterations per CTA
        if (want > 64) want = 64;
        if (want > 1) p.ksplit = (int)want;
    }
    p.iters_per_split = (niter_total + p.ksplit - 1) / p.ksplit;
    p.ksplit = (niter_total + p.iters_per_split - 1) / p.iters_per_split;
    // ring depth: never deeper than the K loop; shallow enough for 2-3 resident CTAs when there are CTAs to overlap
    const long long ctas = tiles * p.ksplit;
    const int budget = ctas > 2LL * sms ? (p.stage_bytes <= 24 * 1024 ? 72 * 1024 : 108 * 1024)
                                        : (ctas > sms ? 108 * 1024 : 200 * 1024);
    int nst = budget / p.stage_bytes;
    if (nst > MAX_STAGES) nst = MAX_STAGES;
    if (nst > p.iters_per_split) nst = p.iters_per_split;
    if (nst < 2) nst = 2;
    p.nstages = nst;
    const int smem_bytes = p.nstages * p.stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
    MK_REQUIRE(smem_bytes <= SMEM_MAX, "mk_conv2d_tc: shared memory plan exceeds 227 KB (%d)", smem_bytes);
    if (t_plan_out) {  // dry run (mk_conv2d_tc_plan): report the launch plan, touch no device state
        int* o = t_plan_out;
        o[0] = p.tilesW * p.tilesH * tilesN; o[1] = grid_y; o[2] = (p.ups ? 4 : 1) * p.ksplit; o[3] = smem_bytes;
        o[4] = p.nstages; o[5] = p.ksplit; o[6] = p.iters_per_split; o[7] = niter_total; o[8] = p.tmem_cols;
        o[9] = p.TW; o[10] = p.TH; o[11] = p.TN; o[12] = b_rows; o[13] = p.stage_bytes; o[14] = p.Ho; o[15] = p.Wo;
        return 0;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_tc: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tmA, tmB, tmB2;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_p, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)Win * ldx * 4, (cuuint64_t)Hin * Win * ldx * 4};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc: activation tensor map rejected (%d)", (int)r);
    }
    {
        const int ntap = R * S * (p.ups ? 4 : 1);
        cuuint64_t dims[3] = {(cuuint64_t)Cin_p, (cuuint64_t)Cout_p, (cuuint64_t)ntap};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_p * 4, (cuuint64_t)Cin_p * Cout_p * 4};
        cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)b_rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wpack_tc), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc: weight tensor map rejected (%d)", (int)r);
        tmB2 = tmB;
        if (p.x3) {   // cross operand behind the hi half: [tap][Cout_p][Cin_p rounded up to 8] 4-byte slots
            const cuuint64_t cin8 = (cuuint64_t)((Cin_p + 7) & ~7);
            cuuint64_t dims2[3] = {cin8, (cuuint64_t)Cout_p, (cuuint64_t)ntap};
            cuuint64_t strides2[2] = {cin8 * 4, cin8 * Cout_p * 4};
            r = encode(&tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                       const_cast<float*>(wpack_tc) + (size_t)ntap * Cout_p * Cin_p, dims2, strides2, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc: cross-operand tensor map rejected (%d)", (int)r);
        }
    }
    static unsigned long long attr_done = 0;
    if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done |= attr_bit;
    }
    if (p.ksplit > 1) {
        const size_t out_pix = (size_t)N * p.Ho * p.Wo * (p.ups ? 4 : 1);
        cudaError_t e = cudaMemsetAsync(y, 0, out_pix * ldy * sizeof(float), (cudaStream_t)stream);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_tc memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)(p.tilesW * p.tilesH * tilesN), (unsigned)grid_y, (unsigned)((p.ups ? 4 : 1) * p.ksplit));
    k_conv_tc<<<grid, 256, smem_bytes, (cudaStream_t)stream>>>(tmA, tmB, tmB2, p);
    return mk_check_launch("mk_conv2d_tc");
}

// Reference-precision variant (fp32-accurate tensor-core convolution): same contract, `wpack_tc` packed with mode | 8
// (hi half followed by the cross operand).
MK_EXPORT int mk_conv2d_tc_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups,
                              const float* wpack_tc, int R, int S, int pad, const float* scale, const float* shift,
                              const float* resid, int ldr, int act, float slope, float* y, int Cout_p, int ldy,
                              void* stream) {
    t_x3 = 1;
    const int rc = mk_conv2d_tc(x, N, Hin, Win, Cin_p, ldx, ups, wpack_tc, R, S, pad, scale, shift, resid, ldr, act, slope,
                                y, Cout_p, ldy, stream);
    t_x3 = 0;
    return rc;
}

// Dry run of mk_conv2d_tc's host-side planning (no device state touched, works without a GPU: 148 SMs assumed):
// out[16] = grid.x, grid.y, grid.z, dynamic smem bytes, ring stages, ksplit, K iterations per split, K iterations,
// TMEM columns, TMA box TW, TH, TN, weight rows per stage, stage bytes, Ho, Wo.  tests/test_tc_plans.py sweeps every
// layer shape of the shipped configurations through it.
MK_EXPORT int mk_conv2d_tc_plan(int N, int Hin, int Win, int Cin_p, int ups, int R, int S, int pad, int act, int Cout_p,
                                int ldy, int* out) {
    MK_REQUIRE(out != nullptr, "mk_conv2d_tc_plan: out is NULL");
    t_x3 = ups >> 1;  // bit 1 of `ups` selects the 3xTF32 plan
    ups &= 1;
    struct Reset { ~Reset() { t_x3 = 0; } } reset;
    t_plan_out = out;
    const int rc = mk_conv2d_tc(nullptr, N, Hin, Win, Cin_p, Cin_p, ups, nullptr, R, S, pad, nullptr, nullptr, nullptr, 0,
                                act, 0.f, nullptr, Cout_p, ldy, nullptr);
    t_plan_out = nullptr;
    return rc;
}
