// EXPERIMENTAL (written at the end of round 1; opt-in through MONKEY_B200_CONV_HALO=1, never on the default path, its
// GPU test is skipped unless that variable is set).  STATUS after its only GPU run (the round's last 20 seconds of
// budget): the kernel runs to completion and writes every output, but the values are wrong (relative error 0.8 on all
// six test shapes) - consistent with 6 of the 9 windows (those whose row shift 16*r + s is not a multiple of 8)
// being read with the wrong swizzle phase.  That run set the descriptor base offset to (start >> 7) & 7.  Reasoning
// afterwards: the in-atom K advance every tcgen05 kernel here uses (start address + 32 B per K step, base offset 0)
// only works if the hardware XORs the 16-byte-chunk bits with the row bits of the FINAL ABSOLUTE shared-memory
// address - and then a row-shifted window of a 1024-byte-aligned buffer needs NO base offset either; adding one
// double-counts the phase.  The default is therefore now base offset 0 (MONKEY_B200_HALO_BASEOFF=1 restores the
// first variant); untested - the round's GPU budget was spent.
//
// Halo-window tensor-core convolution for sm_100a.  k_conv_tc (conv_tc.cu) fetches the shifted 128-pixel A tile once
// PER FILTER TAP: ncu on 48->48 3x3 @256x256 shows 1.66 GB crossing L2->SM for a 100 MB input, lts throughput 59 %,
// tensor pipe 14 % - the small-channel full-resolution layers are L2->SMEM bound.  This kernel loads the tile's HALO
// once per 32-channel chunk and reads every tap as a row-shifted WINDOW of the same shared-memory buffer:
//
//   * output tile = 8 rows x TWv columns, TWv = 16 - (S-1) (14 for 3x3, 13 for 4x4); GEMM row m = 16*row + col, the
//     columns col >= TWv of each row are junk rows whose results the epilogue skips (M efficiency 14/16 for 3x3);
//   * halo = ONE 4-D TMA box {32 ch, 16 w, 8 + (R-1) + 1 h, 1 n} at (w0 - pad, h0 - pad): 128-byte pixel rows,
//     16 pixels per image row, 128B-swizzled (out-of-bounds = zero fill = conv padding).  Pixel (row + r, col + s)
//     of the halo is shared-memory row m + 16*r + s, so tap (r, s) is the same buffer read through a UMMA
//     descriptor whose start address is advanced by (16*r + s) * 128 B;  the window then no longer starts on a
//     1024-byte swizzle-atom boundary (see the status note above for the base-offset question);  the extra halo row
//     covers the overrun of the last window (m = 127, r = R-1, s = S-1);
//   * L2->SM bytes per tile and chunk: (8 + R) * 2 KB instead of R*S * 16 KB (6.5x less for 3x3);
//   * two rings as in wgrad_tc.cu: halo ring (per chunk) and weight ring (per chunk x tap), MMA order chunk-major.
// Envelope: stride-1, no upsample, R == S in {3, 4}, Ho >= 8, Wo >= TWv, linear or fused epilogue as k_conv_tc,
// no split-K (it is meant for the many-tile layers).  Same contract and epilogue as mk_conv2d_tc otherwise.
#include "common.cuh"
#include "../../include/monkey_b200.h"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace mk_tc;

constexpr int HK = 32;                 // fp32 channels per chunk = 128 bytes
constexpr int H_MAX_A = 3, H_MAX_B = 10;
constexpr int H_SMEM_MAX = 227 * 1024;

struct HaloP {
    int N, Ho, Wo, Cout_p, ldy, Cin_p, R, S, pad;
    int TWv, tilesW, tilesH;
    int halo_rows, a_slot, b_slot, a_slots, b_slots, tmem_cols;
    int use_base_offset;  // 1: descriptor base offset = (start >> 7) & 7 for the shifted windows; 0: leave it zero
    const float* scale; const float* shift; const float* resid; int ldr, act; float slope;
    float* y;
};

// K-major SWIZZLE_128B operand whose first row is NOT on a 1024-byte atom boundary: base offset = row phase
__device__ __forceinline__ uint64_t umma_desc_window(const void* smem, int use_base_offset) {
    const uint32_t addr = smem_u32(smem);
    uint64_t d = umma_desc(smem);
    if (use_base_offset) d |= (uint64_t)((addr >> 7) & 7u) << 49;
    return d;
}

__global__ void __launch_bounds__(256) k_conv_tc_halo(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const HaloP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + p.a_slots * p.a_slot;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + p.b_slots * p.b_slot);
    uint64_t* a_empty = a_full + H_MAX_A;
    uint64_t* b_full = a_empty + H_MAX_A;
    uint64_t* b_empty = b_full + H_MAX_B;
    uint64_t* tmem_full = b_empty + H_MAX_B;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tw = t % p.tilesW; t /= p.tilesW;
    const int th = t % p.tilesH; t /= p.tilesH;
    const int w0 = tw * p.TWv, h0 = th * 8, n = t;
    const int cout0 = blockIdx.y * 128;
    const int n_this = min(128, p.Cout_p - cout0);
    const int nchunks = (p.Cin_p + HK - 1) / HK;
    const int ntaps = p.R * p.S;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < H_MAX_A; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < H_MAX_B; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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
            for (int ch = 0; ch < nchunks; ++ch) {
                const int as = ch % p.a_slots;
                mbar_wait(&a_empty[as], ((ch / p.a_slots) & 1) ^ 1);
                mbar_expect_tx(&a_full[as], p.halo_rows * 16 * 128);
                tma_load_4d(a_ring + as * p.a_slot, &tmA, &a_full[as], ch * HK, w0 - p.pad, h0 - p.pad, n);
                for (int tap = 0; tap < ntaps; ++tap, ++bi) {
                    const int bs = bi % p.b_slots;
                    mbar_wait(&b_empty[bs], ((bi / p.b_slots) & 1) ^ 1);
                    mbar_expect_tx(&b_full[bs], p.b_slot);
                    tma_load_3d(b_ring + bs * p.b_slot, &tmB, &b_full[bs], ch * HK, cout0, tap);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        const uint32_t idesc = umma_idesc_tf32(128, (n_this + 15) & ~15);
        int bi = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int as = ch % p.a_slots;
            mbar_wait(&a_full[as], (ch / p.a_slots) & 1);
            int kleft = p.Cin_p - ch * HK;
            if (kleft > HK) kleft = HK;
            const int nk = (kleft + 7) >> 3;
            for (int tap = 0; tap < ntaps; ++tap, ++bi) {
                const int bs = bi % p.b_slots;
                mbar_wait(&b_full[bs], (bi / p.b_slots) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const int r = tap / p.S, s = tap - r * p.S;
                    const uint64_t adesc = umma_desc_window(a_ring + as * p.a_slot + (16 * r + s) * 128, p.use_base_offset);
                    const uint64_t bdesc = umma_desc(b_ring + bs * p.b_slot);
                    for (int k = 0; k < nk; ++k)
                        umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (ch | tap | k) ? 1u : 0u);
                    umma_commit(&b_empty[bs]);
                    if (tap == ntaps - 1) {
                        umma_commit(&a_empty[as]);
                        if (ch == nchunks - 1) umma_commit(tmem_full);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue
        const int q = warp & 3;
        const int m = q * 32 + lane;              // GEMM row = 16 * tile row + tile column
        const int col = m & 15, row = m >> 4;
        const int h = h0 + row, w = w0 + col;
        const bool valid = col < p.TWv && h < p.Ho && w < p.Wo;
        const long long pix = ((long long)n * p.Ho + h) * p.Wo + w;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < n_this; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (!valid) continue;
            const int co = cout0 + c;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                if (c + j >= n_this) break;
                float4 sc = p.scale ? ldg4(p.scale + co + j) : make_float4(1.f, 1.f, 1.f, 1.f);
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

// Returns 0 on success, -2 when the shape is outside the halo kernel's envelope (callers fall back to mk_conv2d_tc).
MK_EXPORT int mk_conv2d_tc_halo(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc,
                                int R, int S, int pad, const float* scale, const float* shift, const float* resid,
                                int ldr, int act, float slope, float* y, int Cout_p, int ldy, void* stream) {
    const int Ho = Hin + 2 * pad - R + 1, Wo = Win + 2 * pad - S + 1;
    const int TWv = 16 - (S - 1);
    if (Cin_p % 4 || ldx % 4 || Cout_p % 4 || ldy % 4 || (resid && ldr % 4) || R != S || (R != 3 && R != 4) || Ho < 8 ||
        Wo < TWv) {
        mk_set_error("mk_conv2d_tc_halo: outside the halo kernel's envelope");
        return -2;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_tc_halo: cuTensorMapEncodeTiled unavailable");
    HaloP p;
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout_p = Cout_p; p.ldy = ldy; p.Cin_p = Cin_p; p.R = R; p.S = S; p.pad = pad;
    p.TWv = TWv; p.tilesW = (Wo + TWv - 1) / TWv; p.tilesH = (Ho + 7) / 8;
    p.halo_rows = 8 + (R - 1) + 1;
    p.a_slot = p.halo_rows * 16 * 128;                      // 22 or 24 KB, a multiple of 1024
    const int b_rows = Cout_p < 128 ? (Cout_p + 15) & ~15 : 128;
    p.b_slot = b_rows * 128;
    p.tmem_cols = b_rows <= 32 ? 32 : (b_rows <= 64 ? 64 : 128);
    p.scale = scale; p.shift = shift; p.resid = resid; p.ldr = ldr; p.act = act; p.slope = slope; p.y = y;
    {
        const char* e = getenv("MONKEY_B200_HALO_BASEOFF");
        p.use_base_offset = (e && e[0] == '1') ? 1 : 0;
    }
    const int nchunks = (Cin_p + HK - 1) / HK;
    p.a_slots = nchunks < 2 ? 1 : 2;
    int budget = 100 * 1024 - p.a_slots * p.a_slot;         // two CTAs per SM
    int bs = budget / p.b_slot;
    if (bs > H_MAX_B) bs = H_MAX_B;
    if (bs > nchunks * R * S) bs = nchunks * R * S;
    if (bs < 2) bs = 2;
    p.b_slots = bs;
    const int smem_bytes = p.a_slots * p.a_slot + p.b_slots * p.b_slot + 512 /*barriers*/ + 1024 /*align*/;
    MK_REQUIRE(smem_bytes <= H_SMEM_MAX, "mk_conv2d_tc_halo: shared memory plan exceeds 227 KB (%d)", smem_bytes);

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin_p, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)Win * ldx * 4, (cuuint64_t)Hin * Win * ldx * 4};
        cuuint32_t box[4] = {(cuuint32_t)HK, 16, (cuuint32_t)p.halo_rows, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: activation tensor map rejected (%d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_p, (cuuint64_t)Cout_p, (cuuint64_t)(R * S)};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_p * 4, (cuuint64_t)Cin_p * Cout_p * 4};
        cuuint32_t box[3] = {(cuuint32_t)HK, (cuuint32_t)b_rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wpack_tc), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MK_REQUIRE(r == CUDA_SUCCESS, "mk_conv2d_tc_halo: weight tensor map rejected (%d)", (int)r);
    }
    static unsigned long long attr_done = 0;
    if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tc_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_MAX);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_tc_halo: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done |= attr_bit;
    }
    dim3 grid((unsigned)(p.tilesW * p.tilesH * N), (unsigned)((Cout_p + 127) / 128), 1);
    k_conv_tc_halo<<<grid, 256, smem_bytes, (cudaStream_t)stream>>>(tmA, tmB, p);
    return mk_check_launch("mk_conv2d_tc_halo");
}
