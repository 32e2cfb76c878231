// Tensor-core weight gradient for sm_100a:   dW[tap][ci][co] = sum_pixels dY[pix][co] * X[pix + tap][ci]
//
// GEMM view: D[M = co][N = ci] accumulated over K = pixels.  In NHWC both operands have their M/N index (channels)
// contiguous and the contraction index (pixels) strided, i.e. they are "MN-major" UMMA operands - tcgen05 takes them
// directly (a_major = b_major = 1 in the instruction descriptor), so no transposed copies of the activations exist:
//   * A stage = four 4-D TMA boxes {32 co, TW, TH, TN} of dY (32 pixels each), B stage = up to four boxes
//     {32 ci, TW, TH, TN} of X shifted by the filter tap; out-of-bounds pixels are zero-filled = conv padding.
//     Each box is [32 pixels][128 B] in the 128B swizzle with 32-byte atoms (TMA SWIZZLE_128B_ATOM_32B ==
//     UMMA SWIZZLE_128B_BASE32B, the only MN-major layout 32-bit operands may use) = one MN block of the
//     canonical layout ((8,n),(4,k)) with LBO = box size (4096 B), SBO = 512 B (4 pixel rows).
//   * 4 x tcgen05.mma.kind::tf32 (K = 8 pixels each) per stage; accumulator [128 co][<=128 ci] fp32 in TMEM.
//   * grid = (co tiles, taps x ci tiles, pixel splits); splits combine with fp32 atomics (red) into the packed
//     gradient, whose layout [tap][Cin_p][Cout_p] makes the epilogue's per-column writes coalesced across lanes.
#include "tc_common.cuh"
#include "../../include/monkey_b200.h"

namespace {
using namespace mk_tc;

constexpr int PC = 32;            // pixels per stage
constexpr int BOX_BYTES = PC * 128;
constexpr int WSTAGES = 6;
constexpr int WSTAGE_BYTES = 8 * BOX_BYTES;  // 4 A boxes + 4 B boxes
constexpr int WSMEM_BYTES = WSTAGES * WSTAGE_BYTES + 1024 + 256;

struct WgTcP {
    int N, Ho, Wo, Cout_p, Cin_p, R, S, pad;
    int TW, TH, TN, tilesW, tilesH, nchunks, chunks_per_split, n_ci_tiles;
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
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WSTAGES * WSTAGE_BYTES);
    uint64_t* empty = full + WSTAGES;
    uint64_t* tmem_full = empty + WSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * 128;
    const int tap = blockIdx.y / p.n_ci_tiles, ci0 = (blockIdx.y % p.n_ci_tiles) * 128;
    const int r = tap / p.S, s = tap - r * p.S;
    const int n_this = min(128, p.Cin_p - ci0);          // multiple of 4; the MMA runs on the next multiple of 16
    const int nb = (n_this + 31) >> 5;                   // B boxes actually needed
    const int q0 = blockIdx.z * p.chunks_per_split;
    const int q1 = min(p.nchunks, q0 + p.chunks_per_split);
    const int niter = q1 - q0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < WSTAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int it = 0; it < niter; ++it) {
                const int stage = it % WSTAGES;
                const uint32_t phase = (it / WSTAGES) & 1;
                int q = q0 + it;
                const int tw = q % p.tilesW; q /= p.tilesW;
                const int th = q % p.tilesH; q /= p.tilesH;
                const int w0 = tw * p.TW, h0 = th * p.TH, n0 = q * p.TN;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a = smem + stage * WSTAGE_BYTES;
                mbar_expect_tx(&full[stage], (4 + nb) * BOX_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_4d(a + j * BOX_BYTES, &tmDy, &full[stage], co0 + j * 32, w0, h0, n0);
                for (int j = 0; j < nb; ++j)
                    tma_load_4d(a + (4 + j) * BOX_BYTES, &tmX, &full[stage], ci0 + j * 32, w0 + s - p.pad, h0 + r - p.pad,
                                n0);
            }
        }
    } else if (warp == 1) {
        // M = 128 (co), N = n_this (ci), both operands MN-major
        const uint32_t idesc = umma_idesc_tf32(128, (n_this + 15) & ~15) | (1u << 15) | (1u << 16);
        for (int it = 0; it < niter; ++it) {
            const int stage = it % WSTAGES;
            const uint32_t phase = (it / WSTAGES) & 1;
            mbar_wait(&full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint8_t* a = smem + stage * WSTAGE_BYTES;
                const uint64_t adesc = umma_desc_mn(a), bdesc = umma_desc_mn(a + 4 * BOX_BYTES);
#pragma unroll
                for (int k = 0; k < PC / 8; ++k)  // 8 pixel rows = 1024 B = 64 sixteen-byte units
                    umma_tf32(tmem_base, adesc + 64 * k, bdesc + 64 * k, idesc, (it | k) ? 1u : 0u);
                umma_commit(&empty[stage]);
                if (it == niter - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && niter > 0) {
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        const bool valid = co < p.Cout_p;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* base = p.dw + ((long long)tap * p.Cin_p + ci0) * p.Cout_p + co;
        for (int c = 0; c < n_this; c += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace

MK_EXPORT int mk_conv2d_wgrad_tc(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy,
                                 int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream) {
    if (Cin_p % 4 || Cout_p % 4 || ldx % 4 || ldy % 4) {
        mk_set_error("mk_conv2d_wgrad_tc: unsupported channel configuration");
        return -2;
    }
    EncodeTiledFn encode = get_encode();
    MK_REQUIRE(encode != nullptr, "mk_conv2d_wgrad_tc: cuTensorMapEncodeTiled unavailable");
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
    const long long tiles = (long long)co_tiles * p.n_ci_tiles * R * S;
    long long splits = mk_cdiv(2LL * mk_num_sms(), tiles);
    if (splits > p.nchunks / 4) splits = p.nchunks / 4;  // at least 4 chunks (128 pixels) per CTA
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    p.chunks_per_split = (int)mk_cdiv(p.nchunks, splits);
    splits = mk_cdiv(p.nchunks, p.chunks_per_split);

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
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WSMEM_BYTES);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dwpack, 0, sizeof(float) * (size_t)R * S * Cin_p * Cout_p, st);
        if (e != cudaSuccess) { mk_set_error("mk_conv2d_wgrad_tc memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)co_tiles, (unsigned)(R * S * p.n_ci_tiles), (unsigned)splits);
    k_wgrad_tc<<<grid, 256, WSMEM_BYTES, st>>>(tmDy, tmX, p);
    return mk_check_launch("mk_conv2d_wgrad_tc");
}
