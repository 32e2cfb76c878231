// Deformation sampling: F.grid_sample (5-D call with D_in = 1, z == 0  ==  2-D bilinear, zeros padding,
// align_corners=True under torch 0.4.1) with the deformation-grid resize of generator.py:51-58 fused in, plus the
// plain NHWC resize used for the keypoint-embedding skips (generator.py:72).
//
// HBM-bound.  NHWC makes every bilinear tap a contiguous channel vector: one thread owns one float4 of channels of
// one output pixel, so the 4 taps are 128-bit loads that coalesce across the lanes of a pixel, and the output is a
// coalesced 128-bit store.  Algorithmic bytes per call: 4*(B*C*h*w + 2*B*d*h*w + B*d*C*h*w).
#include "common.cuh"
#include <stdlib.h>
#include "../../include/monkey_b200.h"

struct Tap {
    int x0, y0;
    float wx1, wy1;  // weight of the x0+1 / y0+1 taps
    float gx, gy;    // normalised coordinates actually sampled
};

// normalised sampling coordinate of output pixel (ho,wo) of frame n from the coarse deformation field
__device__ __forceinline__ float2 fetch_grid(const float* __restrict__ deform, long long n, int h0, int w0, int ho,
                                             int wo, int h, int w, int mode) {
    const float* base = deform + n * (long long)h0 * w0 * 2;
    if (h == h0 && w == w0) {
        return __ldg(reinterpret_cast<const float2*>(base + ((long long)ho * w0 + wo) * 2));
    }
    if (mode == 0) {
        int ys = nearest_src(ho, h0, h), xs = nearest_src(wo, w0, w);
        return __ldg(reinterpret_cast<const float2*>(base + ((long long)ys * w0 + xs) * 2));
    }
    int y0, y1, x0, x1;
    float ly, lx;
    linear_src(ho, h0, h, y0, y1, ly);
    linear_src(wo, w0, w, x0, x1, lx);
    float2 a = __ldg(reinterpret_cast<const float2*>(base + ((long long)y0 * w0 + x0) * 2));
    float2 b = __ldg(reinterpret_cast<const float2*>(base + ((long long)y0 * w0 + x1) * 2));
    float2 c = __ldg(reinterpret_cast<const float2*>(base + ((long long)y1 * w0 + x0) * 2));
    float2 d = __ldg(reinterpret_cast<const float2*>(base + ((long long)y1 * w0 + x1) * 2));
    float hy = 1.f - ly, hx = 1.f - lx;
    // same association order as ATen upsample_bilinear2d: h0lambda*(w0lambda*a + w1lambda*b) + h1lambda*(...)
    float2 r;
    r.x = hy * (hx * a.x + lx * b.x) + ly * (hx * c.x + lx * d.x);
    r.y = hy * (hx * a.y + lx * b.y) + ly * (hx * c.y + lx * d.y);
    return r;
}

__device__ __forceinline__ Tap make_tap(float2 g, int h, int w) {
    Tap t;
    t.gx = g.x; t.gy = g.y;
    float ix = ((g.x + 1.f) / 2.f) * (float)(w - 1);
    float iy = ((g.y + 1.f) / 2.f) * (float)(h - 1);
    float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx; t.y0 = (int)fy;
    t.wx1 = ix - fx; t.wy1 = iy - fy;
    return t;
}

// Work decomposition (forward and backward), warp-cooperative:
//   * a CTA (8 warps) owns a PW x PH patch of output pixels of one frame, a warp P of them (P = 32 for the big levels;
//     halved until the launch has ~48 warps per SM, so the small levels keep their memory-level parallelism);
//   * phase 1: lane l < P fetches the deformation of pixel l (32 consecutive grid entries = coalesced 256 B) and computes
//     its bilinear tap ONCE - the previous one-thread-per-(pixel, float4) mapping recomputed the tap (two float
//     divisions of the resize rule, floor, bounds) in every channel lane: 157 warp instructions per 512 B moved, the
//     kernel was issue-bound (ncu: 56 % issue-active at 18 % DRAM);
//   * phase 2: the warp walks its 32 pixels; the tap of pixel q is broadcast with 4 shuffles and every lane moves
//     float4s of channels: lanes-per-pixel = min(32, pow2ceil(cv)), so 32/lpp pixels are in flight per pass and the
//     four tap loads of a pass are independent of every other pass (deep memory-level parallelism, no dependent
//     grid -> tap chain inside the loop).  Per tap a pixel's lanes read one contiguous run of the NHWC pixel.
//   * the 2-D patch keeps the overlap of neighbouring pixels' taps inside one SM's L1.
struct Patch {
    int tiles_x, tiles_y;
    int P, pw, pw_sh, ph;   // pixels per warp; patch = pw x ph pixels (pw = 1 << pw_sh), 8 warps
    int lpp, lpp_sh;        // lanes per pixel (power of two <= 32) and its log2
    int cgroups;            // channel-vector groups of lpp lanes per pixel: ceil(cv / lpp)
};

static inline Patch make_patch(int h, int w, int cv, long long frames, int vec_per_lane = 1, int resident_warps = 40) {
    Patch t;
    t.lpp = 1; t.lpp_sh = 0;
    const int lanes_needed = (cv + vec_per_lane - 1) / vec_per_lane;
    while (t.lpp < lanes_needed && t.lpp < 32) { t.lpp <<= 1; ++t.lpp_sh; }
    t.cgroups = (cv + t.lpp - 1) / t.lpp;
    // pixels per warp: as many as possible (amortises phase 1) while the launch is either a single resident wave or
    // many waves - 1.4 waves of equal CTAs would leave the second wave 60 % empty (ncu: 42 % warps active)
    const long long resident = (long long)resident_warps * mk_num_sms();
    t.P = 32;
    while (t.P > 1) {
        const double waves = (double)(frames * h * w / t.P) / (double)resident;
        if (waves >= 2.6 && frames * h * w / t.P >= 48LL * mk_num_sms()) break;
        t.P >>= 1;
    }
    t.pw = t.P >= 16 ? 16 : t.P;
    t.pw_sh = 0;
    while ((1 << t.pw_sh) < t.pw) ++t.pw_sh;
    t.ph = 8 * (t.P / t.pw);
    t.tiles_x = (w + t.pw - 1) / t.pw; t.tiles_y = (h + t.ph - 1) / t.ph;
    return t;
}

struct WarpTap {
    int base;      // y0 * w + x0 (pixel offset of the north-west tap inside the source frame)
    int opix;      // ho * w + wo of the output pixel, -1 outside the frame (partial patches)
    int bounds;    // bit0 x0 in range, bit1 x0+1, bit2 y0, bit3 y0+1
    float wx1, wy1;
};

__device__ __forceinline__ WarpTap lane_tap(const float* __restrict__ deform, unsigned n, int h0, int w0, int mode,
                                            int h, int w, unsigned tx, unsigned ty, const Patch& pt, int& ho, int& wo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = warp * pt.P + lane;  // pixel of the patch (lanes >= P idle in phase 1)
    ho = (int)ty * pt.ph + (p >> pt.pw_sh); wo = (int)tx * pt.pw + (p & (pt.pw - 1));
    WarpTap t;
    t.opix = -1; t.base = 0; t.bounds = 0; t.wx1 = 0.f; t.wy1 = 0.f;
    if (lane < pt.P && ho < h && wo < w) {
        const Tap tp = make_tap(fetch_grid(deform, n, h0, w0, ho, wo, h, w, mode), h, w);
        t.opix = ho * w + wo;
        t.base = tp.y0 * w + tp.x0;
        t.bounds = (tp.x0 >= 0 && tp.x0 < w ? 1 : 0) | (tp.x0 + 1 >= 0 && tp.x0 + 1 < w ? 2 : 0) |
                   (tp.y0 >= 0 && tp.y0 < h ? 4 : 0) | (tp.y0 + 1 >= 0 && tp.y0 + 1 < h ? 8 : 0);
        t.wx1 = tp.wx1; t.wy1 = tp.wy1;
    }
    return t;
}

__device__ __forceinline__ WarpTap bcast_tap(const WarpTap& t, int src) {
    WarpTap r;
    r.base = __shfl_sync(0xffffffffu, t.base, src);
    r.opix = __shfl_sync(0xffffffffu, t.opix, src);
    r.bounds = __shfl_sync(0xffffffffu, t.bounds, src);
    r.wx1 = __shfl_sync(0xffffffffu, t.wx1, src);
    r.wy1 = __shfl_sync(0xffffffffu, t.wy1, src);
    return r;
}

// Forward tap as broadcast by the owning lane: every quantity phase 2 needs is precomputed, so a pass is
// 6 shuffles + 4 unconditional 128-bit loads + 16 FMAs + 1 store.  Out-of-range taps are expressed as ZERO WEIGHTS
// on coordinates clamped into the frame (no predicates, no divergent loads); all element offsets are 32-bit
// (the host checks h*w*ld < 2^31).
// 16-byte asynchronous global -> shared copy (LDGSTS): the tap loads of several passes are in flight without holding
// destination registers and without any scoreboard dependency on the blend of the previous pass.  (Left to itself
// ptxas issued tap 0, consumed it, recycled its registers for the addresses of taps 2-3 and so serialised two DRAM
// round trips per pass: ncu long-scoreboard 9.4 stall cycles per issue at 33 % DRAM.)
// CA = cache the line in L1 as well (cp.async.ca): horizontally adjacent output pixels share two of their four taps, and
// with .cg every tap went to the L2 - ncu (profiles/r1_ncu_grid_sample_fwd_v3.md) counts 272.6 MB crossing L2->SM for
// a 67 MB input at 0 % L1 hit rate, i.e. ~9.7 TB/s of the ~12 TB/s the crossbar delivers: the kernel was bound by
// L2->SM bandwidth, not by HBM.
template <bool CA>
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    if (CA)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                     "l"(gsrc)
                     : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                     "l"(gsrc)
                     : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared memory: 8 warps x GS_STAGES passes x 4 taps x V vectors x 32 lanes x 16 B (V=1,S=3: 48 KB; V=2,S=2: 64 KB)

struct FwdTap {
    int o00;       // element offset of the clamped north-west tap inside the source frame: (y0c*w + x0c)*ld
    int out;       // (opix * ldo) << 2 | (y1c != y0c) << 1 | (x1c != x0c); -1 when the pixel is outside the frame
    float w00, w01, w10, w11;
};

// phase 1: lane l < P owns pixel l of the warp's run
__device__ __forceinline__ FwdTap fwd_lane_tap(const float* __restrict__ deform, unsigned n, int h0, int w0, int mode,
                                               int h, int w, unsigned tx, unsigned ty, const Patch& pt, int ld, int ldo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FwdTap mine;
    mine.out = -1; mine.o00 = 0; mine.w00 = mine.w01 = mine.w10 = mine.w11 = 0.f;
    const int p = warp * pt.P + lane;
    const int ho = (int)ty * pt.ph + (p >> pt.pw_sh), wo = (int)tx * pt.pw + (p & (pt.pw - 1));
    if (lane < pt.P && ho < h && wo < w) {
        const Tap tp = make_tap(fetch_grid(deform, n, h0, w0, ho, wo, h, w, mode), h, w);
        const int x1 = tp.x0 + 1, y1 = tp.y0 + 1;
        const bool vx0 = tp.x0 >= 0 && tp.x0 < w, vx1 = x1 >= 0 && x1 < w;
        const bool vy0 = tp.y0 >= 0 && tp.y0 < h, vy1 = y1 >= 0 && y1 < h;
        const float wx0 = 1.f - tp.wx1, wy0 = 1.f - tp.wy1;
        mine.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.f;
        mine.w01 = (vx1 && vy0) ? tp.wx1 * wy0 : 0.f;
        mine.w10 = (vx0 && vy1) ? wx0 * tp.wy1 : 0.f;
        mine.w11 = (vx1 && vy1) ? tp.wx1 * tp.wy1 : 0.f;
        const int x0c = min(max(tp.x0, 0), w - 1), x1c = min(max(x1, 0), w - 1);
        const int y0c = min(max(tp.y0, 0), h - 1), y1c = min(max(y1, 0), h - 1);
        mine.o00 = (y0c * w + x0c) * ld;
        mine.out = (((ho * w + wo) * ldo) << 2) | ((y1c != y0c) ? 2 : 0) | ((x1c != x0c) ? 1 : 0);
    }
    return mine;
}

// generic path (cv > 32: the >= 132-channel levels, all small): direct loads, channel groups of 32 lanes
__global__ void __launch_bounds__(256) k_grid_sample_fwd(const float* __restrict__ inp, int h, int w, int cv, int ld,
                                                         const float* __restrict__ deform, int d, int h0, int w0,
                                                         int mode, float* __restrict__ out, int ldo, const Patch pt,
                                                         const FastDiv ftx, const FastDiv fty) {
    unsigned tx, ty;
    const unsigned t1 = fd_divmod(blockIdx.x, ftx, tx);
    const unsigned n = fd_divmod(t1, fty, ty);
    const float* src = inp + (long long)(n / (unsigned)d) * h * w * ld;
    float* dst = out + (long long)n * h * w * ldo;
    const int lane = threadIdx.x & 31;
    const FwdTap mine = fwd_lane_tap(deform, n, h0, w0, mode, h, w, tx, ty, pt, ld, ldo);
    if (__ballot_sync(0xffffffffu, mine.out >= 0) == 0) return;  // warp entirely outside the frame
    // ---- phase 2
    const int sub = lane >> pt.lpp_sh, cl = lane & (pt.lpp - 1);  // pixel slot of this pass, channel lane
    const int ppp = 32 >> pt.lpp_sh;                              // pixels per pass
    const int row = w * ld, c4 = cl * 4;
    const bool lane_on = cl < cv;
#pragma unroll 2
    for (int q0 = 0; q0 < pt.P; q0 += ppp) {
        const int from = (q0 + sub) & 31;
        const int o00 = __shfl_sync(0xffffffffu, mine.o00, from);
        const int po = __shfl_sync(0xffffffffu, mine.out, from);
        const float w00 = __shfl_sync(0xffffffffu, mine.w00, from), w01 = __shfl_sync(0xffffffffu, mine.w01, from);
        const float w10 = __shfl_sync(0xffffffffu, mine.w10, from), w11 = __shfl_sync(0xffffffffu, mine.w11, from);
        if (po < 0 || !lane_on) continue;
        // byte addresses: one 64-bit base per pass, the other taps are 32-bit byte deltas (0 when clamped)
        const char* p0 = reinterpret_cast<const char*>(src) + (long long)(o00 + c4) * 4;
        const int dxb = (po & 1) ? ld * 4 : 0, dyb = (po & 2) ? row * 4 : 0;
        char* op = reinterpret_cast<char*>(dst) + (long long)((po >> 2) + c4) * 4;
        for (int cq = cl; cq < cv; cq += 32, p0 += 512, op += 512) {
            const float4 v0 = ldg4(reinterpret_cast<const float*>(p0));
            const float4 v1 = ldg4(reinterpret_cast<const float*>(p0 + dxb));
            const float4 v2 = ldg4(reinterpret_cast<const float*>(p0 + dyb));
            const float4 v3 = ldg4(reinterpret_cast<const float*>(p0 + dyb + dxb));
            float4 acc = v0 * w00;  // ATen order: nw, ne, sw, se
            fma4(acc, v1, w01);
            fma4(acc, v2, w10);
            fma4(acc, v3, w11);
            __stcs(reinterpret_cast<float4*>(op), acc);  // streamed: never re-read by this kernel
        }
    }
}

// cv <= 32*V: V float4 of channels per lane (vectors cl, cl + lpp, ...), tap loads staged through shared memory with
// cp.async, GS_STAGES passes in flight per warp.  V = 1 serves every level up to 128 channels (measured: V = 2 on
// the 64-channel level was not faster, 32.8 vs 31.1 us, its larger staging buffer costs a resident CTA); V = 2 covers
// the 132..256-channel levels.
template <int V, int GS_STAGES, bool CA>
__global__ void __launch_bounds__(256) k_grid_sample_fwd_async(const float* __restrict__ inp, int h, int w, int cv,
                                                               int ld, const float* __restrict__ deform, int d, int h0,
                                                               int w0, int mode, float* __restrict__ out, int ldo,
                                                               const Patch pt, const FastDiv ftx, const FastDiv fty) {
    extern __shared__ float4 gs_stage[];  // [8 warps][GS_STAGES][4 taps * V][32 lanes]
    unsigned tx, ty;
    const unsigned t1 = fd_divmod(blockIdx.x, ftx, tx);
    const unsigned n = fd_divmod(t1, fty, ty);
    const float* src = inp + (long long)(n / (unsigned)d) * h * w * ld;
    float* dst = out + (long long)n * h * w * ldo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FwdTap mine = fwd_lane_tap(deform, n, h0, w0, mode, h, w, tx, ty, pt, ld, ldo);
    if (__ballot_sync(0xffffffffu, mine.out >= 0) == 0) return;  // warp entirely outside the frame
    const int sub = lane >> pt.lpp_sh, cl = lane & (pt.lpp - 1);
    const int ppp = 32 >> pt.lpp_sh;
    const int rowb = w * ld * 4, ldb = ld * 4, c4 = cl * 4;
    const int vstep = pt.lpp * 16;  // bytes between this lane's channel vectors
    int nv = 0;                     // channel vectors this lane really owns (0..V)
#pragma unroll
    for (int v = 0; v < V; ++v) nv += (cl + v * pt.lpp < cv) ? 1 : 0;
    const int npass = (pt.P + ppp - 1) / ppp;
    float4* buf = gs_stage + (warp * GS_STAGES) * (4 * V * 32) + lane;

    auto issue = [&](int pass) {
        if (pass < npass) {
            const int from = (pass * ppp + sub) & 31;
            const int o00 = __shfl_sync(0xffffffffu, mine.o00, from);
            const int po = __shfl_sync(0xffffffffu, mine.out, from);
            if (po >= 0 && nv > 0) {
                const char* p0 = reinterpret_cast<const char*>(src) + (long long)(o00 + c4) * 4;
                const int dxb = (po & 1) ? ldb : 0, dyb = (po & 2) ? rowb : 0;
                float4* b = buf + (pass % GS_STAGES) * (4 * V * 32);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if (v < nv) {
                        cp_async16<CA>(b + (4 * v + 0) * 32, p0);
                        cp_async16<CA>(b + (4 * v + 1) * 32, p0 + dxb);
                        cp_async16<CA>(b + (4 * v + 2) * 32, p0 + dyb);
                        cp_async16<CA>(b + (4 * v + 3) * 32, p0 + dyb + dxb);
                    }
                    p0 += vstep;
                }
            }
        }
        cp_async_commit();  // every lane commits every time: group counts stay uniform
    };

#pragma unroll
    for (int s = 0; s < GS_STAGES - 1; ++s) issue(s);
    for (int pass = 0; pass < npass; ++pass) {
        issue(pass + GS_STAGES - 1);
        cp_async_wait<GS_STAGES - 1>();  // the group of `pass` has landed (this lane reads only its own slots)
        const int from = (pass * ppp + sub) & 31;
        const int po = __shfl_sync(0xffffffffu, mine.out, from);
        const float w00 = __shfl_sync(0xffffffffu, mine.w00, from), w01 = __shfl_sync(0xffffffffu, mine.w01, from);
        const float w10 = __shfl_sync(0xffffffffu, mine.w10, from), w11 = __shfl_sync(0xffffffffu, mine.w11, from);
        if (po < 0 || nv == 0) continue;
        const float4* b = buf + (pass % GS_STAGES) * (4 * V * 32);
        char* op = reinterpret_cast<char*>(dst) + (long long)((po >> 2) + c4) * 4;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (v < nv) {
                float4 acc = b[(4 * v + 0) * 32] * w00;  // ATen order: nw, ne, sw, se
                fma4(acc, b[(4 * v + 1) * 32], w01);
                fma4(acc, b[(4 * v + 2) * 32], w10);
                fma4(acc, b[(4 * v + 3) * 32], w11);
                __stcs(reinterpret_cast<float4*>(op), acc);
            }
            op += vstep;
        }
    }
}

template <int V, int GS_STAGES, bool CA>
static int launch_gs_async(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d, int h0,
                           int w0, int mode, float* out, int ldo, cudaStream_t st) {
    const int smem = 8 * GS_STAGES * 4 * V * 32 * (int)sizeof(float4);
    static unsigned long long attr_done = 0;
    if (const unsigned long long attr_bit = mk_attr_needed(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(k_grid_sample_fwd_async<V, GS_STAGES, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { mk_set_error("mk_grid_sample_fwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done |= attr_bit;
    }
    int resident = (227 * 1024 / (smem + 1024)) * 8;  // warps per SM by shared memory
    if (resident > 40) resident = 40;                 // ... and by registers
    const Patch pt = make_patch(h, w, Cp / 4, (long long)B * d, V, resident);
    const long long blocks = (long long)B * d * pt.tiles_x * pt.tiles_y;
    MK_REQUIRE(blocks < (1LL << 31), "mk_grid_sample_fwd: extent too large");
    k_grid_sample_fwd_async<V, GS_STAGES, CA><<<(unsigned)blocks, 256, smem, st>>>(inp, h, w, Cp / 4, ld, deform, d, h0, w0, mode, out,
                                                                    ldo, pt, make_fastdiv(pt.tiles_x),
                                                                    make_fastdiv(pt.tiles_y));
    return mk_check_launch("mk_grid_sample_fwd");
}

MK_EXPORT int mk_grid_sample_fwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d,
                                 int h0, int w0, int mode, float* out, int ldo, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0, "mk_grid_sample_fwd: channels must be x4");
    const long long total = (long long)B * d * h * w * (Cp / 4);
    if (total == 0) return 0;
    MK_REQUIRE((long long)h * w * ld < (1LL << 31) && (long long)h * w * ldo < (1LL << 29),
               "mk_grid_sample_fwd: extent too large");
    const int cv = Cp / 4;
    // L1 allocation of the taps (cp.async.ca), measured on a B200 (L2 flushed, CUDA events): 3 ch x 256^2: 24.6 vs 26.6 us
    // (better), 64 ch x 128^2: 36.0 vs 32.8 us, 128 ch x 64^2: 24.6 vs 21.7 us (worse: with 192 KB of staging buffers
    // per SM the L1 is ~30 KB and thrashes) - so only the image-sized levels use it.  MONKEY_B200_GS_CA=0/1 forces.
    static int force_ca = -2;
    if (force_ca == -2) {
        const char* e = getenv("MONKEY_B200_GS_CA");
        force_ca = e ? (e[0] == '0' ? 0 : 1) : -1;
    }
    const int use_ca = force_ca >= 0 ? force_ca : (cv <= 2 ? 1 : 0);
    if (cv <= 32)
        return use_ca ? launch_gs_async<1, 3, true>(inp, B, h, w, Cp, ld, deform, d, h0, w0, mode, out, ldo, (cudaStream_t)stream)
                      : launch_gs_async<1, 3, false>(inp, B, h, w, Cp, ld, deform, d, h0, w0, mode, out, ldo, (cudaStream_t)stream);
    if (cv <= 64)
        return use_ca ? launch_gs_async<2, 2, true>(inp, B, h, w, Cp, ld, deform, d, h0, w0, mode, out, ldo, (cudaStream_t)stream)
                      : launch_gs_async<2, 2, false>(inp, B, h, w, Cp, ld, deform, d, h0, w0, mode, out, ldo, (cudaStream_t)stream);
    const Patch pt = make_patch(h, w, cv, (long long)B * d);
    const long long blocks = (long long)B * d * pt.tiles_x * pt.tiles_y;
    MK_REQUIRE(blocks < (1LL << 31), "mk_grid_sample_fwd: extent too large");
    k_grid_sample_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        inp, h, w, cv, ld, deform, d, h0, w0, mode, out, ldo, pt, make_fastdiv(pt.tiles_x), make_fastdiv(pt.tiles_y));
    return mk_check_launch("mk_grid_sample_fwd");
}

// backward: dinp via vector atomics (red.global.add.v4.f32), d(grid) reduced over the channel lanes of a pixel
// (segmented warp shuffle over the lpp lanes), then chained through the resize into the coarse deformation gradient
// with scalar atomics.
__device__ __forceinline__ void scatter_grid_grad(float* __restrict__ ddeform, long long n, int h0, int w0, int ho,
                                                  int wo, int h, int w, int mode, float gx, float gy) {
    float* base = ddeform + n * (long long)h0 * w0 * 2;
    if (h == h0 && w == w0) {
        float* q = base + ((long long)ho * w0 + wo) * 2;
        atomicAdd(q, gx); atomicAdd(q + 1, gy);
        return;
    }
    if (mode == 0) {
        int ys = nearest_src(ho, h0, h), xs = nearest_src(wo, w0, w);
        float* q = base + ((long long)ys * w0 + xs) * 2;
        atomicAdd(q, gx); atomicAdd(q + 1, gy);
        return;
    }
    int y0, y1, x0, x1;
    float ly, lx;
    linear_src(ho, h0, h, y0, y1, ly);
    linear_src(wo, w0, w, x0, x1, lx);
    float hy = 1.f - ly, hx = 1.f - lx;
    float* q;
    q = base + ((long long)y0 * w0 + x0) * 2; atomicAdd(q, hy * hx * gx); atomicAdd(q + 1, hy * hx * gy);
    q = base + ((long long)y0 * w0 + x1) * 2; atomicAdd(q, hy * lx * gx); atomicAdd(q + 1, hy * lx * gy);
    q = base + ((long long)y1 * w0 + x0) * 2; atomicAdd(q, ly * hx * gx); atomicAdd(q + 1, ly * hx * gy);
    q = base + ((long long)y1 * w0 + x1) * 2; atomicAdd(q, ly * lx * gx); atomicAdd(q + 1, ly * lx * gy);
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(256) k_grid_sample_bwd(const float* __restrict__ inp, int h, int w, int cv, int ld,
                                                         const float* __restrict__ deform, int d, int h0, int w0,
                                                         int mode, const float* __restrict__ dout, int ldo,
                                                         float* __restrict__ dinp, int lddi,
                                                         float* __restrict__ ddeform, const Patch pt,
                                                         const FastDiv ftx, const FastDiv fty) {
    unsigned tx, ty;
    const unsigned t1 = fd_divmod(blockIdx.x, ftx, tx);
    const unsigned n = fd_divmod(t1, fty, ty);
    const long long sb = (long long)(n / (unsigned)d) * h * w;  // first pixel of the source frame
    int ho, wo;
    const WarpTap mine = lane_tap(deform, n, h0, w0, mode, h, w, tx, ty, pt, ho, wo);
    if (__ballot_sync(0xffffffffu, mine.opix >= 0) == 0) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane >> pt.lpp_sh, cl = lane & (pt.lpp - 1);
    const int ppp = 32 >> pt.lpp_sh;
    float my_gx = 0.f, my_gy = 0.f;  // d(loss)/d(sampling coordinate) of THIS lane's phase-1 pixel, in pixels
    for (int q0 = 0; q0 < pt.P; q0 += ppp) {
        const WarpTap t = bcast_tap(mine, (q0 + sub) & 31);
        float gix = 0.f, giy = 0.f;
        if (t.opix >= 0) {
            const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
            const bool b0 = (t.bounds & 5) == 5, b1 = (t.bounds & 6) == 6, b2 = (t.bounds & 9) == 9,
                       b3 = (t.bounds & 10) == 10;
            const long long o00 = sb + t.base;
            const float* gp = dout + ((long long)n * h * w + t.opix) * ldo;
            for (int g = 0; g < pt.cgroups; ++g) {
                const int cq = g * pt.lpp + cl;
                if (cq >= cv) break;
                const int c = cq * 4;
                const float4 gv = ldg4(gp + c);
                if (b0) {
                    if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + o00 * lddi + c), gv * (wx0 * wy0));
                    const float s = dot4(ldg4(inp + o00 * ld + c), gv);
                    gix -= s * wy0; giy -= s * wx0;
                }
                if (b1) {
                    if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + (o00 + 1) * lddi + c), gv * (t.wx1 * wy0));
                    const float s = dot4(ldg4(inp + (o00 + 1) * ld + c), gv);
                    gix += s * wy0; giy -= s * t.wx1;
                }
                if (b2) {
                    if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + (o00 + w) * lddi + c), gv * (wx0 * t.wy1));
                    const float s = dot4(ldg4(inp + (o00 + w) * ld + c), gv);
                    gix -= s * t.wy1; giy += s * wx0;
                }
                if (b3) {
                    if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + (o00 + w + 1) * lddi + c), gv * (t.wx1 * t.wy1));
                    const float s = dot4(ldg4(inp + (o00 + w + 1) * ld + c), gv);
                    gix += s * t.wy1; giy += s * t.wx1;
                }
            }
        }
        if (ddeform) {
            // sum over the lpp channel lanes of each pixel slot, then hand the total to the lane that owns the pixel
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                if (o < pt.lpp) {
                    gix += __shfl_xor_sync(0xffffffffu, gix, o);
                    giy += __shfl_xor_sync(0xffffffffu, giy, o);
                }
            }
            // owner lane of pixel q is lane q; it reads the total from the first lane of slot (q - q0)
            const int rel = lane - q0;
            const int from = (rel >= 0 && rel < ppp) ? (rel << pt.lpp_sh) : 0;
            const float tx_ = __shfl_sync(0xffffffffu, gix, from), ty_ = __shfl_sync(0xffffffffu, giy, from);
            if (rel >= 0 && rel < ppp) { my_gx = tx_; my_gy = ty_; }
        }
    }
    if (ddeform && mine.opix >= 0)
        scatter_grid_grad(ddeform, n, h0, w0, ho, wo, h, w, mode, my_gx * 0.5f * (float)(w - 1),
                          my_gy * 0.5f * (float)(h - 1));
}

MK_EXPORT int mk_grid_sample_bwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d,
                                 int h0, int w0, int mode, const float* dout, int ldo, float* dinp, int lddi,
                                 float* ddeform, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && (!dinp || lddi % 4 == 0),
               "mk_grid_sample_bwd: channels must be x4");
    const int cv = Cp / 4;
    const long long total = (long long)B * d * h * w * cv;
    if (total == 0) return 0;
    const Patch pt = make_patch(h, w, cv, (long long)B * d);
    const long long blocks = (long long)B * d * pt.tiles_x * pt.tiles_y;
    MK_REQUIRE(blocks < (1LL << 31) && (long long)h * w < (1LL << 30), "mk_grid_sample_bwd: extent too large");
    k_grid_sample_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        inp, h, w, cv, ld, deform, d, h0, w0, mode, dout, ldo, dinp, lddi, ddeform, pt, make_fastdiv(pt.tiles_x),
        make_fastdiv(pt.tiles_y));
    return mk_check_launch("mk_grid_sample_bwd");
}

// ------------------------------------------------------------------------------------------------ plain resize
__global__ void __launch_bounds__(256) k_resize_fwd(const float* __restrict__ x, int h0, int w0, int cv, int ld,
                                                    int mode, float* __restrict__ out, int h, int w, int ldo,
                                                    long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * 4;
        const long long op = i / cv;
        const int wo = (int)(op % w);
        const long long t = op / w;
        const int ho = (int)(t % h);
        const long long n = t / h;
        const float* base = x + n * (long long)h0 * w0 * ld + c;
        float4 r;
        if (mode == 0) {
            int ys = nearest_src(ho, h0, h), xs = nearest_src(wo, w0, w);
            r = ldg4(base + ((long long)ys * w0 + xs) * ld);
        } else {
            int y0, y1, x0, x1;
            float ly, lx;
            linear_src(ho, h0, h, y0, y1, ly);
            linear_src(wo, w0, w, x0, x1, lx);
            float hy = 1.f - ly, hx = 1.f - lx;
            float4 a = ldg4(base + ((long long)y0 * w0 + x0) * ld), b = ldg4(base + ((long long)y0 * w0 + x1) * ld);
            float4 cc = ldg4(base + ((long long)y1 * w0 + x0) * ld), dd = ldg4(base + ((long long)y1 * w0 + x1) * ld);
            r = (a * hx + b * lx) * hy + (cc * hx + dd * lx) * ly;
        }
        st4(out + op * ldo + c, r);
    }
}

MK_EXPORT int mk_resize_fwd(const float* x, int N, int h0, int w0, int Cp, int ld, int mode, float* out, int h, int w,
                            int ldo, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0, "mk_resize_fwd: channels must be x4");
    const long long total = (long long)N * h * w * (Cp / 4);
    if (total == 0) return 0;
    long long blocks = mk_cdiv(total, 256);
    const long long cap = 16LL * mk_num_sms();
    if (blocks > cap) blocks = cap;
    k_resize_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, h0, w0, Cp / 4, ld, mode, out, h, w, ldo, total);
    return mk_check_launch("mk_resize_fwd");
}

__global__ void __launch_bounds__(256) k_resize_bwd(const float* __restrict__ dout, int h, int w, int cv, int ldo,
                                                    int mode, float* __restrict__ dx, int h0, int w0, int ld,
                                                    long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * 4;
        const long long op = i / cv;
        const int wo = (int)(op % w);
        const long long t = op / w;
        const int ho = (int)(t % h);
        const long long n = t / h;
        float* base = dx + n * (long long)h0 * w0 * ld + c;
        const float4 g = ldg4(dout + (long long)op * ldo + c);
        if (mode == 0) {
            int ys = nearest_src(ho, h0, h), xs = nearest_src(wo, w0, w);
            atomicAdd(reinterpret_cast<float4*>(base + ((long long)ys * w0 + xs) * ld), g);
        } else {
            int y0, y1, x0, x1;
            float ly, lx;
            linear_src(ho, h0, h, y0, y1, ly);
            linear_src(wo, w0, w, x0, x1, lx);
            float hy = 1.f - ly, hx = 1.f - lx;
            atomicAdd(reinterpret_cast<float4*>(base + ((long long)y0 * w0 + x0) * ld), g * (hy * hx));
            atomicAdd(reinterpret_cast<float4*>(base + ((long long)y0 * w0 + x1) * ld), g * (hy * lx));
            atomicAdd(reinterpret_cast<float4*>(base + ((long long)y1 * w0 + x0) * ld), g * (ly * hx));
            atomicAdd(reinterpret_cast<float4*>(base + ((long long)y1 * w0 + x1) * ld), g * (ly * lx));
        }
    }
}

MK_EXPORT int mk_resize_bwd(const float* dout, int N, int h, int w, int Cp, int ldo, int mode, float* dx, int h0,
                            int w0, int ld, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0, "mk_resize_bwd: channels must be x4");
    const long long total = (long long)N * h * w * (Cp / 4);
    if (total == 0) return 0;
    long long blocks = mk_cdiv(total, 256);
    const long long cap = 16LL * mk_num_sms();
    if (blocks > cap) blocks = cap;
    k_resize_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, h, w, Cp / 4, ldo, mode, dx, h0, w0, ld, total);
    return mk_check_launch("mk_resize_bwd");
}
