// Deformation sampling: F.grid_sample (5-D call with D_in = 1, z == 0  ==  2-D bilinear, zeros padding,
// align_corners=True under torch 0.4.1) with the deformation-grid resize of generator.py:51-58 fused in, plus the
// plain NHWC resize used for the keypoint-embedding skips (generator.py:72).
//
// HBM-bound.  NHWC makes every bilinear tap a contiguous channel vector: one thread owns one float4 of channels of
// one output pixel, so the 4 taps are 128-bit loads that coalesce across the lanes of a pixel, and the output is a
// coalesced 128-bit store.  Algorithmic bytes per call: 4*(B*C*h*w + 2*B*d*h*w + B*d*C*h*w).
#include "common.cuh"
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

// Work decomposition (forward and backward): one CTA owns a PH x PW patch of output pixels of one frame and all of its
// channel vectors; item = (pixel of the patch, float4 of channels), channel vector fastest, so the lanes of a warp
// are the channels of 1-2 neighbouring pixels (128-bit loads that coalesce into whole 128 B lines per tap).  The
// four taps of neighbouring pixels overlap (x0+1 of pixel w is x0 of pixel w+1, row y0+1 of this row is row y0 of the
// next): with the previous row-linear mapping every tap line was fetched again from L2 by another CTA (ncu: L1 hit
// 21 %, L2->SM traffic 4x the input, the kernel ran at the L2 bandwidth, 48 % of HBM); a 2-D patch keeps that reuse
// inside one SM's L1.
struct Patch {
    int ph, pw, pw_sh, tiles_x, tiles_y, items;  // pw = 1 << pw_sh; items = ph * pw * cv
};

static inline Patch make_patch(int h, int w, int cv) {
    Patch t;
    const int side = cv >= 16 ? 8 : 16;
    t.ph = side; t.pw = side; t.pw_sh = side == 8 ? 3 : 4;
    t.tiles_x = (w + t.pw - 1) / t.pw; t.tiles_y = (h + t.ph - 1) / t.ph;
    t.items = t.ph * t.pw * cv;
    return t;
}

__global__ void __launch_bounds__(256) k_grid_sample_fwd(const float* __restrict__ inp, int h, int w, int cv, int ld,
                                                         const float* __restrict__ deform, int d, int h0, int w0,
                                                         int mode, float* __restrict__ out, int ldo, const Patch pt,
                                                         const FastDiv fcv, const FastDiv ftx, const FastDiv fty) {
    unsigned tx, ty;
    const unsigned t1 = fd_divmod(blockIdx.x, ftx, tx);
    const unsigned n = fd_divmod(t1, fty, ty);
    const float* src = inp + (long long)(n / (unsigned)d) * h * w * ld;
    float* dst = out + (long long)n * h * w * ldo;
#pragma unroll 2
    for (unsigned item = threadIdx.x; item < (unsigned)pt.items; item += 256u) {
        unsigned cq;
        const unsigned p = fd_divmod(item, fcv, cq);
        const int ho = (int)(ty * pt.ph + (p >> pt.pw_sh)), wo = (int)(tx * pt.pw + (p & (pt.pw - 1)));
        if (ho >= h || wo >= w) continue;
        const int c = (int)cq * 4;
        const Tap tp = make_tap(fetch_grid(deform, n, h0, w0, ho, wo, h, w, mode), h, w);
        const bool xin0 = tp.x0 >= 0 && tp.x0 < w, xin1 = tp.x0 + 1 >= 0 && tp.x0 + 1 < w;
        const bool yin0 = tp.y0 >= 0 && tp.y0 < h, yin1 = tp.y0 + 1 >= 0 && tp.y0 + 1 < h;
        const float wx0 = 1.f - tp.wx1, wy0 = 1.f - tp.wy1;
        const float* s0 = src + (tp.y0 * w + tp.x0) * (long long)ld + c;
        // issue the four tap loads before blending (ATen order: nw, ne, sw, se)
        const float4 v0 = (yin0 && xin0) ? ldg4(s0) : f4zero();
        const float4 v1 = (yin0 && xin1) ? ldg4(s0 + ld) : f4zero();
        const float4 v2 = (yin1 && xin0) ? ldg4(s0 + (long long)w * ld) : f4zero();
        const float4 v3 = (yin1 && xin1) ? ldg4(s0 + (long long)w * ld + ld) : f4zero();
        float4 acc = f4zero();
        fma4(acc, v0, wx0 * wy0);
        fma4(acc, v1, tp.wx1 * wy0);
        fma4(acc, v2, wx0 * tp.wy1);
        fma4(acc, v3, tp.wx1 * tp.wy1);
        __stcs(reinterpret_cast<float4*>(dst + ((long long)ho * w + wo) * ldo + c), acc);  // streamed: never re-read here
    }
}

MK_EXPORT int mk_grid_sample_fwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d,
                                 int h0, int w0, int mode, float* out, int ldo, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0, "mk_grid_sample_fwd: channels must be x4");
    const long long total = (long long)B * d * h * w * (Cp / 4);
    if (total == 0) return 0;
    const Patch pt = make_patch(h, w, Cp / 4);
    const long long blocks = (long long)B * d * pt.tiles_x * pt.tiles_y;
    MK_REQUIRE(blocks < (1LL << 31) && (long long)h * w * ld < (1LL << 31), "mk_grid_sample_fwd: extent too large");
    k_grid_sample_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(inp, h, w, Cp / 4, ld, deform, d, h0, w0, mode,
                                                                          out, ldo, pt, make_fastdiv(Cp / 4),
                                                                          make_fastdiv(pt.tiles_x),
                                                                          make_fastdiv(pt.tiles_y));
    return mk_check_launch("mk_grid_sample_fwd");
}

// backward: dinp via vector atomics (red.global.add.v4.f32), d(grid) reduced over the channel lanes of a pixel
// (segmented warp shuffle when the lane group is a power of two <= 32), then chained through the resize into the
// coarse deformation gradient with scalar atomics.
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
                                                         float* __restrict__ ddeform, int seg, const Patch pt,
                                                         const FastDiv fcv, const FastDiv ftx, const FastDiv fty) {
    // same patch decomposition as the forward kernel.  pt.items is a multiple of 32 and the item loop advances whole
    // warps, so the lanes of one pixel (cv consecutive items) never straddle a loop boundary; pixels outside the
    // image (partial patches) take part in the shuffles with zero contributions.
    unsigned tx, ty;
    const unsigned t1 = fd_divmod(blockIdx.x, ftx, tx);
    const unsigned n = fd_divmod(t1, fty, ty);
    const long long sb = (long long)(n / (unsigned)d) * h * w;
    for (unsigned item = threadIdx.x; item < (unsigned)pt.items; item += 256u) {
        unsigned cq;
        const unsigned p = fd_divmod(item, fcv, cq);
        const int ho = (int)(ty * pt.ph + (p >> pt.pw_sh)), wo = (int)(tx * pt.pw + (p & (pt.pw - 1)));
        const bool inside = ho < h && wo < w;
        const int c = (int)cq * 4;
        float gix = 0.f, giy = 0.f;
        if (inside) {
            const Tap tp = make_tap(fetch_grid(deform, n, h0, w0, ho, wo, h, w, mode), h, w);
            const bool xin0 = tp.x0 >= 0 && tp.x0 < w, xin1 = tp.x0 + 1 >= 0 && tp.x0 + 1 < w;
            const bool yin0 = tp.y0 >= 0 && tp.y0 < h, yin1 = tp.y0 + 1 >= 0 && tp.y0 + 1 < h;
            const float wx0 = 1.f - tp.wx1, wy0 = 1.f - tp.wy1;
            const float4 g = ldg4(dout + ((long long)n * h * w + (long long)ho * w + wo) * ldo + c);
            const long long o00 = sb + (long long)tp.y0 * w + tp.x0, o01 = o00 + 1, o10 = o00 + w, o11 = o10 + 1;
            if (yin0 && xin0) {
                if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + o00 * lddi + c), g * (wx0 * wy0));
                float s = dot4(ldg4(inp + o00 * ld + c), g);
                gix -= s * wy0; giy -= s * wx0;
            }
            if (yin0 && xin1) {
                if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + o01 * lddi + c), g * (tp.wx1 * wy0));
                float s = dot4(ldg4(inp + o01 * ld + c), g);
                gix += s * wy0; giy -= s * tp.wx1;
            }
            if (yin1 && xin0) {
                if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + o10 * lddi + c), g * (wx0 * tp.wy1));
                float s = dot4(ldg4(inp + o10 * ld + c), g);
                gix -= s * tp.wy1; giy += s * wx0;
            }
            if (yin1 && xin1) {
                if (dinp) atomicAdd(reinterpret_cast<float4*>(dinp + o11 * lddi + c), g * (tp.wx1 * tp.wy1));
                float s = dot4(ldg4(inp + o11 * ld + c), g);
                gix += s * tp.wy1; giy += s * tp.wx1;
            }
        }
        if (ddeform) {
            if (seg > 1) {
                // lanes of one pixel are `seg` consecutive lanes (seg = cv, power of two <= 32) or a whole warp
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    if (o < seg) {
                        gix += __shfl_xor_sync(0xffffffffu, gix, o);
                        giy += __shfl_xor_sync(0xffffffffu, giy, o);
                    }
                }
                if (((threadIdx.x & 31) & (seg - 1)) != 0) continue;
            }
            if (inside)
                scatter_grid_grad(ddeform, n, h0, w0, ho, wo, h, w, mode, gix * 0.5f * (float)(w - 1),
                                  giy * 0.5f * (float)(h - 1));
        }
    }
}

MK_EXPORT int mk_grid_sample_bwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d,
                                 int h0, int w0, int mode, const float* dout, int ldo, float* dinp, int lddi,
                                 float* ddeform, void* stream) {
    MK_REQUIRE(Cp % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && (!dinp || lddi % 4 == 0),
               "mk_grid_sample_bwd: channels must be x4");
    const int cv = Cp / 4;
    const long long total = (long long)B * d * h * w * cv;
    if (total == 0) return 0;
    const Patch pt = make_patch(h, w, cv);
    // segmented shuffle reduction over the channel lanes of a pixel: cv a power of two <= 32 (a pixel = `cv` aligned
    // lanes) or a multiple of 32 (whole warps per pixel); otherwise per-thread atomics (seg = 1).  Patches hold a
    // multiple of 32 items, so warps are never partial.
    int seg = 1;
    if ((cv & (cv - 1)) == 0 && cv <= 32) seg = cv;
    else if (cv % 32 == 0) seg = 32;
    const long long blocks = (long long)B * d * pt.tiles_x * pt.tiles_y;
    MK_REQUIRE(blocks < (1LL << 31), "mk_grid_sample_bwd: extent too large");
    k_grid_sample_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        inp, h, w, cv, ld, deform, d, h0, w0, mode, dout, ldo, dinp, lddi, ddeform, seg, pt, make_fastdiv(cv),
        make_fastdiv(pt.tiles_x), make_fastdiv(pt.tiles_y));
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
