/*
 * monkey_b200.h - C ABI of libmonkey_b200.so (sm_100a).
 *
 * The reference (AliaksandrSiarohin/monkey-net) has no FFI layer: its hot path is Python modules that call
 * torch/ATen/cuDNN ops.  The drop-in boundary is therefore the Python module API
 * (`modules/*`, `sync_batchnorm/*`); this header is the C ABI *underneath* that API - one entry
 * point per library call site the reference makes on the path, each citing the reference line it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless noted; sizes are element counts;
 *   - activations are NHWC: `[N][H][W][ld]`, `N` = B*D frames (the reference's Conv3d kernels are (1,k,k), so D
 *     folds into batch), `ld` = pixel stride in floats (>= physical channel count), all physical channel
 *     counts are multiples of 4 with zero padding channels;
 *   - `stream` is a `cudaStream_t` passed as `void*`; all work is asynchronous on it;
 *   - return value 0 = success, otherwise a cudaError_t / negative argument-error code; `mk_last_error()` has
 *     the text (thread-local).  The Python wrapper raises RuntimeError on non-zero.
 *   - the library is re-entrant and holds no global state besides a per-thread error string.
 */
#ifndef MONKEY_B200_H
#define MONKEY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* mk_last_error(void);
int mk_version(void);
int mk_fill_zero(void* ptr, long long bytes, void* stream);
/* benchmark helper: evict the L2 by READING `bytes` (>= 2x the 126 MB L2) of `buf`; unlike a memset it leaves no
 * dirty lines behind to be written back during the timed kernel.  `buf` must be 16-byte aligned. */
int mk_l2_evict(const void* buf, long long bytes, void* stream);

/* ---- layout edge: reference NCDHW tensors <-> internal NHWC (SURVEY 8(b) "convert only at this edge") ----
 * src is a 5-D (B,C,D,H,W) tensor with arbitrary element strides; `step` implements the nearest down-scale
 * F.interpolate(scale_factor=(1,1/step,1/step)) (keypoint_detector.py:99, dense_motion_module.py:44,
 * movement_embedding.py:44, discriminator.py:67): dst[n=b*D+d][h][w][c] = src[b][c][d][h*step][w*step].
 * dst is [B*D][H/step][W/step][ld]; channels C..Cp-1 are written as zero. */
int mk_ncdhw_to_nhwc(const float* src, int B, int C, int D, int H, int W,
                     long long sb, long long sc, long long sd, long long sh, long long sw,
                     int step, float* dst, int Cp, int ld, void* stream);
/* inverse (also the backward of the above): dst[b][c][d][h*step][w*step] (+)= src[n][h][w][c]; dst strides given.
 * When step > 1 the caller zero-fills dst first. */
int mk_nhwc_to_ncdhw(const float* src, int ld, int B, int C, int D, int Hs, int Ws, int step,
                     float* dst, long long sb, long long sc, long long sd, long long sh, long long sw, void* stream);
/* strided channel-slice copy between NHWC buffers: dst[p][0..C) = src[p][0..C) (C % 4 == 0) - concat / split. */
int mk_copy_channels(const float* src, int lds, float* dst, int ldd, long long npix, int C, void* stream);
/* dst[p][j] = map[j] >= 0 ? src[p][map[j]] : 0 for j < Cd - compacts a concat-with-holes tensor (torch.cat at
 * util.py:187 followed by the ResBlocks, generator.py:78-79); with the inverse map it scatters gradients back. */
int mk_gather_channels(const float* src, int lds, const int* map, float* dst, int ldd, long long npix, int Cd,
                       void* stream);
/* dst[p][c] += src[p][c] */
int mk_add_channels(const float* src, int lds, float* dst, int ldd, long long npix, int C, void* stream);

/* ---- convolution: nn.Conv3d (1,k,k) (util.py:52-55,79-80,98-99,176-177; discriminator.py:17-18;
 *      dense_motion_module.py:26-27 grouped 1x1; generator.py:48 1x1) -------------------------------------------
 * Weight packing: w is the reference parameter (Co, Ci/groups, 1, R, S).  wpack is [R*S][Kin_p][Kout_p].
 *   mode 0 (forward):  Kin = input channels,  Kout = output channels, tap = r*S+s.
 *   mode 1 (dgrad):    Kin = output channels, Kout = input channels,  tap flipped (R-1-r, S-1-s).
 *   mode 4:            sub-pixel forward pack for the upsampled 3x3 conv, [4 parities][4 taps][Kout_p][Kin_p], TF32.
 *   modes 2 / 3:       modes 0 / 1 in the tensor-core layout [R*S][Kout_p][Kin_p], rounded to TF32 (mk_conv2d_tc).
 *   mode | 8 (with 2, 3, 4): 3xTF32 pack - wpack holds 2x the floats, the TF32 remainders lo = rna(w - hi) in the
 *                      same layout behind the hi half (mk_conv2d_tc_x3).
 * `cin_map[Cin_p]` maps each PHYSICAL input channel to its logical index or -1 (padding / concat holes);
 * physical output channel j is logical j for j < Co, padding otherwise.  Grouped convs are packed block-diagonal. */
int mk_pack_weight(const float* w, int Co, int Cig, int R, int S, int groups, const int* cin_map, int Cin_p,
                   int Cout_p, int mode, float* wpack, const float* bias /* (Co) or NULL */,
                   float* bias_p /* [Cout_p] zero-padded copy, or NULL */, void* stream);
/* gather the packed weight gradient (mode-0 layout) back into the parameter layout (Co,Ci/groups,1,R,S);
 * cin_inv[Ci] maps each LOGICAL input channel to its physical position (NULL = identity). */
int mk_unpack_wgrad(const float* dwpack, int Co, int Cig, int R, int S, int groups, const int* cin_inv, int Cin_p,
                    int Cout_p, float* dw, void* stream);
/* the same, accumulating (+=) into the parameter's gradient tensor `grad` (parameter layout): autograd's
 * AccumulateGrad for conv weights done by the unpack kernel itself (torch/autograd: one add kernel per parameter). */
int mk_unpack_wgrad_acc(const float* dwpack, int Co, int Cig, int R, int S, int groups, const int* cin_inv, int Cin_p,
                        int Cout_p, float* grad, void* stream);
/* dz = dy * y * (1 - y)  (backward of the fused sigmoid epilogue, generator.py:80); n floats, n % 4 == 0 */
int mk_sigmoid_bwd(const float* y, const float* dy, float* dz, long long n, void* stream);
/* y = epilogue(conv(x)); implicit GEMM, fp32 FFMA (exact-parity path).
 *   ups=1: x is nearest-upsampled x2 on the fly (util.py:84 F.interpolate folded into the gather).
 *   (Hl,Wl) = logical input size = (Hin,Win) << ups;  Ho = Hl + 2*pad - R + 1.
 *   epilogue per output channel: v = acc*scale[c] + shift[c] (scale==NULL -> 1, shift==NULL -> 0; shift is the
 *   bias or the folded eval-mode BN), then + resid, then act (0 none, 1 relu/leaky with `slope`, 2 sigmoid),
 *   then pool (0 none, 1 = 2x2 average, floor; 2 = 2x2 sum) -> y is [N][Ho>>1][Wo>>1][ldy] when pooled. */
int mk_conv2d(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups,
              const float* wpack, int R, int S, int pad,
              const float* scale, const float* shift, const float* resid, int ldr, int act, float slope,
              float* y, int Cout_p, int ldy, int pool, void* stream);
/* Tensor-core variant of mk_conv2d for stride-1 convs without upsample/pool: 4-D TMA boxes of the NHWC activation
 * (out-of-bounds zero fill == conv padding) -> 128B-swizzled smem -> tcgen05.mma kind::tf32, TMEM accumulators.
 * wpack_tc = mk_pack_weight mode 2 (forward) / 3 (dgrad): layout [tap][Kout_p][Kin_p], values rounded to TF32.
 * ups=1 (R=S=3, pad=1 only): conv3x3(nearest-upsample-x2(x)) (util.py:71-88) computed as FOUR 2x2 sub-pixel convs on
 * the low-resolution grid, one per output parity, with pre-summed taps (mk_pack_weight mode 4:
 * [parity][2x2 tap][Kout_p][Kin_p]) - 2.25x fewer FLOPs than convolving the upsampled tensor.
 * Any channel counts that are multiples of 4 (ragged counts ride on the TMA zero fill); returns -2 (and touches
 * nothing) for unaligned strides so the caller can use mk_conv2d. */
int mk_conv2d_tc(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups, const float* wpack_tc, int R,
                 int S, int pad, const float* scale, const float* shift, const float* resid, int ldr, int act, float slope,
                 float* y, int Cout_p, int ldy, void* stream);
/* 3xTF32 ("reference precision") variant of mk_conv2d_tc: every fp32 operand is split hi + lo (hi = rna_tf32(v),
 * lo = rna_tf32(v - hi)) and A_lo*B_hi + A_hi*B_lo + A_hi*B_hi accumulate in the same TMEM tile - fp32-accurate
 * products (2^-22 relative) at a third of the TF32 issue rate.  wpack_tc = mk_pack_weight mode (2|3|4) | 8: the lo
 * half follows the hi half (2x the floats); the activation tile is split in shared memory by idle warps. */
int mk_conv2d_tc_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups, const float* wpack_tc, int R,
                    int S, int pad, const float* scale, const float* shift, const float* resid, int ldr, int act,
                    float slope, float* y, int Cout_p, int ldy, void* stream);
/* Dry runs of the two tensor-core entry points' host-side planning (tiling, split-K / pixel splits, ring depths,
 * shared-memory and TMEM budgets); they touch no device state and work without a GPU (148 SMs assumed).  out[16]:
 * see csrc/conv_tc.cu / csrc/wgrad_tc.cu.  tests/test_tc_plans.py sweeps every layer of the shipped configs.
 * The 3xTF32 plans are selected with `ups | 2` (conv) and `pad | 256` (wgrad). */
int mk_conv2d_tc_plan(int N, int Hin, int Win, int Cin_p, int ups, int R, int S, int pad, int act, int Cout_p, int ldy,
                      int* out);
int mk_conv2d_wgrad_tc_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int R, int S, int pad, int* out);
/* Halo-window PERSISTENT tensor-core convolution (csrc/conv_halo.cu), the default for the many-tile stride-1 layers.
 * Same contract and epilogue as mk_conv2d_tc (ups = 0).  One TMA box per 32-channel chunk lands the (8*RB + R) x 16
 * pixel halo of a super-tile; every filter tap of every 8 x (17-S) row-block is a row-shifted window of that buffer
 * (UMMA descriptor start address, 128B swizzle on absolute address bits), the weights stay resident in shared memory
 * for the life of the CTA when they fit, TMEM accumulators are double-buffered and the output leaves through TMA
 * stores.  Returns -2 (nothing touched) outside its envelope (R, S <= 4) or when the layer has too few tiles for a
 * persistent launch - callers then use mk_conv2d_tc.  The _x3 variant is the 3xTF32 mode (wpack mode | 8). */
int mk_conv2d_tc_halo(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc, int R, int S,
                      int pad, const float* scale, const float* shift, const float* resid, int ldr, int act,
                      float slope, float* y, int Cout_p, int ldy, void* stream);
int mk_conv2d_tc_halo_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_tc, int R,
                         int S, int pad, const float* scale, const float* shift, const float* resid, int ldr, int act,
                         float slope, float* y, int Cout_p, int ldy, void* stream);
/* conv3x3(nearest_x2(x)), pad 1 (modules/util.py:84-85) on the halo-window kernel: four sub-pixel 2x2 passes over the
 * low-resolution grid, each storing its output parity through a 5-D TMA map.  wpack_ups = mk_pack_weight mode 4 (| 8 for
 * the _x3 entry); y [N][2 Hin][2 Win][ldy].  -2 (nothing launched) outside the envelope: callers use mk_conv2d_tc(ups=1). */
int mk_conv2d_tc_halo_ups(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_ups,
                          const float* scale, const float* shift, int act, float slope, float* y, int Cout_p, int ldy,
                          void* stream);
int mk_conv2d_tc_halo_ups_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* wpack_ups,
                             const float* scale, const float* shift, int act, float slope, float* y, int Cout_p, int ldy,
                             void* stream);
/* dry run of its planner: out[16] = grid.x, cout tiles, RB, smem bytes, halo stages, weight slots, resident?, TMEM
 * columns, tiles, halo rows, halo stage bytes, weight slot bytes, valid tile width, output groups, acc columns, x3 */
int mk_conv2d_tc_halo_plan(int N, int Hin, int Win, int Cin_p, int R, int S, int pad, int Cout_p, int has_resid, int x3,
                           int* out);
int mk_conv2d_tc_halo_ups_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int x3, int* out /* host, 16 ints */);
/* dwpack[R*S][Cin_p][Cout_p] = sum over pixels of im2col(x)^T dy  (zero-filled inside). */
int mk_conv2d_wgrad(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, int ups,
                    const float* dy, int Cout_p, int ldy, int R, int S, int pad, float* dwpack, void* stream);

/* Tensor-core variant of mk_conv2d_wgrad (no upsample): both GEMM operands are consumed MN-major straight from the
 * NHWC tensors (tcgen05 a_major = b_major = 1), K = pixels, split over pixel ranges with fp32 atomics.  Same output
 * layout as mk_conv2d_wgrad.  Channel counts must be multiples of 4; returns -2 otherwise. */
int mk_conv2d_wgrad_tc(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                       int ldy, int R, int S, int pad, float* dwpack, void* stream);
/* Halo-window tensor-core weight gradient (csrc/wgrad_halo.cu), the default for 3x3 / 4x4 layers with enough pixels:
 * the X halo of a pixel tile is loaded once per 32-channel chunk and every filter tap is a row-shifted window of it; the
 * tap ROWS ride on the MMA's M dimension (four 32-channel blocks one image row apart, descriptor leading-dimension
 * offset), so D[32 r + ci][co] = dW[(r, s)][ci][co] and N = co is as narrow as the layer.  Same output layout as
 * mk_conv2d_wgrad.  Returns -2 (nothing touched) outside its envelope - callers then use mk_conv2d_wgrad_tc.  _x3 =
 * 3xTF32 (both operands split hi / lo in shared memory). */
int mk_conv2d_wgrad_halo(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                         int ldy, int R, int S, int pad, float* dwpack, void* stream);
int mk_conv2d_wgrad_halo_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                            int ldy, int R, int S, int pad, float* dwpack, void* stream);
/* weight gradient of conv3x3(nearest_x2(x)), pad 1, as four sub-pixel passes on the LOW-resolution grid (16 tap-pixel
 * products per low-res pixel instead of 36 at full resolution, no upsampled copy of x): x [N][Hin][Win][ldx],
 * dy [N][2 Hin][2 Win][ldy] read through a 5-D TMA map; dwpack_ups [16 = parity x 2x2 tap][Cin_p][Cout_p] is the gradient
 * of the mode-4 pack, folded onto the parameter's 3x3 taps by mk_unpack_wgrad_ups (accumulate != 0: += into `dw`).
 * -2 (nothing launched) outside the envelope (Hin % 8, few tiles): callers upsample x and use mk_conv2d_wgrad_halo. */
int mk_conv2d_wgrad_halo_ups(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                             int ldy, float* dwpack_ups, void* stream);
int mk_conv2d_wgrad_halo_ups_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                                int ldy, float* dwpack_ups, void* stream);
int mk_unpack_wgrad_ups(const float* dwpack_ups, int Co, int Ci, const int* cin_inv, int Cin_p, int Cout_p, float* dw,
                        int accumulate, void* stream);
/* dry run: out[16] = co tiles, ci groups, pixel splits, smem bytes, stages, tile rows, ci chunks per CTA, dY boxes, TMEM
 * columns, tiles, tiles per split, UMMA N, stage bytes, x3, valid tile width, tile rows of the image */
int mk_conv2d_wgrad_halo_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int R, int S, int pad, int x3, int* out);
int mk_conv2d_wgrad_halo_ups_plan(int N, int Hin, int Win, int Cin_p, int Cout_p, int x3, int* out /* host, 16 ints */);
/* 3xTF32 variant (see mk_conv2d_tc_x3): both operand tiles are split hi + lo in shared memory. */
int mk_conv2d_wgrad_tc_x3(const float* x, int N, int Hin, int Win, int Cin_p, int ldx, const float* dy, int Cout_p,
                          int ldy, int R, int S, int pad, float* dwpack, void* stream);

/* ---- normalisation: BatchNorm over (N,D,H,W) (sync_batchnorm/batchnorm.py:48-78 -> F.batch_norm semantics on the
 *      global batch) and InstanceNorm3d (discriminator.py:20) -------------------------------------------------
 * colstats: sums[g][0][c] = sum x, sums[g][1][c] = sum x^2 over the pixels of group g; groups = 1 (batch norm,
 * bias gradients) or N (instance norm).  `sums` is [groups][2][Cp].  With several GPUs the [2][Cp] block is what
 * gets all-reduced (ONE NCCL all-reduce per BN layer per direction). */
int mk_colstats(const float* x, int ld, int N, long long hw, int Cp, int per_frame, float* sums, void* stream);
/* the same sums accumulated and returned in DOUBLE precision (forward statistics of batch / instance norm: the
 * variance E[x^2]-E[x]^2 cancels in fp32 when mean^2 >> var).  This [groups][2][Cp] double block is what the
 * forward all-reduce carries. */
int mk_colstats_f64(const float* x, int ld, int N, long long hw, int Cp, int per_frame, double* sums, void* stream);
/* mean/invstd/scale/shift from (possibly all-reduced) double-precision sums of mk_colstats_f64.  count = pixels per group (global).  gamma/beta are
 * the logical-length-C parameters; padding channels get scale = shift = 0.  If running_mean != NULL the running
 * stats are updated (momentum, unbiased variance) and *num_batches_tracked (int64) is incremented.
 * out = [groups][4][Cp]: mean, invstd, scale, shift. */
int mk_norm_finalize(const double* sums, int groups, int C, int Cp, double count, const float* gamma,
                     const float* beta, float eps, float* running_mean, float* running_var, float momentum,
                     long long* num_batches_tracked, float* out, void* stream);
/* eval-mode BN: scale/shift from running stats (batchnorm.py:50-53 with training=False); out as above, groups=1 */
int mk_norm_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                        int C, int Cp, float eps, float* out, void* stream);
/* out = pool(act(x*scale+shift)); params = [groups][4][Cp] as above (params==NULL: identity affine);
 * act: slope < 0 -> none, else v>0?v:slope*v (0 = ReLU util.py:64,66,86,105,124; 0.2 = LeakyReLU
 * discriminator.py:29).  pool=1: AvgPool (1,2,2) floor (util.py:101, discriminator.py:30). */
int mk_norm_apply(const float* x, int ldx, int N, int H, int W, int Cp, const float* params, int per_frame,
                  float slope, int pool, float* out, int ldo, void* stream);
/* backward, phase 1: sums[g][0][c] = sum dz, sums[g][1][c] = sum dz*xhat where z = x*scale+shift,
 * dz = act'(z) * unpool(dout). */
int mk_norm_bwd_reduce(const float* x, int ldx, const float* dout, int ldd, int N, int H, int W, int Cp,
                       const float* params, int per_frame, float slope, int pool, float* sums, void* stream);
/* backward, phase 2: dx = scale*(dz - sum_dz/count - xhat*sum_dzxhat/count) (normed=1) or dx = dz (normed=0,
 * params may be NULL).  sums are the (all-reduced) phase-1 sums; count is global. */
int mk_norm_bwd_apply(const float* x, int ldx, const float* dout, int ldd, int N, int H, int W, int Cp,
                      const float* params, const float* sums, double count, int per_frame, int normed,
                      float slope, int pool, float* dx, int lddx, void* stream);

/* ---- F.grid_sample 5-D with the grid resize fused (generator.py:51-58) ------------------------------------------
 * inp [B][h][w][ld] ; deform [B*d][h0][w0][2] (x,y in [-1,1]); out [B*d][h][w][ldo]; frame n samples source n/d.
 * mode 0: grid resized to (h,w) by nearest (F.interpolate default), 1: bilinear align_corners=False ('trilinear'
 * with D preserved).  Sampling: bilinear, zeros padding, align_corners=True (torch 0.4.1 semantics). */
int mk_grid_sample_fwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d, int h0,
                       int w0, int mode, float* out, int ldo, void* stream);
/* dinp (zero-filled by caller, accumulated with vector atomics) and ddeform (zero-filled by caller). */
int mk_grid_sample_bwd(const float* inp, int B, int h, int w, int Cp, int ld, const float* deform, int d, int h0,
                       int w0, int mode, const float* dout, int ldo, float* dinp, int lddi, float* ddeform,
                       void* stream);
/* F.interpolate(size=...) of an NHWC map (generator.py:72 kp_skips): mode 0 nearest, 1 bilinear(ac=False). */
int mk_resize_fwd(const float* x, int N, int h0, int w0, int Cp, int ld, int mode, float* out, int h, int w,
                  int ldo, void* stream);
int mk_resize_bwd(const float* dout, int N, int h, int w, int Cp, int ldo, int mode, float* dx, int h0, int w0,
                  int ld, void* stream); /* dx zero-filled by caller */

/* ---- keypoint head: softmax(heat/T) over H*W + gaussian2kp (keypoint_detector.py:43-78,101-107) ----------------
 * logits [N][H][W][ld] (K logical channels).  mean [N][K][2], var [N][K][4] (row-major 2x2) ==
 * (B,D,K,2)/(B,D,K,2,2) contiguous.  var_mode 0 'matrix', 1 'single' (var [N][K][1]).  clip <= 0: no clip_variance.
 * aux [N][K][8]: softmax max, sum-exp, raw (unclipped) covariance (4), spare - saved for backward.
 * scratch: *out of mk_kp_head_scratch_floats(N, H, W, K, out) floats of device memory for the chunk partials of the
 * many-block kernels (all channels of a pixel chunk per block; partials combined in a fixed order); NULL selects the
 * one-block-per-(frame, keypoint) kernels. */
int mk_kp_head_scratch_floats(int N, int H, int W, int K, long long* out /* host */);
int mk_kp_head_fwd(const float* logits, int N, int H, int W, int K, int ld, float inv_temperature, int var_mode,
                   float clip, float* mean, float* var, float* aux, float* scratch, void* stream);
int mk_kp_head_bwd(const float* logits, int N, int H, int W, int K, int ld, float inv_temperature, int var_mode,
                   float clip, const float* mean, const float* aux, const float* dmean, const float* dvar,
                   float* dlogits /* [N][H][W][ld], pads zeroed inside */, float* scratch, void* stream);

/* ---- movement embedding (movement_embedding.py:42-92) incl. kp2gaussian (keypoint_detector.py:7-40) ------------
 * flags bit0 use_heatmap, bit1 use_difference, bit2 use_deformed_source_image, bit3 add_bg_feature_map,
 * bit4 heatmap_type=='difference'.  var_mode 0 matrix / 1 single / 2 constant (kp_variance float).
 * kp_d_* are (B,d,K,..) = [N][K][..], kp_s_* are (B,1,K,..).  norm_const > 0: divide by it; == 0: 'sum'
 * normalisation using heat_sums [2][N][K] from mk_kp_heat_sums.  src [B][h][w][lds] (C image channels) may be NULL
 * unless bit2.  out [N][h][w][ldo], channels slot*F+f (slot-major), remaining pad channels up to Cout_p zeroed. */
int mk_kp_heat_sums(const float* kd_mean, const float* kd_var, const float* ks_mean, const float* ks_var, int B,
                    int d, int K, int h, int w, int var_mode, float const_var, float* heat_sums, void* stream);
int mk_movement_embed_fwd(const float* src, int lds, int C, const float* kd_mean, const float* kd_var,
                          const float* ks_mean, const float* ks_var, int B, int d, int K, int h, int w, int flags,
                          int var_mode, float const_var, float norm_const, const float* heat_sums, float* out,
                          int Cout_p, int ldo, void* stream);
/* gradients w.r.t. the four kp tensors (zero-filled by caller; source grads accumulate over d). */
int mk_movement_embed_bwd(const float* src, int lds, int C, const float* kd_mean, const float* kd_var,
                          const float* ks_mean, const float* ks_var, int B, int d, int K, int h, int w, int flags,
                          int var_mode, float const_var, float norm_const, const float* heat_sums,
                          const float* dout, int ldo, float* d_kd_mean, float* d_kd_var, float* d_ks_mean,
                          float* d_ks_var, void* stream);

/* ---- dense-motion head (dense_motion_module.py:52-76): channel softmax mask x keypoint shifts + correction +
 *      identity grid.  pred [N][h][w][ld] with (K+1)*use_mask + 2*use_correction logical channels.
 *      deform out [N][h][w][2]. */
int mk_flow_head_fwd(const float* pred, int ld, const float* kd_mean, const float* ks_mean, int B, int d, int K,
                     int h, int w, int use_mask, int use_correction, float* deform, void* stream);
int mk_flow_head_bwd(const float* pred, int ld, const float* kd_mean, const float* ks_mean, int B, int d, int K,
                     int h, int w, int use_mask, int use_correction, const float* ddeform, float* dpred,
                     float* d_kd_mean, float* d_ks_mean /* both zero-filled by caller */, void* stream);

/* ---- losses (modules/losses.py:4-24): per-sample means over logical 5-D (B,C,D,H,W) operands with strides.
 * kind 0: |a-b|   1: (1-a)^2   2: (1-a)^2 + b^2 ;  out[b] = weight * mean.   */
int mk_loss_fwd(int kind, const float* a, const long long* stride_a, const float* b, const long long* stride_b,
                int B, int C, int D, int H, int W, float weight, float* out, void* stream);
/* da/db written with the operand's own strides (either may be NULL); gout[b] is d loss / d out[b]. */
int mk_loss_bwd(int kind, const float* a, const long long* stride_a, const float* b, const long long* stride_b,
                int B, int C, int D, int H, int W, float weight, const float* gout, float* da, float* db,
                void* stream);

/* ---- multi-GPU: one-shot all-reduce of the packed BN statistics over NVLink peer memory (csrc/p2p.cu), replacing one
 * NCCL call per BN layer per direction (sync_batchnorm/batchnorm.py:90-125).  Every rank allocates
 * mk_stats_allreduce_bytes() bytes of SYMMETRIC memory (zero-filled once); `peers` is a HOST array of `world` device
 * pointers (rank r's buffer as mapped in this process); `seq` a zero-initialised device uint64 owned by the
 * communicator.  Sums `n` doubles (is_double) or floats at `local` in place, in rank order (bit-identical on all
 * ranks); the sequence number lives on the device, so the launch is CUDA-graph capturable. */
int mk_stats_allreduce_bytes(void);
int mk_stats_allreduce(void* local, int n, int is_double, const unsigned long long* peers, int rank, int world,
                       unsigned long long* seq, void* stream);

/* ---- optimiser (train.py:81-83,118-136): fused Adam over one flat fp32 span. */
int mk_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float bias_c1, float bias_c2, void* stream);

/* Adam over ONE flat parameter group, CUDA-graph capturable (train.py:81-83: betas (0.5, 0.999), eps 1e-8, no weight
 * decay; update identical to torch.optim.Adam).  p, g, m, v: flat fp32 buffers of n elements (16-byte aligned);
 * *step (device int64) is the number of updates done so far and is advanced by the kernel; *ticket (device uint32,
 * zero-initialised) is scratch.  zero_grad != 0 clears g in the same pass (optimizer.zero_grad() of train.py:118).
 * lr_dev (device float, may be NULL): when given it overrides `lr`, so a captured CUDA graph follows the
 * MultiStepLR schedule of train.py:92-97,146-148 without re-capture. */
int mk_adam_flat(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 long long* step, unsigned* ticket, int zero_grad, const float* lr_dev, void* stream);

/* ---- data edge (frames_dataset.py:14-29): decoded uint8 image of T frames concatenated
 * horizontally, `img` [H][T*w][Cs] (Cs = 1 gray, 2 gray+alpha, 3 RGB, 4 RGBA; DEVICE pointer) -> `dst` [T][H][w][Cp] fp32
 * NHWC frames in [0,1]: gray replicated to RGB, alpha dropped, value/255 exactly as img_as_float32, channels 3..Cp-1
 * zero.  The uint8 image is what crosses PCIe (4x fewer bytes than the reference's float32 frames). */
int mk_stacked_u8_to_nhwc(const unsigned char* img, int H, int T, int w, int Cs, float* dst, int Cp, void* stream);

#ifdef __cplusplus
}
#endif
#endif
