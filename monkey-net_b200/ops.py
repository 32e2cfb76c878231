"""Autograd plumbing over the C ABI (include/monkey_b200.h).

PyTorch is used here for device memory (torch.empty), the current CUDA stream, the autograd tape and
torch.distributed - nothing else.  Every tensor op on the hot path is a kernel from libmonkey_b200.so; there is no
CPU or ATen fallback (a CPU tensor raises).

Internal activation layout: NHWC fp32 `[N][H][W][Cp]`, N = B*D frames, physical channel count padded to a multiple
of 4 with zero channels.  `Act` carries the tensor plus its channel segments `((logical, padded), ...)` - a concat
of padded tensors has holes, which the weight packer skips via a physical->logical channel map.
"""
import os

import torch

from . import lib
from . import dist as mkdist


def pad4(c):
    return (c + 3) & ~3


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, offset=0):
    if t is None:
        return None
    return t.data_ptr() + 4 * offset


def _check(t, name='tensor'):
    if not t.is_cuda:
        raise RuntimeError('monkey-net_b200: %s is on %s - the B200 path has no CPU fallback' % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError('monkey-net_b200: %s must be float32, got %s' % (name, t.dtype))


def _empty(*shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _zeros(*shape, like):
    t = torch.empty(shape, dtype=torch.float32, device=like.device)
    lib.call('mk_fill_zero', t.data_ptr(), t.numel() * 4, _stream())
    return t


_MAP_CACHE = {}

# Convolution arithmetic (MONKEY_B200_CONV / set_conv_mode):
#   'tf32'   = tcgen05 tensor-core kernels, TF32 operands (10-bit mantissa), fp32 accumulate in TMEM, for the forward,
#              input-gradient and weight-gradient convolutions - the inference default: generator output within the
#              north-star 1e-3 of the fp32 reference (tests/test_gpu_3_tc.py);
#   'tf32x3' = the same kernels in 3xTF32 mode (operands split hi + lo, three MMAs per product): fp32-accurate
#              tensor-core convolutions - the TRAINING default (train-mode batch norm over near-constant channels
#              amplifies 1xTF32 rounding beyond the reference's precision);
#   'fp32'   = exact FFMA kernels (parity 1e-5, tests/test_gpu_1_ops.py).
# 'auto' (default) = 'tf32x3' while autograd is recording, 'tf32' under torch.no_grad().
CONV_MODES = ('fp32', 'tf32', 'tf32x3', 'auto')
CONV_MODE = os.environ.get('MONKEY_B200_CONV', 'auto')


_PRECISE = [0]


class reference_precision:
    """Context: under the 'auto' policy run the enclosed convolutions fp32-accurately (3xTF32) even without autograd.
    The two GEOMETRY networks use it (KP_PRECISE): the keypoint detector - the north-star asks for bit-exact keypoint
    pixel indices, and a 1xTF32 hourglass moves the soft-argmax by ~5e-5, enough to flip a rounded pixel coordinate
    next to a .5 boundary - and the dense-motion network, whose sampling coordinates are amplified by image gradients
    (`video_deformed` off by 2e-3..2e-2 with 1xTF32).  The appearance path meets the 1e-3 frame bar with 1xTF32."""

    def __enter__(self):
        _PRECISE[0] += 1

    def __exit__(self, *exc):
        _PRECISE[0] -= 1
        return False


KP_PRECISE = os.environ.get('MONKEY_B200_KP_PRECISE', '1') != '0'


def conv_mode():
    """the arithmetic the next convolution will run in"""
    if CONV_MODE == 'auto':
        return 'tf32x3' if (torch.is_grad_enabled() or _PRECISE[0] > 0) else 'tf32'
    return CONV_MODE


# Halo-window persistent tensor-core conv (csrc/conv_halo.cu): the default for the many-tile stride-1 layers; the entry
# point declines (-2, before touching device state) shapes outside its envelope and few-tile layers, which run on
# mk_conv2d_tc (split-K).  MONKEY_B200_CONV_HALO=0 switches it off.
CONV_HALO = os.environ.get('MONKEY_B200_CONV_HALO', '1') != '0'
CONV_HALO_UPS = os.environ.get('MONKEY_B200_CONV_HALO_UPS', '1') != '0'   # the upsampled convs on the halo kernel (A/B switch)


def _tc_pack_numel(taps, kout, kin, x3):
    """floats of a tensor-core weight pack [tap][Kout][Kin]; reference precision (mk_pack_weight mode | 8) appends the
    cross operand of the BF16 correction MMA: [tap][Kout][Kin rounded up to 8] 4-byte slots (csrc/conv_halo.cu)"""
    n = taps * kout * kin
    if x3:
        n += taps * kout * ((kin + 7) & ~7)
    return n


def _tc_launch(xp, N, Hin, Win, Cp, ups, wp, R, S, pad, scale, shift, resid_ptr, ldr, act, slope, yp, Cop, st, x3=False):
    """one tensor-core convolution launch on raw pointers: halo kernel when it takes the shape, else the per-tap one"""
    if CONV_HALO and not ups:
        rc = lib.call_soft('mk_conv2d_tc_halo_x3' if x3 else 'mk_conv2d_tc_halo', (-2,), xp, N, Hin, Win, Cp, Cp, wp,
                           R, S, pad, scale, shift, resid_ptr, ldr, act, slope, yp, Cop, Cop, st)
        if rc == 0:
            return
    elif CONV_HALO and CONV_HALO_UPS and resid_ptr is None:
        # conv3x3(nearest_x2(x)): four sub-pixel 2x2 passes of the halo kernel on the low-resolution grid
        rc = lib.call_soft('mk_conv2d_tc_halo_ups_x3' if x3 else 'mk_conv2d_tc_halo_ups', (-2,), xp, N, Hin, Win, Cp, Cp,
                           wp, scale, shift, act, slope, yp, Cop, Cop, st)
        if rc == 0:
            return
    lib.call('mk_conv2d_tc_x3' if x3 else 'mk_conv2d_tc', xp, N, Hin, Win, Cp, Cp, ups, wp, R, S, pad, scale, shift,
             resid_ptr, ldr, act, slope, yp, Cop, Cop, st)


def _conv_tc_call(x, N, Hin, Win, Cp, ups, wpack, R, S, pad, scale, shift, resid_ptr, ldr, act, slope, y, Cop, groups, st,
                  x3=False):
    _tc_launch(x.data_ptr(), N, Hin, Win, Cp, ups, wpack.data_ptr(), R, S, pad, scale, shift, resid_ptr, ldr, act, slope,
               y.data_ptr(), Cop, st, x3)


def set_conv_mode(mode):
    global CONV_MODE
    assert mode in CONV_MODES, mode
    CONV_MODE = mode


def _tc_ok(cin_p, cout_p, ups, pool, k=3, groups=1, mode=None):
    """Forward envelope of mk_conv2d_tc: everything but the fused-pool epilogue; the upsampled conv runs as four
    sub-pixel 2x2 convs (3x3 ungrouped kernels)."""
    if (mode or conv_mode()) == 'fp32' or pool:
        return False
    return (k == 3 and groups == 1) if ups else True


_X4 = {}


def _times4_params(cp, device):
    """norm-apply parameter block [mean, invstd, scale, shift] with scale = 4: average-pool(4*x) == 2x2 sum."""
    key = (cp, device)
    if key not in _X4:
        t = torch.zeros(4, cp, dtype=torch.float32, device=device)
        t[2] = 4.0
        _X4[key] = t.reshape(-1)
    return _X4[key]


def _channel_maps(segs, device):
    """(phys->logical int32 [Cp], logical->phys int32 [C]) device tensors for a segment tuple; None for identity."""
    if all(logical == padded for logical, padded in segs):   # no padding holes anywhere: physical == logical
        return None, None
    key = (segs, device)
    if key not in _MAP_CACHE:
        fwd, inv, base = [], [], 0
        for logical, padded in segs:
            for j in range(padded):
                if j < logical:
                    inv.append(len(fwd))
                    fwd.append(base + j)
                else:
                    fwd.append(-1)
            base += logical
        _MAP_CACHE[key] = (torch.tensor(fwd, dtype=torch.int32, device=device),
                           torch.tensor(inv, dtype=torch.int32, device=device))
    return _MAP_CACHE[key]


class Act:
    """NHWC activation: tensor [N,H,W,Cp] + channel segments."""
    __slots__ = ('t', 'segs')

    def __init__(self, t, segs):
        self.t, self.segs = t, tuple(segs)

    @property
    def C(self):
        return sum(s[0] for s in self.segs)

    @property
    def Cp(self):
        return self.t.shape[3]

    @property
    def shape(self):
        return self.t.shape


# ====================================================================================================== layout edge
class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x5, step):
        _check(x5, 'input')
        B, C, D, H, W = x5.shape
        Cp = pad4(C)
        out = _empty(B * D, H // step, W // step, Cp, like=x5)
        sb, sc, sd, sh, sw = x5.stride()
        lib.call('mk_ncdhw_to_nhwc', x5.data_ptr(), B, C, D, H, W, sb, sc, sd, sh, sw, step, out.data_ptr(), Cp, Cp,
                 _stream())
        ctx.meta = (B, C, D, H, W, step)
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, D, H, W, step = ctx.meta
        g = g.contiguous()
        # gradient returned NHWC-backed (channels-last memory, logical NCDHW view): consumers are stride-aware
        if step == 1:
            return _nhwc_as_ncdhw(g, B, C), None
        dx = _zeros(B, C, D, H, W, like=g)
        sb, sc, sd, sh, sw = dx.stride()
        lib.call('mk_nhwc_to_ncdhw', g.data_ptr(), g.shape[3], B, C, D, H // step, W // step, step, dx.data_ptr(), sb,
                 sc, sd, sh, sw, _stream())
        return dx, None


def _nhwc_as_ncdhw(t, B, C):
    """zero-copy logical (B,C,D,H,W) view of an NHWC tensor [B*D,H,W,Cp]."""
    N, H, W, Cp = t.shape
    return t.view(B, N // B, H, W, Cp).permute(0, 4, 1, 2, 3)[:, :C]


def _is_nhwc_backed(x5):
    """True if x5 (B,C,D,H,W) is exactly the view produced by _nhwc_as_ncdhw of a dense [B*D,H,W,pad4(C)] buffer."""
    B, C, D, H, W = x5.shape
    Cp = pad4(C)
    return x5.stride() == (D * H * W * Cp, 1, H * W * Cp, W * Cp, Cp) and x5.storage_offset() % Cp == 0


def to_nhwc(x5, step=1):
    """(B,C,D,H,W) reference-layout tensor (any strides) -> Act; `step` = nearest down-scale 1/scale_factor."""
    return Act(_ToNHWC.apply(x5, step), ((x5.shape[1], pad4(x5.shape[1])),))


class _FromNHWC(torch.autograd.Function):
    """Expose an NHWC activation as the reference's logical (B,C,D,H,W) tensor without copying."""

    @staticmethod
    def forward(ctx, t, B, C):
        ctx.meta = (B, C, t.shape)
        return _nhwc_as_ncdhw(t, B, C)

    @staticmethod
    def backward(ctx, g5):
        B, C, shape = ctx.meta
        N, H, W, Cp = shape
        D = N // B
        out = _empty(N, H, W, Cp, like=g5)
        sb, sc, sd, sh, sw = g5.stride()
        lib.call('mk_ncdhw_to_nhwc', g5.data_ptr(), B, C, D, H, W, sb, sc, sd, sh, sw, 1, out.data_ptr(), Cp, Cp,
                 _stream())
        return out, None, None


def from_nhwc(a, B):
    assert len(a.segs) == 1
    return _FromNHWC.apply(a.t, B, a.segs[0][0])


# ====================================================================================================== convolution
# Direct parameter gradients: inside `direct_param_grads()` the conv weight gradient is ACCUMULATED by the unpack kernel
# straight into `weight.grad` (when that tensor exists, e.g. FlatAdam's views of its flat gradient buffer, zeroed by the
# fused Adam launch) and autograd receives None for the weight - the same accumulate-into-.grad semantics as
# AccumulateGrad, minus one temporary and one `add_` launch per parameter (62 of the 247 ATen elementwise launches of a
# taichi training iteration).  Only meaningful under `.backward()`; never enabled for `torch.autograd.grad` callers.
_DIRECT_GRAD = [0]


class direct_param_grads:
    def __enter__(self):
        _DIRECT_GRAD[0] += 1

    def __exit__(self, *exc):
        _DIRECT_GRAD[0] -= 1
        return False


def _direct_grad_target(param):
    if _DIRECT_GRAD[0] <= 0:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != param.shape or g.device != param.device:
        return None
    return g


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, resid, segs, pad, groups, ups, act, pool, mode, bias_grad_zero=False):
        # `mode` is resolved by the caller: inside an autograd Function grad mode is always off
        _check(x, 'conv input')
        _check(weight, 'conv weight')
        N, Hin, Win, Cp = x.shape
        Co, Cig, _, R, S = weight.shape
        Cop = pad4(Co)
        cmap, cinv = _channel_maps(segs, x.device)
        st = _stream()
        wpack = _empty(R * S * Cp * Cop, like=x)
        bias_p = _empty(Cop, like=x) if bias is not None else None
        # tensor-core path: pack mode 2 = [tap][Cout_p][Cin_p] (TF32), mode 4 = sub-pixel pack of the upsampled conv
        tc = _tc_ok(Cp, Cop, ups, pool, R, groups, mode)
        x3 = tc and mode == 'tf32x3'
        if tc:
            wpack = _empty(_tc_pack_numel(16 if ups else R * S, Cop, Cp, x3), like=x)
        lib.call('mk_pack_weight', weight.data_ptr(), Co, Cig, R, S, groups, _ptr(cmap), Cp, Cop,
                 ((4 if ups else 2) | (8 if x3 else 0)) if tc else 0, wpack.data_ptr(), _ptr(bias), _ptr(bias_p), st)
        Hl, Wl = Hin << ups, Win << ups
        Ho, Wo = Hl + 2 * pad - R + 1, Wl + 2 * pad - S + 1
        if pool:
            y = _empty(N, Ho >> 1, Wo >> 1, Cop, like=x)
        else:
            y = _empty(N, Ho, Wo, Cop, like=x)
        act_code = {None: 0, 'relu': 1, 'sigmoid': 2}[act]
        if tc:
            _conv_tc_call(x, N, Hin, Win, Cp, ups, wpack, R, S, pad, None, _ptr(bias_p), _ptr(resid),
                          Cop if resid is not None else 0, act_code, 0.0, y, Cop, groups, st, x3=x3)
        else:
            lib.call('mk_conv2d', x.data_ptr(), N, Hin, Win, Cp, Cp, ups, wpack.data_ptr(), R, S, pad, None,
                     _ptr(bias_p), _ptr(resid), Cop if resid is not None else 0, act_code, 0.0, y.data_ptr(), Cop, Cop,
                     pool, st)
        # the backward runs in the arithmetic of its forward (autograd executes it with grad mode off)
        ctx.mode = mode
        ctx.bias_grad_zero = bool(bias_grad_zero)
        ctx.cfg = (segs, pad, groups, ups, act, pool, bias is not None, resid is not None)
        ctx.save_for_backward(x, weight, y if act == 'sigmoid' else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        segs, pad, groups, ups, act, pool, has_bias, has_resid = ctx.cfg
        x, weight, y = ctx.saved_tensors
        if pool or act == 'relu':
            raise RuntimeError('conv with fused relu/pool epilogue is inference-only')
        dy = dy.contiguous()
        N, Hin, Win, Cp = x.shape
        Co, Cig, _, R, S = weight.shape
        Cop = pad4(Co)
        st = _stream()
        if act == 'sigmoid':
            dz = torch.empty_like(dy)
            lib.call('mk_sigmoid_bwd', y.data_ptr(), dy.data_ptr(), dz.data_ptr(), dy.numel(), st)
            dy = dz
        cmap, cinv = _channel_maps(segs, x.device)
        dx = dw = db = None
        tc = ctx.mode != 'fp32'
        x3 = ctx.mode == 'tf32x3'
        if ctx.needs_input_grad[0]:
            wt = _empty(_tc_pack_numel(R * S, Cp, Cop, x3) if tc else R * S * Cop * Cp, like=x)
            lib.call('mk_pack_weight', weight.data_ptr(), Co, Cig, R, S, groups, _ptr(cmap), Cp, Cop,
                     (3 | (8 if x3 else 0)) if tc else 1, wt.data_ptr(), None, None, st)
            dx = _empty(N, Hin, Win, Cp, like=x)
            # dgrad = correlation of dy with the flipped/transposed kernel, padding R-1-pad.  The transpose of the
            # nearest-x2 upsample is a 2x2 sum: fused as the fp32 kernel's pooled epilogue; on the tensor-core path
            # the full-resolution gradient is pooled by the (x4, average) mode of the norm-apply kernel.
            if tc and ups:
                full = _empty(N, dy.shape[1], dy.shape[2], Cp, like=x)
                _tc_launch(dy.data_ptr(), N, dy.shape[1], dy.shape[2], Cop, 0, wt.data_ptr(), R, S, R - 1 - pad, None,
                           None, None, 0, 0, 0.0, full.data_ptr(), Cp, st, x3)
                lib.call('mk_norm_apply', full.data_ptr(), Cp, N, dy.shape[1], dy.shape[2], Cp,
                         _times4_params(Cp, x.device).data_ptr(), 0, -1.0, 1, dx.data_ptr(), Cp, st)
            elif tc:
                _tc_launch(dy.data_ptr(), N, dy.shape[1], dy.shape[2], Cop, 0, wt.data_ptr(), R, S, R - 1 - pad, None,
                           None, None, 0, 0, 0.0, dx.data_ptr(), Cp, st, x3)
            else:
                lib.call('mk_conv2d', dy.data_ptr(), N, dy.shape[1], dy.shape[2], Cop, Cop, 0, wt.data_ptr(), R, S,
                         R - 1 - pad, None, None, None, 0, 0, 0.0, dx.data_ptr(), Cp, Cp, 2 if ups else 0, st)
        ups_done = False
        if (ctx.needs_input_grad[1] and tc and ups and CONV_HALO and CONV_HALO_UPS and R == 3 and S == 3 and pad == 1
                and groups == 1):
            # upsampled conv: the weight gradient on the low-resolution grid (four sub-pixel passes, no upsampled copy
            # of x), folded onto the 3x3 taps by the adjoint of the sub-pixel pack
            dwp16 = _empty(16 * Cp * Cop, like=x)
            rc = lib.call_soft('mk_conv2d_wgrad_halo_ups_x3' if x3 else 'mk_conv2d_wgrad_halo_ups', (-2,), x.data_ptr(), N,
                               Hin, Win, Cp, Cp, dy.data_ptr(), Cop, Cop, dwp16.data_ptr(), st)
            if rc == 0:
                tgt = _direct_grad_target(weight)
                if tgt is None:
                    dw = torch.empty_like(weight)
                lib.call('mk_unpack_wgrad_ups', dwp16.data_ptr(), Co, Cig, _ptr(cinv), Cp, Cop,
                         (tgt if tgt is not None else dw).data_ptr(), 1 if tgt is not None else 0, st)
                ups_done = True
        if ctx.needs_input_grad[1] and not ups_done:
            dwp = _empty(R * S * Cp * Cop, like=x)
            if tc:
                xin = x
                if ups:  # the tensor-core wgrad reads the upsampled activation as a plain tensor (one resize kernel)
                    xin = _empty(N, Hin * 2, Win * 2, Cp, like=x)
                    lib.call('mk_resize_fwd', x.data_ptr(), N, Hin, Win, Cp, Cp, 0, xin.data_ptr(), Hin * 2, Win * 2, Cp,
                             st)
                rc = -2
                if CONV_HALO:
                    rc = lib.call_soft('mk_conv2d_wgrad_halo_x3' if x3 else 'mk_conv2d_wgrad_halo', (-2,), xin.data_ptr(),
                                       N, xin.shape[1], xin.shape[2], Cp, Cp, dy.data_ptr(), Cop, Cop, R, S, pad,
                                       dwp.data_ptr(), st)
                if rc != 0:
                    lib.call('mk_conv2d_wgrad_tc_x3' if x3 else 'mk_conv2d_wgrad_tc', xin.data_ptr(), N, xin.shape[1],
                             xin.shape[2], Cp, Cp, dy.data_ptr(), Cop, Cop, R, S, pad, dwp.data_ptr(), st)
            else:
                lib.call('mk_conv2d_wgrad', x.data_ptr(), N, Hin, Win, Cp, Cp, ups, dy.data_ptr(), Cop, Cop, R, S, pad,
                         dwp.data_ptr(), st)
            tgt = _direct_grad_target(weight)
            if tgt is not None:
                # the optimiser's gradient buffer is the destination: no temporary, no AccumulateGrad add kernel
                lib.call('mk_unpack_wgrad_acc', dwp.data_ptr(), Co, Cig, R, S, groups, _ptr(cinv), Cp, Cop,
                         tgt.data_ptr(), st)
            else:
                dw = torch.empty_like(weight)
                lib.call('mk_unpack_wgrad', dwp.data_ptr(), Co, Cig, R, S, groups, _ptr(cinv), Cp, Cop, dw.data_ptr(),
                         st)
        if has_bias and ctx.needs_input_grad[2] and not ctx.bias_grad_zero:
            sums = _empty(2 * Cop, like=x)
            lib.call('mk_colstats', dy.data_ptr(), Cop, N, dy.shape[1] * dy.shape[2], Cop, 0, sums.data_ptr(), st)
            db = sums[:Co]
        dres = dy if has_resid and ctx.needs_input_grad[3] else None
        return dx, dw, db, dres, None, None, None, None, None, None, None, None


def conv(a, weight, bias, pad, groups=1, ups=False, resid=None, act=None, pool=0, feeds_train_norm=False):
    """nn.Conv3d (1,k,k) as a per-frame 2-D conv; returns a single-segment Act.
    `feeds_train_norm`: the output goes straight into a training-mode batch / instance norm.  The norm subtracts the
    per-channel mean, so d(loss)/d(bias) = sum over pixels of the norm's input gradient is EXACTLY zero; the reference
    computes it anyway and gets rounding noise of order 1e-9 (tests/helpers.py:structurally_zero_grad).  We return the
    exact value - no gradient - and save one reduction pass over dy per such layer (74 launches at taichi@256)."""
    if INFER_FUSION and not torch.is_grad_enabled():
        return conv_infer(a, weight, bias, pad, groups=groups, ups=ups, resid=resid,
                          act={None: 0, 'relu': 1, 'sigmoid': 2}[act], pool=pool)
    y = _Conv.apply(a.t, weight, bias, resid.t if resid is not None else None, a.segs, pad, groups, int(ups), act,
                    pool, conv_mode(), bool(feeds_train_norm))
    co = weight.shape[0]
    return Act(y, ((co, pad4(co)),))


# ---------------------------------------------------------------------------------------------- inference fast path
# Under torch.no_grad() (transfer.py / reconstruction.py run the nets in eval mode without autograd) nothing has to
# be saved for backward and the parameters do not change between calls:
#   * the packed GEMM weights and the epilogue vectors are CACHED per parameter version instead of re-packed by a
#     kernel on every call;
#   * conv -> eval-BatchNorm -> ReLU is ONE launch: the BN of running statistics folds into the conv epilogue,
#     scale = gamma * invstd, shift = beta + (bias - running_mean) * scale (batchnorm.py:50-53 with training=False),
#     which removes a full read + write pass per block.
# The tensor-core kernel splits K only for a linear epilogue, so layers whose tile count cannot fill the machine
# (deep 2x2 ... 8x8 levels) keep conv (split-K) + norm_apply.
INFER_FUSION = os.environ.get('MONKEY_B200_INFER_FUSION', '1') != '0'
_INFER_CACHE = {}
# Kernels update parameters and BN running statistics through raw pointers (mk_norm_finalize, mk_adam_flat, CUDA-graph
# replays of a training step), which torch's per-tensor version counters do not see: every such update bumps this
# epoch and thereby invalidates the cached packs.
_PARAM_EPOCH = [0]


def note_parameters_changed():
    _PARAM_EPOCH[0] += 1


def _versions(*ts):
    return (_PARAM_EPOCH[0],) + tuple((t.data_ptr(), t._version) if t is not None else None for t in ts)


def _infer_pack(weight, bias, segs, groups, ups, norm, Cp, tc, x3=False):
    """(packed weights, scale or None, shift or None) for this conv (+ folded eval BN), cached on parameter versions."""
    import weakref
    key = (id(weight), segs, int(ups), bool(tc), bool(x3), id(norm) if norm is not None else None)
    nt = (norm.weight, norm.bias, norm.running_mean, norm.running_var) if norm is not None else ()
    ver = _versions(weight, bias, *nt)
    hit = _INFER_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[1]() is weight:
        return hit[2]
    Co, Cig, _, R, S = weight.shape
    Cop = pad4(Co)
    st = _stream()
    cmap, _ = _channel_maps(segs, weight.device)
    wpack = _empty(_tc_pack_numel(16 if ups else R * S, Cop, Cp, x3) if tc else R * S * Cp * Cop, like=weight)
    bias_p = _empty(Cop, like=weight) if bias is not None else None
    lib.call('mk_pack_weight', weight.data_ptr(), Co, Cig, R, S, groups, _ptr(cmap), Cp, Cop,
             ((4 if ups else 2) | (8 if x3 else 0)) if tc else 0, wpack.data_ptr(), _ptr(bias), _ptr(bias_p), st)
    scale = None
    shift = bias_p
    if norm is not None:
        with torch.no_grad():
            sc = norm.weight.detach() * torch.rsqrt(norm.running_var.detach().double() + 1e-5).float()
            b0 = bias.detach() if bias is not None else torch.zeros_like(sc)
            sh = norm.bias.detach() + (b0 - norm.running_mean.detach()) * sc
            scale = torch.zeros(Cop, dtype=torch.float32, device=weight.device)
            shift = torch.zeros(Cop, dtype=torch.float32, device=weight.device)
            scale[:Co] = sc
            shift[:Co] = sh
    val = (wpack, scale, shift)
    _INFER_CACHE[key] = (ver, weakref.ref(weight), val)
    return val


_SPLIT_CACHE = {}


def _tc_would_split(N, Hin, Win, Cp, ups, R, S, pad, Cop):
    """would mk_conv2d_tc run this layer split-K with a linear epilogue?  Asked from the library's own planner
    (mk_conv2d_tc_plan: a host-side dry run, no launch) instead of re-deriving its rule here."""
    key = (N, Hin, Win, Cp, ups, R, S, pad, Cop)
    if key not in _SPLIT_CACHE:
        import ctypes
        out = (ctypes.c_int * 16)()
        lib.query('mk_conv2d_tc_plan', N, Hin, Win, Cp, ups, R, S, pad, 0, Cop, Cop, out)
        _SPLIT_CACHE[key] = out[5] > 1
    return _SPLIT_CACHE[key]


def conv_infer(a, weight, bias, pad, groups=1, ups=False, resid=None, act=0, slope=0.0, pool=0, norm=None):
    """Inference-only convolution (no autograd): y = pool(act((conv(x) + bias) [folded eval-BN of `norm`] + resid))."""
    x = a.t
    _check(x, 'conv input')
    N, Hin, Win, Cp = x.shape
    Co, Cig, _, R, S = weight.shape
    Cop = pad4(Co)
    ups = int(bool(ups))
    tc = _tc_ok(Cp, Cop, ups, 0, R, groups)
    x3 = tc and conv_mode() == 'tf32x3'
    wpack, scale, shift = _infer_pack(weight, bias, a.segs, groups, ups, norm, Cp, tc, x3)
    Hl, Wl = Hin << ups, Win << ups
    Ho, Wo = Hl + 2 * pad - R + 1, Wl + 2 * pad - S + 1
    st = _stream()
    rp, ldr = (resid.t.data_ptr(), Cop) if resid is not None else (None, 0)
    if tc:
        y = _empty(N, Ho, Wo, Cop, like=x)
        _conv_tc_call(x, N, Hin, Win, Cp, ups, wpack, R, S, pad, _ptr(scale), _ptr(shift), rp, ldr, act, float(slope),
                      y, Cop, groups, st, x3=x3)
        out = Act(y, ((Co, Cop),))
        if pool:
            out = norm_act(out, None, mode='none', pool=1)
        return out
    y = _empty(N, Ho >> 1, Wo >> 1, Cop, like=x) if pool else _empty(N, Ho, Wo, Cop, like=x)
    lib.call('mk_conv2d', x.data_ptr(), N, Hin, Win, Cp, Cp, ups, wpack.data_ptr(), R, S, pad, _ptr(scale), _ptr(shift),
             rp, ldr, act, float(slope), y.data_ptr(), Cop, Cop, int(pool), st)
    return Act(y, ((Co, Cop),))


def conv_bn_relu(a, conv_mod, norm_mod, pad, groups=1, ups=False, pool=0, extras=()):
    """conv -> BatchNorm -> ReLU [-> 2x2 average pool] [-> concat(extras)]: the body of DownBlock3D / UpBlock3D /
    SameBlock3D (util.py:91-126).  Training or autograd: conv + the two-phase norm kernels.  Inference (eval-mode BN,
    no autograd): one fused conv launch where the layer has enough tiles, see above."""
    infer = INFER_FUSION and not torch.is_grad_enabled() and not norm_mod.training
    if infer:
        N, Hin, Win, Cp = a.t.shape
        Co, _, _, R, S = conv_mod.weight.shape
        u = int(bool(ups))
        tc = _tc_ok(Cp, pad4(Co), u, 0, R, groups)
        if not (tc and _tc_would_split(N, Hin, Win, Cp, u, R, S, pad, pad4(Co))):
            y = conv_infer(a, conv_mod.weight, conv_mod.bias, pad, groups=groups, ups=ups, act=1, slope=0.0,
                           pool=pool, norm=norm_mod)
            return norm_act(y, None, mode='none', extras=extras) if extras else y
    y = conv(a, conv_mod.weight, conv_mod.bias, pad=pad, groups=groups, ups=ups, feeds_train_norm=bool(norm_mod.training))
    return norm_act(y, norm_mod, mode='bn', slope=0.0, pool=pool, extras=extras)


# ====================================================================================================== norm + act
class _NormAct(torch.autograd.Function):
    """out = concat(pool(act(norm(x))), *extras).  mode: 'bn' (batch norm; train or eval), 'in', 'none'."""

    @staticmethod
    def forward(ctx, x, gamma, beta, cfg, *extras):
        mode, training, slope, pool, C, module = cfg
        _check(x, 'norm input')
        N, H, W, Cp = x.shape
        st = _stream()
        params = None
        per_frame = 1 if mode == 'in' else 0
        count = float(N * H * W)
        if mode == 'bn' and training:
            sums = torch.empty(2 * Cp, dtype=torch.float64, device=x.device)  # double: E[x^2]-E[x]^2 cancels in fp32
            lib.call('mk_colstats_f64', x.data_ptr(), Cp, N, H * W, Cp, 0, sums.data_ptr(), st)
            count *= mkdist.all_reduce_stats(sums)
            params = _empty(4 * Cp, like=x)
            lib.call('mk_norm_finalize', sums.data_ptr(), 1, C, Cp, count, _ptr(gamma), _ptr(beta), 1e-5,
                     module.running_mean.data_ptr(), module.running_var.data_ptr(), 0.1,
                     module.num_batches_tracked.data_ptr(), params.data_ptr(), st)
            note_parameters_changed()  # running statistics written through raw pointers
        elif mode == 'bn':
            params = _empty(4 * Cp, like=x)
            lib.call('mk_norm_eval_params', module.running_mean.data_ptr(), module.running_var.data_ptr(), _ptr(gamma),
                     _ptr(beta), C, Cp, 1e-5, params.data_ptr(), st)
        elif mode == 'in':
            sums = torch.empty(N * 2 * Cp, dtype=torch.float64, device=x.device)
            lib.call('mk_colstats_f64', x.data_ptr(), Cp, N, H * W, Cp, 1, sums.data_ptr(), st)
            count = float(H * W)
            params = _empty(N * 4 * Cp, like=x)
            lib.call('mk_norm_finalize', sums.data_ptr(), N, C, Cp, count, _ptr(gamma), _ptr(beta), 1e-5, None, None,
                     0.0, None, params.data_ptr(), st)
        Ho, Wo = (H >> 1, W >> 1) if pool else (H, W)
        ctot = Cp + sum(e.shape[3] for e in extras)
        out = _empty(N, Ho, Wo, ctot, like=x)
        lib.call('mk_norm_apply', x.data_ptr(), Cp, N, H, W, Cp, _ptr(params), per_frame, slope, pool, out.data_ptr(),
                 ctot, st)
        off = Cp
        for e in extras:
            lib.call('mk_copy_channels', e.data_ptr(), e.shape[3], _ptr(out, off), ctot, N * Ho * Wo, e.shape[3], st)
            off += e.shape[3]
        ctx.cfg = (mode, training, slope, pool, C, count, [e.shape[3] for e in extras])
        ctx.save_for_backward(x, params)
        return out

    @staticmethod
    def backward(ctx, dout):
        mode, training, slope, pool, C, count, ext_c = ctx.cfg
        x, params = ctx.saved_tensors
        dout = dout.contiguous()
        N, H, W, Cp = x.shape
        ctot = dout.shape[3]
        st = _stream()
        npix_out = dout.shape[0] * dout.shape[1] * dout.shape[2]
        normed = (mode == 'bn' and training) or mode == 'in'
        per_frame = 1 if mode == 'in' else 0
        dgamma = dbeta = None
        sums = None
        if normed:
            groups = N if per_frame else 1
            sums = _empty(groups * 2 * Cp, like=x)
            lib.call('mk_norm_bwd_reduce', x.data_ptr(), Cp, dout.data_ptr(), ctot, N, H, W, Cp, params.data_ptr(),
                     per_frame, slope, pool, sums.data_ptr(), st)
            if per_frame:
                tot = _empty(2 * 2 * Cp, like=x)
                lib.call('mk_colstats', sums.data_ptr(), 2 * Cp, N, 1, 2 * Cp, 0, tot.data_ptr(), st)
                dbeta, dgamma = tot[:C], tot[Cp:Cp + C]
            else:
                # parameter gradients are the LOCAL sums (the gradient all-reduce averages them later); the input
                # gradient needs the GLOBAL sums -> one all-reduce of [2][Cp] per layer in backward as well
                if mkdist.stats_world() > 1:
                    local = sums.clone()
                    mkdist.all_reduce_stats(sums)
                else:
                    local = sums
                dbeta, dgamma = local[:C], local[Cp:Cp + C]
        elif mode == 'bn' and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            # eval-mode BN with trainable affine parameters (fine-tuning with frozen statistics): F.batch_norm(
            # training=False) still yields dgamma = sum dz*xhat, dbeta = sum dz with xhat from the RUNNING statistics -
            # exactly what the reduce kernel computes from the eval parameter block.  dx needs no sums in this mode.
            esums = _empty(2 * Cp, like=x)
            lib.call('mk_norm_bwd_reduce', x.data_ptr(), Cp, dout.data_ptr(), ctot, N, H, W, Cp, params.data_ptr(),
                     0, slope, pool, esums.data_ptr(), st)
            dbeta, dgamma = esums[:C], esums[Cp:Cp + C]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _empty(N, H, W, Cp, like=x)
            lib.call('mk_norm_bwd_apply', x.data_ptr(), Cp, dout.data_ptr(), ctot, N, H, W, Cp, _ptr(params),
                     _ptr(sums), count, per_frame, 1 if normed else 0, slope, pool, dx.data_ptr(), Cp, st)
        dextras = []
        off = Cp
        for i, c in enumerate(ext_c):
            if ctx.needs_input_grad[4 + i]:
                g = _empty(dout.shape[0], dout.shape[1], dout.shape[2], c, like=x)
                lib.call('mk_copy_channels', _ptr(dout, off), ctot, g.data_ptr(), c, npix_out, c, st)
                dextras.append(g)
            else:
                dextras.append(None)
            off += c
        if not ctx.needs_input_grad[1]:
            dgamma = None
        if not ctx.needs_input_grad[2]:
            dbeta = None
        return (dx, dgamma, dbeta, None) + tuple(dextras)


def norm_act(a, norm=None, mode='none', slope=-1.0, pool=0, extras=()):
    """`norm` is the parameter-holding module (weight, bias, running stats); extras are Acts concatenated after."""
    assert len(a.segs) == 1
    C = a.segs[0][0]
    gamma = norm.weight if norm is not None else None
    beta = norm.bias if norm is not None else None
    training = bool(norm.training) if norm is not None else False
    cfg = (mode, training, float(slope), int(pool), C, norm if mode == 'bn' else None)
    out = _NormAct.apply(a.t, gamma, beta, cfg, *[e.t for e in extras])
    segs = a.segs
    for e in extras:
        segs = segs + e.segs
    return Act(out, segs)


class _Concat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *ts):
        N, H, W = ts[0].shape[:3]
        ctot = sum(t.shape[3] for t in ts)
        out = _empty(N, H, W, ctot, like=ts[0])
        off = 0
        for t in ts:
            lib.call('mk_copy_channels', t.data_ptr(), t.shape[3], _ptr(out, off), ctot, N * H * W, t.shape[3],
                     _stream())
            off += t.shape[3]
        ctx.cs = [t.shape[3] for t in ts]
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N, H, W, ctot = g.shape
        outs, off = [], 0
        for i, c in enumerate(ctx.cs):
            if ctx.needs_input_grad[i]:
                t = _empty(N, H, W, c, like=g)
                lib.call('mk_copy_channels', _ptr(g, off), ctot, t.data_ptr(), c, N * H * W, c, _stream())
                outs.append(t)
            else:
                outs.append(None)
            off += c
        return tuple(outs)


def concat(acts):
    segs = ()
    for a in acts:
        segs = segs + a.segs
    return Act(_Concat.apply(*[a.t for a in acts]), segs)


class _Compact(torch.autograd.Function):
    """Gather the logical channels of a concat-with-holes tensor into one dense zero-padded segment."""

    @staticmethod
    def forward(ctx, x, segs):
        N, H, W, Cp = x.shape
        cmap, cinv = _channel_maps(segs, x.device)
        C = cinv.numel()
        Cd = pad4(C)
        key = ('compact', segs, x.device)
        if key not in _MAP_CACHE:
            src_of = torch.full((Cd,), -1, dtype=torch.int32)
            src_of[:C] = cinv.cpu()
            _MAP_CACHE[key] = src_of.to(x.device)
        out = _empty(N, H, W, Cd, like=x)
        lib.call('mk_gather_channels', x.data_ptr(), Cp, _MAP_CACHE[key].data_ptr(), out.data_ptr(), Cd, N * H * W, Cd,
                 _stream())
        ctx.meta = (segs, Cp)
        return out

    @staticmethod
    def backward(ctx, g):
        segs, Cp = ctx.meta
        g = g.contiguous()
        N, H, W, Cd = g.shape
        cmap, _ = _channel_maps(segs, g.device)  # phys -> logical (== position in the compact tensor) or -1
        dx = _empty(N, H, W, Cp, like=g)
        lib.call('mk_gather_channels', g.data_ptr(), Cd, cmap.data_ptr(), dx.data_ptr(), Cp, N * H * W, Cp, _stream())
        return dx, None


def compact(a):
    if len(a.segs) == 1:
        return a
    c = a.C
    return Act(_Compact.apply(a.t, a.segs), ((c, pad4(c)),))


# ====================================================================================================== sampling
class _GridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, deform, d, mode):
        _check(inp, 'grid_sample input')
        B, h, w, Cp = inp.shape
        N, h0, w0, _ = deform.shape
        assert N == B * d
        out = _empty(N, h, w, Cp, like=inp)
        lib.call('mk_grid_sample_fwd', inp.data_ptr(), B, h, w, Cp, Cp, deform.data_ptr(), d, h0, w0, mode,
                 out.data_ptr(), Cp, _stream())
        ctx.meta = (d, mode)
        ctx.save_for_backward(inp, deform)
        return out

    @staticmethod
    def backward(ctx, g):
        d, mode = ctx.meta
        inp, deform = ctx.saved_tensors
        g = g.contiguous()
        B, h, w, Cp = inp.shape
        N, h0, w0, _ = deform.shape
        dinp = _zeros(B, h, w, Cp, like=inp) if ctx.needs_input_grad[0] else None
        ddef = _zeros(N, h0, w0, 2, like=inp) if ctx.needs_input_grad[1] else None
        lib.call('mk_grid_sample_bwd', inp.data_ptr(), B, h, w, Cp, Cp, deform.data_ptr(), d, h0, w0, mode,
                 g.data_ptr(), Cp, _ptr(dinp), Cp, _ptr(ddef), _stream())
        return dinp, ddef, None, None


def grid_sample(a, deform, d, mode):
    """a: Act [B,h,w,Cp]; deform [B*d,h0,w0,2]; mode 'nearest' | 'trilinear' (grid resize rule)."""
    return Act(_GridSample.apply(a.t, deform, d, 0 if mode == 'nearest' else 1), a.segs)


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w, mode):
        N, h0, w0, Cp = x.shape
        out = _empty(N, h, w, Cp, like=x)
        lib.call('mk_resize_fwd', x.data_ptr(), N, h0, w0, Cp, Cp, mode, out.data_ptr(), h, w, Cp, _stream())
        ctx.meta = (h0, w0, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        h0, w0, mode = ctx.meta
        g = g.contiguous()
        N, h, w, Cp = g.shape
        dx = _zeros(N, h0, w0, Cp, like=g)
        lib.call('mk_resize_bwd', g.data_ptr(), N, h, w, Cp, Cp, mode, dx.data_ptr(), h0, w0, Cp, _stream())
        return dx, None, None, None


def resize(a, h, w, mode):
    if a.t.shape[1] == h and a.t.shape[2] == w:
        return a
    return Act(_Resize.apply(a.t, h, w, 0 if mode == 'nearest' else 1), a.segs)


# ====================================================================================================== keypoints
def _kp_scratch(N, H, W, K, like):
    """chunk partials of the many-block keypoint-head kernels (size asked from the library; a host-side query, not a
    launch)"""
    import ctypes
    n = ctypes.c_longlong(0)
    lib.query('mk_kp_head_scratch_floats', N, H, W, K, ctypes.byref(n))
    return _empty(int(n.value), like=like)


class _KPHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, B, D, K, inv_t, var_mode, clip):
        _check(logits, 'kp logits')
        N, H, W, ld = logits.shape
        mean = _empty(B, D, K, 2, like=logits)
        var = _empty(B, D, K, 2, 2, like=logits) if var_mode == 0 else _empty(B, D, K, 1, 1, like=logits)
        aux = _empty(N, K, 8, like=logits)
        lib.call('mk_kp_head_fwd', logits.data_ptr(), N, H, W, K, ld, inv_t, var_mode, clip, mean.data_ptr(),
                 var.data_ptr(), aux.data_ptr(), _kp_scratch(N, H, W, K, logits).data_ptr(), _stream())
        ctx.meta = (K, inv_t, var_mode, clip)
        ctx.save_for_backward(logits, mean, aux)
        return mean, var

    @staticmethod
    def backward(ctx, dmean, dvar):
        K, inv_t, var_mode, clip = ctx.meta
        logits, mean, aux = ctx.saved_tensors
        N, H, W, ld = logits.shape
        dmean = dmean.contiguous() if dmean is not None else _zeros(*mean.shape, like=logits)
        if dvar is not None:
            dvar = dvar.contiguous()
        else:
            dvar = _zeros(N * K * (4 if var_mode == 0 else 1), like=logits)
        dl = _empty(N, H, W, ld, like=logits)
        lib.call('mk_kp_head_bwd', logits.data_ptr(), N, H, W, K, ld, inv_t, var_mode, clip, mean.data_ptr(),
                 aux.data_ptr(), dmean.data_ptr(), dvar.data_ptr(), dl.data_ptr(),
                 _kp_scratch(N, H, W, K, logits).data_ptr(), _stream())
        return dl, None, None, None, None, None, None


def kp_head(a, B, D, K, temperature, kp_variance, clip_variance):
    var_mode = {'matrix': 0, 'single': 1}.get(kp_variance, 0)
    mean, var = _KPHead.apply(a.t, B, D, K, 1.0 / temperature, var_mode, float(clip_variance or 0.0))
    out = {'mean': mean}
    if kp_variance in ('matrix', 'single'):
        out['var'] = var
    return out


def _kp_c(t):
    return t if t.is_contiguous() else t.contiguous()


class _Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, kd_mean, kd_var, ks_mean, ks_var, cfg):
        flags, var_mode, const_var, norm_const, C, h, w = cfg
        _check(kd_mean, 'kp mean')
        B, d, K, _ = kd_mean.shape
        st = _stream()
        slots = K + (1 if flags & 8 else 0)
        F = (1 if flags & 1 else 0) + (2 if flags & 2 else 0) + (C if flags & 4 else 0)
        Cout_p = pad4(slots * F)
        heat_sums = None
        if norm_const == 0.0 and (flags & 1):
            heat_sums = _empty(2 * B * d * K, like=kd_mean)
            lib.call('mk_kp_heat_sums', kd_mean.data_ptr(), _ptr(kd_var), ks_mean.data_ptr(), _ptr(ks_var), B, d, K, h,
                     w, var_mode, const_var, heat_sums.data_ptr(), st)
        out = _empty(B * d, h, w, Cout_p, like=kd_mean)
        lib.call('mk_movement_embed_fwd', _ptr(src), src.shape[3] if src is not None else 0, C, kd_mean.data_ptr(),
                 _ptr(kd_var), ks_mean.data_ptr(), _ptr(ks_var), B, d, K, h, w, flags, var_mode, const_var, norm_const,
                 _ptr(heat_sums), out.data_ptr(), Cout_p, Cout_p, st)
        ctx.cfg = cfg
        ctx.save_for_backward(src, kd_mean, kd_var, ks_mean, ks_var, heat_sums)
        return out

    @staticmethod
    def backward(ctx, g):
        flags, var_mode, const_var, norm_const, C, h, w = ctx.cfg
        src, kd_mean, kd_var, ks_mean, ks_var, heat_sums = ctx.saved_tensors
        g = g.contiguous()
        B, d, K, _ = kd_mean.shape
        d_kd_mean = _zeros(*kd_mean.shape, like=g)
        d_ks_mean = _zeros(*ks_mean.shape, like=g)
        d_kd_var = _zeros(*kd_var.shape, like=g) if kd_var is not None else None
        d_ks_var = _zeros(*ks_var.shape, like=g) if ks_var is not None else None
        lib.call('mk_movement_embed_bwd', _ptr(src), src.shape[3] if src is not None else 0, C, kd_mean.data_ptr(),
                 _ptr(kd_var), ks_mean.data_ptr(), _ptr(ks_var), B, d, K, h, w, flags, var_mode, const_var, norm_const,
                 _ptr(heat_sums), g.data_ptr(), g.shape[3], d_kd_mean.data_ptr(), _ptr(d_kd_var),
                 d_ks_mean.data_ptr(), _ptr(d_ks_var), _stream())
        return None, d_kd_mean, d_kd_var, d_ks_mean, d_ks_var, None


def _require_single_source_frame(ks_mean):
    """The kernels index kp_source as (B,1,K,..) - every caller of the reference passes the d=1 source keypoints
    (train.py:14-21 `v[:, :1]`, transfer.py:67, reconstruction.py:16).  The reference would broadcast a d>1 source
    against the driving frames; that never-used case is rejected instead of silently reading the wrong keypoints."""
    if ks_mean.dim() != 4 or ks_mean.shape[1] != 1:
        raise NotImplementedError('monkey-net_b200: kp_source must hold exactly one frame (B,1,K,2), got %s'
                                  % (tuple(ks_mean.shape),))


def movement_embed(src, kp_driving, kp_source, h, w, num_channels, kp_variance, use_heatmap, use_difference,
                   use_deformed, add_bg, heatmap_type, norm_const):
    """movement_embedding.py:42-92 -> Act [B*d,h,w,pad4(slots*F)], channels slot-major / feature-minor."""
    flags = (1 if use_heatmap else 0) | (2 if use_difference else 0) | (4 if use_deformed else 0) | \
            (8 if add_bg else 0) | (16 if heatmap_type == 'difference' else 0)
    if kp_variance == 'matrix':
        var_mode, const_var = 0, 0.0
    elif kp_variance == 'single':
        var_mode, const_var = 1, 0.0
    else:
        var_mode, const_var = 2, float(kp_variance)
    nc = 0.0 if norm_const == 'sum' else float(norm_const)
    kd_mean, ks_mean = _kp_c(kp_driving['mean']), _kp_c(kp_source['mean'])
    _require_single_source_frame(ks_mean)
    kd_var = _kp_c(kp_driving['var']) if var_mode != 2 and use_heatmap else None
    ks_var = _kp_c(kp_source['var']) if var_mode != 2 and use_heatmap else None
    cfg = (flags, var_mode, const_var, nc, num_channels, h, w)
    out = _Embed.apply(src.t if (src is not None and use_deformed) else None, kd_mean, kd_var, ks_mean, ks_var, cfg)
    K = kd_mean.shape[2]
    slots = K + (1 if add_bg else 0)
    F = (1 if use_heatmap else 0) + (2 if use_difference else 0) + (num_channels if use_deformed else 0)
    return Act(out, ((slots * F, pad4(slots * F)),))


class _FlowHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, kd_mean, ks_mean, use_mask, use_corr):
        B, d, K, _ = kd_mean.shape
        N, h, w, ld = pred.shape
        deform = _empty(N, h, w, 2, like=pred)
        lib.call('mk_flow_head_fwd', pred.data_ptr(), ld, kd_mean.data_ptr(), ks_mean.data_ptr(), B, d, K, h, w,
                 use_mask, use_corr, deform.data_ptr(), _stream())
        ctx.meta = (use_mask, use_corr)
        ctx.save_for_backward(pred, kd_mean, ks_mean)
        return deform

    @staticmethod
    def backward(ctx, g):
        use_mask, use_corr = ctx.meta
        pred, kd_mean, ks_mean = ctx.saved_tensors
        g = g.contiguous()
        B, d, K, _ = kd_mean.shape
        N, h, w, ld = pred.shape
        dpred = _empty(N, h, w, ld, like=pred)
        dkd = _zeros(*kd_mean.shape, like=pred)
        dks = _zeros(*ks_mean.shape, like=pred)
        lib.call('mk_flow_head_bwd', pred.data_ptr(), ld, kd_mean.data_ptr(), ks_mean.data_ptr(), B, d, K, h, w,
                 use_mask, use_corr, g.data_ptr(), dpred.data_ptr(), dkd.data_ptr(), dks.data_ptr(), _stream())
        return dpred, dkd, dks, None, None


def flow_head(a, kp_driving, kp_source, use_mask, use_correction):
    """dense_motion_module.py:52-76 -> deformation [B*d,h,w,2] (the zero z column is added at the API edge)."""
    ks_mean = _kp_c(kp_source['mean'])
    _require_single_source_frame(ks_mean)
    return _FlowHead.apply(a.t, _kp_c(kp_driving['mean']), ks_mean, int(bool(use_mask)), int(bool(use_correction)))


# ====================================================================================================== losses
def _strides5(t):
    import ctypes
    return (ctypes.c_longlong * 5)(*t.stride())


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, weight, a, b):
        _check(a, 'loss operand')
        B, C, D, H, W = a.shape
        out = _empty(B, like=a)
        sa = _strides5(a)
        sb = _strides5(b) if b is not None else None
        lib.call('mk_loss_fwd', kind, a.data_ptr(), sa, _ptr(b), sb, B, C, D, H, W, weight, out.data_ptr(), _stream())
        ctx.meta = (kind, weight)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        kind, weight = ctx.meta
        a, b = ctx.saved_tensors
        B, C, D, H, W = a.shape
        g = g.contiguous()

        def prep(t):
            """(operand with kernel-friendly strides, gradient buffer sharing exactly those strides)"""
            if _is_nhwc_backed(t):  # discriminator maps / generated frames: keep the channels-last memory
                cp = pad4(C)
                buf = _zeros(B * D, H, W, cp, like=t) if cp != C else _empty(B * D, H, W, cp, like=t)
                return t, _nhwc_as_ncdhw(buf, B, C)
            if not t.is_contiguous():
                t = t.contiguous()
            return t, torch.empty_like(t)

        da = db = None
        if ctx.needs_input_grad[2]:
            a, da = prep(a)
        if b is not None and ctx.needs_input_grad[3]:
            b, db = prep(b)
        sa = _strides5(a)
        sb = _strides5(b) if b is not None else None
        lib.call('mk_loss_bwd', kind, a.data_ptr(), sa, _ptr(b), sb, B, C, D, H, W, weight, g.data_ptr(), _ptr(da),
                 _ptr(db), _stream())
        return None, None, da, db


def loss_mean(kind, a, b, weight):
    """kind 'l1' |a-b|, 'gen_gan' (1-a)^2, 'disc_gan' (1-a)^2 + b^2 ; returns weight * per-sample mean, shape (B,)."""
    k = {'l1': 0, 'gen_gan': 1, 'disc_gan': 2}[kind]
    return _Loss.apply(k, float(weight), a, b)
