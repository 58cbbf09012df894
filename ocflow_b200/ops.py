"""torch.autograd wrappers over the C ABI (include/ocflow_b200.h).  CUDA fp32 tensors only.

PyTorch is plumbing here: it owns device memory and streams and records the autograd graph; every
forward and backward below is one call into libocflow_b200.so on torch's current CUDA stream.
There is no CPU path: a non-CUDA or non-fp32 tensor raises TypeError.
"""
import ctypes
import os

import torch

from . import _lib

NORM_NORMALIZE, NORM_CENTER, NORM_ACROSS_CHANNELS, NORM_ACROSS_IMAGES = 1, 2, 4, 8
WARP_ALIGN_CORNERS, WARP_IS_MASK = 1, 2


def _req(t, name, dims=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor: ocflow_b200 has no CPU path (got device %s)" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    if dims is not None and t.dim() != dims:
        raise ValueError("%s must have %d dimensions (got shape %s)" % (name, dims, tuple(t.shape)))
    return t.contiguous()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# -------------------------------------------------------------------------------------------------
# cost volume
# -------------------------------------------------------------------------------------------------
class _CostVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, d, slope):
        B, C, H, W = f1.shape
        nd = 2 * d + 1
        out = torch.empty((B, nd * nd, H, W), device=f1.device, dtype=torch.float32)
        # fused LeakyReLU: the d = 4 kernels also emit the sign of every cost-volume element as a bitmask (8 pixels per
        # byte); the backward needs nothing else, so the 81-plane output is NOT kept alive by autograd (1/32 of the bytes)
        mask = None
        if slope != 1.0 and d == 4:
            mask = torch.empty((B, nd * nd, H, (W + 7) // 8), device=f1.device, dtype=torch.uint8)
        with torch.cuda.device_of(f1):
            _lib.call("ocf_corr_fwd", _p(f1), _p(f2), _p(out), B, C, H, W, d, 0, float(slope), None, _p(mask), _stream())
        ctx.d, ctx.slope, ctx.has_mask = d, float(slope), mask is not None
        if mask is not None:
            ctx.save_for_backward(f1, f2, mask)
        elif slope != 1.0:
            ctx.save_for_backward(f1, f2, out)
        else:
            ctx.save_for_backward(f1, f2)
        return out

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        f1, f2 = saved[0], saved[1]
        act = saved[2] if len(saved) == 3 and not ctx.has_mask else None
        mask = saved[2] if ctx.has_mask else None
        B, C, H, W = f1.shape
        # The incoming gradient is usually a channel slice of the decoder's concat gradient (CatBackward hands out
        # `narrow` views): dense per batch item, only the batch stride differs.  The C ABI takes that stride, so the
        # slice is consumed in place instead of being re-packed by a 2 x 81 x H x W copy per level.
        K = g.shape[1]
        g_bstride = 0
        if not g.is_contiguous():
            st = g.stride()
            if (st[3] == 1 and st[2] == W and st[1] == H * W and st[0] >= K * H * W and st[0] % 4 == 0
                    and g.data_ptr() % 16 == 0 and g.dtype == torch.float32):
                g_bstride = st[0]
            else:
                g = g.contiguous()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        df1 = torch.empty_like(f1) if need1 else None
        df2 = torch.empty_like(f2) if need2 else None
        if need1 or need2:
            with torch.cuda.device_of(f1):
                _lib.call("ocf_corr_bwd", _p(g), _p(act), _p(f1), _p(f2), _p(df1), _p(df2), B, C, H, W, ctx.d, g_bstride, 0, ctx.slope, _p(mask), _stream())
        return df1, df2, None, None


def cost_volume(f1, f2, max_displacement=4, leaky_slope=1.0):
    f1 = _req(f1, "features1", 4)
    f2 = _req(f2, "features2", 4)
    if f1.shape != f2.shape:
        raise ValueError("features1 and features2 must have the same shape (got %s vs %s)" % (tuple(f1.shape), tuple(f2.shape)))
    if f1.numel() == 0:
        # empty batch / empty image: the reference's slice-multiply-mean chain returns an empty cost volume
        nd = 2 * int(max_displacement) + 1
        return (f1.sum() + f2.sum()) * 0 + f1.new_zeros((f1.shape[0], nd * nd, f1.shape[2], f1.shape[3]))
    W = f1.shape[3]
    if W % 4 != 0 and int(max_displacement) == 4 and W >= 16:
        # Ragged rows (KITTI / Sintel pyramids: 621, 311, 39 ...) cannot be described to the TMA unit (global strides must be
        # multiples of 16 bytes).  Zero columns on the right are exactly the reference's out-of-image zeros, so the rows are
        # re-pitched to a multiple of 4 and the regular kernels run; autograd slices / pads the gradients the same way.
        pad = (-W) % 4
        out = _CostVolume.apply(torch.nn.functional.pad(f1, (0, pad)), torch.nn.functional.pad(f2, (0, pad)),
                                int(max_displacement), float(leaky_slope))
        return out[..., :W]
    return _CostVolume.apply(f1, f2, int(max_displacement), float(leaky_slope))


# -------------------------------------------------------------------------------------------------
# feature normalisation
# -------------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


class _Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flags, *xs):
        T = len(xs)
        B, C, H, W = xs[0].shape
        G = 1 if flags & NORM_ACROSS_CHANNELS else C
        NG = T * B * G
        ys = [torch.empty_like(x) for x in xs]
        stats = torch.empty(8 * NG + 8, device=xs[0].device, dtype=torch.float32)
        with torch.cuda.device_of(xs[0]):
            _lib.call("ocf_normalize_fwd", ctypes.cast(_ptr_array(xs), ctypes.c_void_p), ctypes.cast(_ptr_array(ys), ctypes.c_void_p),
                      T, B, C, H, W, flags, _p(stats), _stream())
        ctx.flags, ctx.NG = flags, NG
        ctx.save_for_backward(stats, *xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gs):
        stats, xs = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        T = len(xs)
        B, C, H, W = xs[0].shape
        gs = [torch.zeros_like(x) if g is None else g.contiguous() for g, x in zip(gs, xs)]
        dxs = [torch.empty_like(x) for x in xs]
        red = torch.empty(8 * ctx.NG, device=xs[0].device, dtype=torch.float32)
        with torch.cuda.device_of(xs[0]):
            _lib.call("ocf_normalize_bwd", ctypes.cast(_ptr_array(gs), ctypes.c_void_p), ctypes.cast(_ptr_array(xs), ctypes.c_void_p),
                      ctypes.cast(_ptr_array(dxs), ctypes.c_void_p), T, B, C, H, W, ctx.flags, _p(stats), _p(red), _stream())
        return (None,) + tuple(dxs)


def normalize_features(feature_list, normalize=True, center=True, moments_across_channels=True, moments_across_images=True):
    xs = [_req(f, "feature_list[%d]" % i, 4) for i, f in enumerate(feature_list)]
    if not xs:
        return []
    if not (normalize or center):
        return list(feature_list)
    for x in xs[1:]:
        if x.shape != xs[0].shape:
            raise ValueError("all tensors of feature_list must share one shape")
    flags = (NORM_NORMALIZE if normalize else 0) | (NORM_CENTER if center else 0) | \
        (NORM_ACROSS_CHANNELS if moments_across_channels else 0) | (NORM_ACROSS_IMAGES if moments_across_images else 0)
    return list(_Normalize.apply(flags, *xs))


# -------------------------------------------------------------------------------------------------
# fused FlowNetCV pyramid level: warp -> normalize_features -> compute_cost_volume -> LeakyReLU -> cat
# (cost_volume_flow_net.py:171-173 and 186-190 / 201-205 / 216-220 / 231-235)
# -------------------------------------------------------------------------------------------------
_NORM_DEFAULT = NORM_NORMALIZE | NORM_CENTER | NORM_ACROSS_CHANNELS | NORM_ACROSS_IMAGES
# Which correlation the fused level runs.  False (default): statistics -> apply (c1n straight into the concat buffer) -> fp32 FMA
# correlation (TMA-fed, reads c1n in place): 4 launches, the fastest form at every FlowNetCV level on B200.  True: statistics ->
# tcgen05 3xTF32 correlation normalising on load: 3 launches, 1e-6 instead of 1e-7 relative error, and faster only for ragged
# row lengths (W % 4 != 0, where TMA cannot describe the rows) -- those always take it.
LEVEL_TENSOR_CORES = os.environ.get("OCF_LEVEL_TC", "0") == "1"


class _LevelFused(torch.autograd.Function):
    """x = cat(LeakyReLU(corr(c1n, c2n)), c1n[, up_flow, up_feat]) with [c1n, c2n] = normalize_features([c1, warp(c2, up_flow*scale)]).

    Forward: warp, statistics, apply (c1n written straight into the concat buffer), correlation reading c1n in place and writing
    the cost volume into the same buffer -- or, with LEVEL_TENSOR_CORES / ragged rows, statistics + one tensor-core correlation
    that normalises on load; either way no torch.cat copy of the 81 + C widest channels.  Backward: correlation backward (c1n read in place from the buffer), the normalisation
    backward through the statistics, the warp backward."""

    @staticmethod
    def forward(ctx, c1, c2, up_flow, up_feat, scale, slope):
        B, C, H, W = c1.shape
        has_flow = up_flow is not None
        # coarsest level (cost_volume_flow_net.py:171-176): the decoder input is the cost volume alone
        width = 81 + C + up_flow.shape[1] + up_feat.shape[1] if has_flow else 81
        X = torch.empty((B, width, H, W), device=c1.device, dtype=torch.float32)
        with torch.cuda.device_of(c1):
            if has_flow:
                w2 = torch.empty_like(c2)
                _lib.call("ocf_warp_fwd", _p(c2), _p(up_flow), None, _p(w2), B, C, H, W, 0, float(scale), _stream())
            else:
                w2 = c2
            NG = 2 * B
            stats = torch.empty(8 * NG + 8, device=c1.device, dtype=torch.float32)
            _lib.call("ocf_normalize_stats", ctypes.cast(_ptr_array([c1, w2]), ctypes.c_void_p), 2, B, C, H, W, _NORM_DEFAULT, _p(stats), _stream())
            f2n = torch.empty_like(c2)
            mask = torch.empty((B, 81, H, (W + 7) // 8), device=c1.device, dtype=torch.uint8)
            f1n = X[:, 81:81 + C] if has_flow else torch.empty_like(c1)
            if LEVEL_TENSOR_CORES or W % 4 != 0:
                norm = stats[6 * NG:6 * NG + 2]      # {mean, inv_std}: one scalar pair for all groups (moments_across_images)
                _lib.call("ocf_level_corr_fwd", _p(c1), _p(w2), _p(norm), _p(X), X.stride(0), _p(f1n), f1n.stride(0), _p(f2n), _p(mask),
                          B, C, H, W, float(slope), _stream())
            else:
                # fp32 FMA form: the apply pass leaves c1n where the concat wants it, the TMA-fed correlation reads it in place
                bstr = (ctypes.c_longlong * 2)(f1n.stride(0), 0)
                _lib.call("ocf_normalize_apply", ctypes.cast(_ptr_array([c1, w2]), ctypes.c_void_p), ctypes.cast(_ptr_array([f1n, f2n]), ctypes.c_void_p),
                          ctypes.cast(bstr, ctypes.c_void_p), 2, B, C, H, W, _NORM_DEFAULT, _p(stats), _stream())
                _lib.call("ocf_corr_fwd_strided", _p(f1n), f1n.stride(0), _p(f2n), _p(X), X.stride(0), _p(mask), B, C, H, W, float(slope), _stream())
            if has_flow:
                X[:, 81 + C:81 + C + up_flow.shape[1]].copy_(up_flow)
                X[:, 81 + C + up_flow.shape[1]:].copy_(up_feat)
        ctx.cfg = (float(scale), float(slope), has_flow, up_flow.shape[1] if has_flow else 0)
        ctx.save_for_backward(c1, c2, up_flow if has_flow else c1.new_empty(0), w2, f2n, mask, stats, X if has_flow else f1n)
        return X

    @staticmethod
    def backward(ctx, gX):
        c1, c2, up_flow, w2, f2n, mask, stats, keep = ctx.saved_tensors
        scale, slope, has_flow, nflow = ctx.cfg
        B, C, H, W = c1.shape
        NG = 2 * B
        if not (gX.stride(3) == 1 and gX.stride(2) == W and gX.stride(1) == H * W and gX.stride(0) % 4 == 0 and gX.data_ptr() % 16 == 0):
            gX = gX.contiguous()
        f1n = keep[:, 81:81 + C] if has_flow else keep
        dfn1 = torch.empty_like(c1)
        dfn2 = torch.empty_like(c2)
        with torch.cuda.device_of(c1):
            _lib.call("ocf_level_corr_bwd", _p(gX), gX.stride(0), _p(mask), _p(f1n), f1n.stride(0), _p(f2n), _p(dfn1), _p(dfn2), B, C, H, W,
                      slope, _stream())
            if has_flow:
                dfn1.add_(gX[:, 81:81 + C])      # c1n is also an output (concatenated into the decoder input)
            dc1 = torch.empty_like(c1)
            dw2 = torch.empty_like(c2)
            red = torch.empty(8 * NG, device=c1.device, dtype=torch.float32)
            _lib.call("ocf_normalize_bwd", ctypes.cast(_ptr_array([dfn1, dfn2]), ctypes.c_void_p), ctypes.cast(_ptr_array([c1, w2]), ctypes.c_void_p),
                      ctypes.cast(_ptr_array([dc1, dw2]), ctypes.c_void_p), 2, B, C, H, W, _NORM_DEFAULT, _p(stats), _p(red), _stream())
            if not has_flow:
                return dc1, dw2, None, None, None, None
            need_c2, need_flow = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
            dc2 = torch.empty_like(c2) if need_c2 else None
            dflow = torch.empty_like(up_flow) if need_flow else None
            if need_c2 or need_flow:
                _lib.call("ocf_warp_bwd", _p(dw2), _p(c2), _p(up_flow), None, _p(dc2), _p(dflow), None, B, C, H, W, 0, scale, _stream())
            if need_flow:
                dflow.add_(gX[:, 81 + C:81 + C + nflow])
            dfeat = gX[:, 81 + C + nflow:] if ctx.needs_input_grad[3] else None
        return dc1, dc2, dflow, dfeat, None, None


def level_fused(c1, c2, up_flow=None, up_feat=None, flow_scale=1.0, leaky_slope=0.1):
    """One FlowNetCV decoder level input (d = 4): see _LevelFused.  up_flow / up_feat are None at the coarsest level."""
    c1 = _req(c1, "c1", 4)
    c2 = _req(c2, "c2", 4)
    if c1.shape != c2.shape:
        raise ValueError("c1 and c2 must have the same shape")
    if (up_flow is None) != (up_feat is None):
        raise ValueError("up_flow and up_feat go together")
    if up_flow is not None:
        up_flow = _req(up_flow, "up_flow", 4)
        up_feat = _req(up_feat, "up_feat", 4)
        B, C, H, W = c1.shape
        if up_flow.shape != (B, 2, H, W) or up_feat.shape[0] != B or up_feat.shape[2:] != (H, W):
            raise ValueError("up_flow must be [B,2,H,W] and up_feat [B,*,H,W] matching c1")
    return _LevelFused.apply(c1, c2, up_flow, up_feat, float(flow_scale), float(leaky_slope))


# -------------------------------------------------------------------------------------------------
# bilinear backward warp
# -------------------------------------------------------------------------------------------------
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, flow, occ, flags, scale):
        B, C, H, W = img.shape
        out = torch.empty_like(img)
        with torch.cuda.device_of(img):
            _lib.call("ocf_warp_fwd", _p(img), _p(flow), _p(occ), _p(out), B, C, H, W, flags, float(scale), _stream())
        ctx.flags, ctx.scale, ctx.has_occ = flags, float(scale), occ is not None
        if occ is not None:
            ctx.save_for_backward(img, flow, occ)
        else:
            ctx.save_for_backward(img, flow)
        return out

    @staticmethod
    def backward(ctx, g):
        img, flow = ctx.saved_tensors[0], ctx.saved_tensors[1]
        occ = ctx.saved_tensors[2] if ctx.has_occ else None
        B, C, H, W = img.shape
        g = g.contiguous()
        need_img, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_occ = ctx.has_occ and ctx.needs_input_grad[2]
        d_img = torch.empty_like(img) if need_img else None
        d_flow = torch.empty_like(flow) if need_flow else None
        d_occ = torch.empty_like(occ) if need_occ else None
        if need_img or need_flow or need_occ:
            with torch.cuda.device_of(img):
                _lib.call("ocf_warp_bwd", _p(g), _p(img), _p(flow), _p(occ), _p(d_img), _p(d_flow), _p(d_occ), B, C, H, W, ctx.flags,
                          ctx.scale, _stream())
        return d_img, d_flow, d_occ, None, None


def warp(img, flow, align_corners=True, is_mask=False, occ=None, flow_scale=1.0):
    img = _req(img, "img", 4)
    flow = _req(flow, "flow", 4)
    B, C, H, W = img.shape
    if flow.shape != (B, 2, H, W):
        raise ValueError("flow must be [B,2,H,W] matching img (got %s for img %s)" % (tuple(flow.shape), tuple(img.shape)))
    if occ is not None:
        occ = _req(occ, "occ", 4)
        if occ.shape != (B, 1, H, W):
            raise ValueError("occ must be [B,1,H,W]")
    if img.numel() == 0:
        return (img.sum() + flow.sum()) * 0 + torch.zeros_like(img)      # empty batch: nothing to sample (grad-connected like the op)
    flags = (WARP_ALIGN_CORNERS if align_corners else 0) | (WARP_IS_MASK if is_mask else 0)
    return _Warp.apply(img, flow, occ, flags, float(flow_scale))


# -------------------------------------------------------------------------------------------------
# bilinear resize, align_corners=True (the F.interpolate glue of cost_volume_flow_net.py:245 and models/model.py:396)
# -------------------------------------------------------------------------------------------------
class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo, mul):
        B, C, Hi, Wi = x.shape
        out = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32)
        with torch.cuda.device_of(x):
            _lib.call("ocf_resize_bilinear_fwd", _p(x), _p(out), B * C, Hi, Wi, Ho, Wo, float(mul), _stream())
        ctx.cfg = (B, C, Hi, Wi, Ho, Wo, float(mul))
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, Hi, Wi, Ho, Wo, mul = ctx.cfg
        g = g.contiguous()
        gin = torch.empty((B, C, Hi, Wi), device=g.device, dtype=torch.float32)
        with torch.cuda.device_of(g):
            _lib.call("ocf_resize_bilinear_bwd", _p(g), _p(gin), B * C, Hi, Wi, Ho, Wo, mul, _stream())
        return gin, None, None, None


def resize_bilinear(x, size=None, scale_factor=None, mul=1.0):
    """mul * F.interpolate(x, size / scale_factor, mode='bilinear', align_corners=True) in one launch (gather backward)."""
    x = _req(x, "x", 4)
    if (size is None) == (scale_factor is None):
        raise ValueError("exactly one of size and scale_factor must be given")
    if size is None:
        Ho, Wo = int(x.shape[2] * scale_factor), int(x.shape[3] * scale_factor)   # floor, as F.interpolate
    else:
        Ho, Wo = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    if Ho < 1 or Wo < 1:
        raise ValueError("output size must be positive (got %dx%d)" % (Ho, Wo))
    return _Resize.apply(x, Ho, Wo, float(mul))


# -------------------------------------------------------------------------------------------------
# bias + LeakyReLU epilogue of the convolution blocks (cost_volume_flow_net.py:11-15)
# -------------------------------------------------------------------------------------------------
class _BiasLeakyReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, slope):
        B, C, H, W = x.shape
        # in place: x is the fresh output of the (bias-free) convolution, whose backward does not need it
        with torch.cuda.device_of(x):
            _lib.call("ocf_bias_lrelu_fwd", _p(x), _p(bias), _p(x), B, C, H * W, float(slope), _stream())
        ctx.slope = float(slope)
        ctx.mark_dirty(x)
        ctx.save_for_backward(x)
        return x

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        B, C, H, W = y.shape
        g = g.contiguous()
        dx = torch.empty_like(y)
        dbias = torch.empty(C, device=y.device, dtype=torch.float32)
        with torch.cuda.device_of(y):
            _lib.call("ocf_bias_lrelu_bwd", _p(g), _p(y), _p(dx), _p(dbias), B, C, H * W, ctx.slope, _stream())
        return dx, dbias, None


def bias_leaky_relu_(x, bias, negative_slope=0.1):
    """lrelu(x + bias[None, :, None, None]) computed in place on x (the output of a bias-free convolution)."""
    x = _req(x, "x", 4)
    bias = _req(bias, "bias", 1)
    if bias.shape[0] != x.shape[1]:
        raise ValueError("bias must have one entry per channel")
    if x.numel() == 0:
        return x
    return _BiasLeakyReLU.apply(x, bias, float(negative_slope))


# -------------------------------------------------------------------------------------------------
# range map / occlusion  (forward only: the reference calls it under no_grad, models/model.py:381-391)
# -------------------------------------------------------------------------------------------------
def range_map(flow, with_occlusion=False):
    flow = _req(flow.detach(), "flow", 4)
    B, two, H, W = flow.shape
    if two != 2:
        raise ValueError("flow must be [B,2,H,W]")
    rmap = torch.empty((B, 1, H, W), device=flow.device, dtype=torch.float32)
    occ = torch.empty_like(rmap) if with_occlusion else None
    if rmap.numel() == 0:
        return (rmap, occ) if with_occlusion else rmap
    with torch.cuda.device_of(flow):
        _lib.call("ocf_range_map", _p(flow), _p(rmap), _p(occ), B, H, W, _stream())
    if with_occlusion:
        _lib.launch_count += 1
        return rmap, occ
    return rmap


def flow_to_warp(flow_bhw2):
    flow = _req(flow_bhw2.detach(), "flow", 4)
    B, H, W, two = flow.shape
    if two != 2:
        raise ValueError("flow must be [B,H,W,2]")
    out = torch.empty_like(flow)
    with torch.cuda.device_of(flow):
        _lib.call("ocf_flow_to_warp", _p(flow), _p(out), B, H, W, _stream())
    return out


# -------------------------------------------------------------------------------------------------
# Charbonnier / photometric
# -------------------------------------------------------------------------------------------------
class _RobustL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        y = torch.empty_like(x)
        with torch.cuda.device_of(x):
            _lib.call("ocf_robust_l1_fwd", _p(x), _p(y), x.numel(), float(alpha), _stream())
        ctx.alpha = float(alpha)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        g = g.contiguous()
        with torch.cuda.device_of(x):
            _lib.call("ocf_robust_l1_bwd", _p(g), _p(x), _p(gx), x.numel(), ctx.alpha, _stream())
        return gx, None


def robust_l1(x, alpha=0.001):
    x = _req(x, "x")
    if x.numel() == 0:
        return torch.empty_like(x)
    return _RobustL1.apply(x, float(alpha))


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, img, occ, alpha):
        B, C, H, W = pred.shape
        sums = torch.empty(2, device=pred.device, dtype=torch.float64)
        with torch.cuda.device_of(pred):
            _lib.call("ocf_photometric_fwd", _p(pred), _p(img), _p(occ), _p(sums), B, C, H, W, float(alpha), _stream())
        if occ is not None:
            den = sums[1] * 3 + 1e-16           # the literal 3 of models/model.py:43
        else:
            den = sums[1] * 0 + float(B * C * H * W)   # torch.mean; built from a device scalar (graph-capture safe)
        loss = (sums[0] / den).to(torch.float32)
        ctx.alpha, ctx.has_occ = float(alpha), occ is not None
        if occ is not None:
            ctx.save_for_backward(pred, img, sums, den, occ)
        else:
            ctx.save_for_backward(pred, img, sums, den)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, img, sums, den = ctx.saved_tensors[:4]
        occ = ctx.saved_tensors[4] if ctx.has_occ else None
        B, C, H, W = pred.shape
        g64 = g.to(torch.float64)
        k0 = g64 / den
        k1 = -3.0 * g64 * sums[0] / (den * den) if ctx.has_occ else torch.zeros_like(k0)
        coef = torch.stack((k0, k1)).to(torch.float32)
        need_pred, need_img = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_occ = ctx.has_occ and ctx.needs_input_grad[2]
        d_pred = torch.empty_like(pred) if need_pred else None
        d_img = torch.empty_like(img) if need_img else None
        d_occ = torch.empty_like(occ) if need_occ else None
        if need_pred or need_img or need_occ:
            with torch.cuda.device_of(pred):
                _lib.call("ocf_photometric_bwd", _p(pred), _p(img), _p(occ), _p(coef), _p(d_pred), _p(d_img), _p(d_occ), B, C, H, W,
                          ctx.alpha, _stream())
        return d_pred, d_img, d_occ, None


def photometric_error(img_pred, img, occ=None, alpha=0.001):
    pred = _req(img_pred, "img_pred", 4)
    img = _req(img, "img", 4)
    if pred.shape != img.shape:
        raise ValueError("img_pred and img must have the same shape")
    if occ is not None:
        occ = _req(occ, "occ", 4)
        if occ.shape != (pred.shape[0], 1, pred.shape[2], pred.shape[3]):
            raise ValueError("occ must be [B,1,H,W]")
    return _Photometric.apply(pred, img, occ, float(alpha))


# -------------------------------------------------------------------------------------------------
# smoothness
# -------------------------------------------------------------------------------------------------
class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, flow, order, alpha_edge, alpha_rho):
        B, Ci, H, W = img.shape
        Cf = flow.shape[1]
        sums = torch.empty(2, device=img.device, dtype=torch.float64)
        with torch.cuda.device_of(img):
            _lib.call("ocf_smooth_fwd", _p(img), _p(flow), _p(sums), B, Ci, Cf, H, W, order, float(alpha_edge), float(alpha_rho), _stream())
        nx = float(B * Cf * H * (W - order))
        ny = float(B * Cf * (H - order) * W)
        ctx.cfg = (order, float(alpha_edge), float(alpha_rho), nx, ny)
        ctx.save_for_backward(img, flow)
        # python-scalar arithmetic only: no host->device tensor creation, so the op stays CUDA-graph capturable
        return (sums[0] * (0.5 / nx) + sums[1] * (0.5 / ny)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        img, flow = ctx.saved_tensors
        order, ae, ar, nx, ny = ctx.cfg
        B, Ci, H, W = img.shape
        Cf = flow.shape[1]
        coef = torch.stack((g * (0.5 / nx), g * (0.5 / ny))).to(torch.float32).contiguous()
        need_img, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_img = torch.empty_like(img) if need_img else None
        d_flow = torch.empty_like(flow) if need_flow else None
        if need_img or need_flow:
            with torch.cuda.device_of(img):
                _lib.call("ocf_smooth_bwd", _p(img), _p(flow), _p(coef), _p(d_img), _p(d_flow), B, Ci, Cf, H, W, order, ae, ar, _stream())
        return d_img, d_flow, None, None, None


def smoothness_loss(img, flow, order, alpha=100.0, alpha_rho=0.001):
    img = _req(img, "img", 4)
    flow = _req(flow, "flow", 4)
    if img.shape[0] != flow.shape[0] or img.shape[2:] != flow.shape[2:]:
        raise ValueError("img and flow must share batch and spatial size")
    if img.shape[2] <= order or img.shape[3] <= order:
        raise ValueError("image too small for a stride-%d gradient" % order)
    return _Smooth.apply(img, flow, int(order), float(alpha), float(alpha_rho))


class _Gradient(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, stride):
        B, C, H, W = img.shape
        dx = torch.empty((B, C, H, W - stride), device=img.device, dtype=torch.float32)
        dy = torch.empty((B, C, H - stride, W), device=img.device, dtype=torch.float32)
        with torch.cuda.device_of(img):
            _lib.call("ocf_gradient", _p(img), _p(dx), _p(dy), B, C, H, W, stride, _stream())
        ctx.stride = stride
        return dx, dy

    @staticmethod
    def backward(ctx, gdx, gdy):
        # adjoint of a forward difference is a (negated) backward difference: plain tensor glue
        s = ctx.stride
        B, C, H, Wm = gdx.shape
        W = Wm + s
        g = gdx.new_zeros((B, C, H, W))
        g[:, :, :, s:] += gdx
        g[:, :, :, :-s] -= gdx
        g[:, :, s:, :] += gdy
        g[:, :, :-s, :] -= gdy
        return g, None


def gradient(img, stride=1):
    img = _req(img, "img", 4)
    return _Gradient.apply(img, int(stride))


# -------------------------------------------------------------------------------------------------
# supervised pair losses
# -------------------------------------------------------------------------------------------------
PAIR_L1, PAIR_MSE, PAIR_BCE, PAIR_FOCAL = 0, 1, 2, 3


class _PairLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, kind):
        s = torch.empty(1, device=a.device, dtype=torch.float64)
        need = ctx.needs_input_grad[0]
        grad = torch.empty_like(a) if need else None
        with torch.cuda.device_of(a):
            _lib.call("ocf_pair_loss", _p(a), _p(b), _p(s), _p(grad), a.numel(), kind, _stream())
        ctx.n = a.numel()
        if need:
            ctx.save_for_backward(grad)
        return (s[0] / a.numel()).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * (g / ctx.n), None, None


def pair_loss(a, b, kind):
    a = _req(a, "input")
    b = _req(b.detach(), "target")
    if a.shape != b.shape:
        raise ValueError("input and target must have the same shape")
    return _PairLoss.apply(a, b, int(kind))


# -------------------------------------------------------------------------------------------------
# SSIM (inpainting_metrics/ssim/ssim.py)
# -------------------------------------------------------------------------------------------------
class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, window, size_average):
        B, C, H, W = img1.shape
        Ho, Wo = H + 2 * (window // 2) - window + 1, W + 2 * (window // 2) - window + 1
        sums = torch.empty(B, device=img1.device, dtype=torch.float64)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        coef = torch.empty((B, C, 4, Ho, Wo), device=img1.device, dtype=torch.float32) if need else None
        with torch.cuda.device_of(img1):
            _lib.call("ocf_ssim_fwd", _p(img1), _p(img2), _p(sums), _p(coef), B, C, H, W, window, _stream())
        ctx.window, ctx.size_average, ctx.numel = window, size_average, C * Ho * Wo
        if need:
            ctx.save_for_backward(img1, img2, coef)
        per_item = sums / float(C * Ho * Wo)
        # ssim_map.mean() == mean over items of the per-item means (all items have the same number of elements)
        return (per_item.mean() if size_average else per_item).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        img1, img2, coef = ctx.saved_tensors
        B, C, H, W = img1.shape
        if ctx.size_average:
            scale = (g.to(torch.float32) / float(B * ctx.numel)).expand(B).contiguous()
        else:
            scale = (g.to(torch.float32) / float(ctx.numel)).contiguous()
        d1 = torch.empty_like(img1) if ctx.needs_input_grad[0] else None
        d2 = torch.empty_like(img2) if ctx.needs_input_grad[1] else None
        with torch.cuda.device_of(img1):
            _lib.call("ocf_ssim_bwd", _p(img1), _p(img2), _p(coef), _p(scale), _p(d1), _p(d2), B, C, H, W, ctx.window, _stream())
        return d1, d2, None, None


def ssim(img1, img2, window_size=11, size_average=True):
    img1 = _req(img1, "img1", 4)
    img2 = _req(img2, "img2", 4)
    if img1.shape != img2.shape:
        raise ValueError("img1 and img2 must have the same shape")
    if not 1 <= int(window_size) <= 15:
        raise ValueError("window_size must be in [1, 15] (got %d)" % window_size)
    return _SSIM.apply(img1, img2, int(window_size), bool(size_average))


# -------------------------------------------------------------------------------------------------
# soft census term (no reference definition: parity unpinned, oracle census_loss)
# -------------------------------------------------------------------------------------------------
class _Census(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, img, occ, max_distance):
        B, C, H, W = pred.shape
        sums = torch.empty(2, device=pred.device, dtype=torch.float64)
        with torch.cuda.device_of(pred):
            _lib.call("ocf_census_fwd", _p(pred), _p(img), _p(occ), _p(sums), B, C, H, W, max_distance, _stream())
        den = sums[1] + 1e-16
        ctx.m, ctx.has_occ = max_distance, occ is not None
        if occ is not None:
            ctx.save_for_backward(pred, img, den, occ)
        else:
            ctx.save_for_backward(pred, img, den)
        return (sums[0] / den).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        pred, img, den = ctx.saved_tensors[:3]
        occ = ctx.saved_tensors[3] if ctx.has_occ else None
        B, C, H, W = pred.shape
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        coef = (g.to(torch.float64) / den).to(torch.float32).reshape(1).contiguous()
        d_pred = torch.empty_like(pred)
        with torch.cuda.device_of(pred):
            _lib.call("ocf_census_bwd", _p(pred), _p(img), _p(occ), _p(coef), _p(d_pred), B, C, H, W, ctx.m, _stream())
        return d_pred, None, None, None


def census_loss(img_pred, img, occ=None, max_distance=3):
    """Occlusion-weighted soft census distance between img_pred and img (gradient flows to img_pred only)."""
    pred = _req(img_pred, "img_pred", 4)
    img = _req(img.detach(), "img", 4)
    if pred.shape != img.shape:
        raise ValueError("img_pred and img must have the same shape")
    if occ is not None:
        occ = _req(occ.detach(), "occ", 4)
        if occ.shape != (pred.shape[0], 1, pred.shape[2], pred.shape[3]):
            raise ValueError("occ must be [B,1,H,W]")
    if not 1 <= int(max_distance) <= 3:
        raise ValueError("max_distance must be 1, 2 or 3 (got %d)" % max_distance)
    return _Census.apply(pred, img, occ, int(max_distance))


# -------------------------------------------------------------------------------------------------
# fused occlusion-aware photometric pass (models/model.py:379-407 in one kernel)
# -------------------------------------------------------------------------------------------------
class _OccPhotoFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, flow, rmap, flow_gt, occ_gt, alpha):
        B, C, H, W = img1.shape
        sums = torch.empty(8, device=img1.device, dtype=torch.float64)
        need = ctx.needs_input_grad[2]
        dflow = torch.empty_like(flow) if need else None
        with torch.cuda.device_of(img1):
            _lib.call("ocf_occ_photo_fused", _p(img1), _p(img2), _p(flow), _p(rmap), _p(flow_gt), _p(occ_gt), _p(sums), _p(dflow), None,
                      B, C, H, W, float(alpha), _stream())
        den_vis = sums[1] * 3 + 1e-16
        den_occ = sums[3] * 3 + 1e-16
        photo = (sums[0] / den_vis).to(torch.float32)
        photo_occ = (sums[2] / den_occ).to(torch.float32)
        # terms whose ground truth is absent are NaN, not 0.0 (a zero would read as a perfect score)
        nan = float("nan")
        mse = (sums[4] / float(B * 2 * H * W)).to(torch.float32) if flow_gt is not None else sums[4].to(torch.float32) * 0 + nan
        bce = (sums[5] / float(B * H * W)).to(torch.float32) if occ_gt is not None else sums[5].to(torch.float32) * 0 + nan
        if need:
            ctx.save_for_backward(dflow, den_vis)
        ctx.mark_non_differentiable(photo_occ, mse, bce)
        return photo, photo_occ, mse, bce

    @staticmethod
    def backward(ctx, g_photo, g_pocc, g_mse, g_bce):
        dflow, den_vis = ctx.saved_tensors
        k = (g_photo.to(torch.float64) / den_vis).to(torch.float32)
        return None, None, dflow * k, None, None, None, None


def occ_photo_fused(img1, img2, flow, rmap=None, flow_gt=None, occ_gt=None, alpha=0.001):
    """(photo, photo_occ, flow_mse, occ_bce) of general_step_occ_aware in one pass; grad flows to `flow` via photo only
    (the other three are logged-only scalars in the reference, models/model.py:403-407)."""
    img1 = _req(img1.detach(), "img1", 4)
    img2 = _req(img2.detach(), "img2", 4)
    flow = _req(flow, "flow", 4)
    if img1.shape[1] != 3 or img2.shape != img1.shape:
        # the denominators carry the reference's literal 3 (models/model.py:43): the equivalence with photometric_error /
        # torch.mean only holds for 3-channel images
        raise ValueError("occ_photo_fused expects two [B,3,H,W] images (got %s and %s)" % (tuple(img1.shape), tuple(img2.shape)))
    rmap = None if rmap is None else _req(rmap.detach(), "range_map", 4)
    flow_gt = None if flow_gt is None else _req(flow_gt.detach(), "flow_gt", 4)
    occ_gt = None if occ_gt is None else _req(occ_gt.detach(), "occ_gt", 4)
    return _OccPhotoFused.apply(img1, img2, flow, rmap, flow_gt, occ_gt, float(alpha))
