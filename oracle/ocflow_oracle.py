"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OCFlow hot path (the parity oracle).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this file; the product package ``ocflow_b200`` never does
(tests/test_host_logic.py::test_product_never_imports_oracle_or_falls_back enforces it).  Everything here is plain, dtype-generic PyTorch on the CPU -- the reference's algorithm
is floating-point tensor math whose arithmetic lives in ATen, so a torch restatement (fp32 for the
parity bar, fp64 as tie-breaker) is the natural oracle; nothing here calls ``F.grid_sample``,
``scatter_add_`` on nonzero masks, or any reference code.

PINNING: every function below is checked against the *real* reference, imported in the build
container by ``oracle/ref_loader.py`` (tests/test_oracle_vs_reference.py, skipped where
/root/reference is absent) and against the committed fixtures ``tests/golden/*.pt`` that
``oracle/make_golden.py`` generated from the real reference (tests/test_oracle_golden.py, runs
everywhere).  One symbol is *parity unpinned*: ``CostVolumeLayer`` -- its source file
(models/networks/cost_volume_net.py) is missing from the reference; semantics are inferred from the
call sites (cost_volume_flow_occ_net.py:53,56,188; flow_occ_net_c.py:26,29,99).

Citations are ``file:line`` relative to the reference root.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# a-1  cost volume                                        models/networks/correlation_layer.py:7-40
# --------------------------------------------------------------------------------------------


def cost_volume(f1, f2, max_displacement=4):
    """out[b, (dy+d)*(2d+1)+(dx+d), y, x] = mean_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx], 0 outside the image.

    Restated without padding: each displacement plane is filled only on the rectangle where the
    shifted f2 pixel exists (correlation_layer.py:33-39 pads with zeros and slices instead).
    """
    B, C, H, W = f1.shape
    d = int(max_displacement)
    n = 2 * d + 1
    out = f1.new_zeros(B, n * n, H, W)
    for iy in range(n):
        dy = iy - d
        ya, yb = max(0, -dy), min(H, H - dy)
        if ya >= yb:
            continue
        for ix in range(n):
            dx = ix - d
            xa, xb = max(0, -dx), min(W, W - dx)
            if xa >= xb:
                continue
            prod = f1[:, :, ya:yb, xa:xb] * f2[:, :, ya + dy:yb + dy, xa + dx:xb + dx]
            out[:, iy * n + ix, ya:yb, xa:xb] = prod.mean(dim=1)
    return out


# --------------------------------------------------------------------------------------------
# a-2  feature normalisation                               models/networks/correlation_layer.py:42-82
# --------------------------------------------------------------------------------------------


def normalize_features(feature_list, normalize=True, center=True, moments_across_channels=True,
                       moments_across_images=True):
    """Biased per-sample (or per-sample-per-channel) moments; optionally averaged to ONE scalar pair
    over all tensors/samples(/channels) (correlation_layer.py:66-68); std = sqrt(var + 1e-16) (:70)."""
    dims = (1, 2, 3) if moments_across_channels else (2, 3)
    means, variances = [], []
    for f in feature_list:
        mu = f.mean(dim=dims, keepdim=True)
        var = ((f - mu) ** 2).mean(dim=dims, keepdim=True)
        means.append(mu)
        variances.append(var)
    if moments_across_images:
        gm = torch.stack(means).mean()
        gv = torch.stack(variances).mean()
        means = [gm for _ in feature_list]
        variances = [gv for _ in feature_list]
    out = list(feature_list)
    if center:
        out = [f - m for f, m in zip(out, means)]
    if normalize:
        out = [f / torch.sqrt(v + 1e-16) for f, v in zip(out, variances)]
    return out


# --------------------------------------------------------------------------------------------
# a-4 / a-5  bilinear backward warp      utils.py:20-58, models/model.py:191-221,
#            cost_volume_flow_net.py:121-151, pwc_net.py:6-29 (+ 7 more identical bodies)
# --------------------------------------------------------------------------------------------


def sample_coords(flow, align_corners):
    """Pixel-space sampling coordinates exactly as reference + ATen produce them.

    Reference: g = 2*(x+u)/max(W-1,1) - 1 (utils.py:43-44).  ATen grid_sampler un-normalises with
    ((g+1)/2)*(W-1) when align_corners else ((g+1)*W-1)/2 (ATen/native/GridSampler.h
    grid_sampler_unnormalize).  Kept as separate ops, in this order, so fp32 rounding matches.
    """
    B, _, H, W = flow.shape
    xs = torch.arange(W, dtype=flow.dtype, device=flow.device).view(1, 1, W)
    ys = torch.arange(H, dtype=flow.dtype, device=flow.device).view(1, H, 1)
    gx = 2.0 * (xs + flow[:, 0]) / max(W - 1, 1) - 1.0
    gy = 2.0 * (ys + flow[:, 1]) / max(H - 1, 1) - 1.0
    if align_corners:
        ix = ((gx + 1.0) / 2.0) * (W - 1)
        iy = ((gy + 1.0) / 2.0) * (H - 1)
    else:
        ix = ((gx + 1.0) * W - 1.0) / 2.0
        iy = ((gy + 1.0) * H - 1.0) / 2.0
    return ix, iy


def bilinear_sample(img, ix, iy):
    """Zero-padded bilinear gather: each of the 4 taps is dropped independently when outside."""
    B, C, H, W = img.shape
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx = ix - x0
    fy = iy - y0
    flat = img.reshape(B, C, H * W)
    out = img.new_zeros(B, C, ix.shape[-2], ix.shape[-1])
    for ddy, wy in ((0, 1.0 - fy), (1, fy)):
        for ddx, wx in ((0, 1.0 - fx), (1, fx)):
            xi = (x0 + ddx)
            yi = (y0 + ddy)
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().reshape(B, 1, -1).expand(B, C, -1)
            tap = torch.gather(flat, 2, lin).reshape(out.shape)
            out = out + tap * (wx * wy * ok.to(img.dtype)).unsqueeze(1)
    return out


def warp(img, flow, align_corners=True, is_mask=False):
    """align_corners=True: utils.py:20-58 / models/model.py:191-221 / pwc_net.py:6-29;
    align_corners=False: cost_volume_flow_net.py:121-151 and the 5 other network copies.
    is_mask (utils.py:49-57): zero every output pixel whose bilinear footprint is not fully inside."""
    ix, iy = sample_coords(flow, align_corners)
    out = bilinear_sample(img, ix, iy)
    if is_mask:
        cover = bilinear_sample(torch.ones_like(img), ix, iy)
        cover = torch.where(cover < 0.9999, torch.zeros_like(cover), cover)
        cover = torch.where(cover > 0, torch.ones_like(cover), cover)
        out = out * cover
    return out


# --------------------------------------------------------------------------------------------
# a-6 / a-7  range map + occlusion mask                     models/model.py:223-305, :391
# --------------------------------------------------------------------------------------------


def resize_bilinear(x, size=None, scale_factor=None):
    """F.interpolate(x, size / scale_factor, mode='bilinear', align_corners=True) restated from ATen's upsample_bilinear2d
    (aten/src/ATen/native/UpSample.h: area_pixel_compute_scale, compute_source_index_and_lambda): the glue calls of
    models/networks/cost_volume_flow_net.py:245 and models/model.py:396.  Pinned against F.interpolate itself in
    tests/test_oracle_golden.py."""
    B, C, Hi, Wi = x.shape
    if size is None:
        Ho, Wo = int(Hi * scale_factor), int(Wi * scale_factor)
    else:
        Ho, Wo = (size, size) if isinstance(size, int) else size

    def axis(n_in, n_out):
        r = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        src = torch.arange(n_out, dtype=x.dtype) * torch.tensor(r, dtype=x.dtype)
        i0 = src.floor().long().clamp(max=n_in - 1)
        i1 = i0 + (i0 < n_in - 1).long()
        l1 = src - i0.to(x.dtype)
        return i0, i1, 1.0 - l1, l1
    y0, y1, ly0, ly1 = axis(Hi, Ho)
    x0, x1, lx0, lx1 = axis(Wi, Wo)
    top = x[:, :, y0][:, :, :, x0] * lx0 + x[:, :, y0][:, :, :, x1] * lx1
    bot = x[:, :, y1][:, :, :, x0] * lx0 + x[:, :, y1][:, :, :, x1] * lx1
    return top * ly0[:, None] + bot * ly1[:, None]


def flow_to_warp(flow_bhw2):
    """Endpoints (x+u, y+v) of a [B,H,W,2] flow (models/model.py:223-241)."""
    B, H, W, _ = flow_bhw2.shape
    xs = torch.arange(W, dtype=flow_bhw2.dtype, device=flow_bhw2.device).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, dtype=flow_bhw2.dtype, device=flow_bhw2.device).view(1, H, 1).expand(B, H, W)
    return torch.stack((xs, ys), dim=-1) + flow_bhw2


def range_map(flow):
    """Bilinear forward splat of 1.0 from every pixel to (x+u, y+v); taps outside are dropped
    (models/model.py:243-305).  flow [B,2,H,W] -> [B,1,H,W]."""
    B, _, H, W = flow.shape
    ends = flow_to_warp(flow.permute(0, 2, 3, 1))
    base = torch.floor(ends)
    frac = ends - base
    bx = base[..., 0].to(torch.int64)
    by = base[..., 1].to(torch.int64)
    counts = flow.new_zeros(B * H * W)
    boff = (torch.arange(B, device=flow.device) * (H * W)).view(B, 1, 1)
    for di in (0, 1):
        wx = frac[..., 0] if di else 1.0 - frac[..., 0]
        for dj in (0, 1):
            wy = frac[..., 1] if dj else 1.0 - frac[..., 1]
            tx, ty = bx + di, by + dj
            ok = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
            lin = boff + ty.clamp(0, H - 1) * W + tx.clamp(0, W - 1)
            counts.index_put_((lin.reshape(-1),), (wx * wy * ok.to(flow.dtype)).reshape(-1), accumulate=True)
    return counts.view(B, 1, H, W)


def occlusion_from_range_map(rmap):
    """1 = occluded, 0 = visible (models/model.py:390-391)."""
    return 1.0 - rmap.clamp(0.0, 1.0)


# --------------------------------------------------------------------------------------------
# a-8  Charbonnier / photometric                models/model.py:27-46, utils.py:8-18
# --------------------------------------------------------------------------------------------


def robust_l1(x, alpha=0.001):
    return (x * x + alpha * alpha) ** 0.5


def charbonnier_loss(x, alpha=0.001, reduction=True):
    y = torch.sqrt(x * x + alpha * alpha)
    return y.mean() if reduction else y


def photometric_error(img_pred, img, occ=None):
    err = robust_l1(img_pred - img)
    if occ is None:
        return err.mean()
    vis = 1.0 - occ
    return (err * vis).sum() / (vis.sum() * 3 + 1e-16)


# --------------------------------------------------------------------------------------------
# a-9  smoothness                                           models/model.py:53-114
# --------------------------------------------------------------------------------------------


def gradient(img, stride=1):
    s = int(stride)
    return img[..., :, s:] - img[..., :, :-s], img[..., s:, :] - img[..., :-s, :]


def _edge_weight(g, alpha):
    return torch.exp(-((alpha * g) ** 2).mean(dim=1, keepdim=True))


def first_order_smoothness_loss(img, flow, alpha=100.0):
    igx, igy = gradient(img)
    fgx, fgy = gradient(flow)
    return 0.5 * ((_edge_weight(igx, alpha) * robust_l1(fgx)).mean() + (_edge_weight(igy, alpha) * robust_l1(fgy)).mean())


def second_order_smoothness_loss(img, flow, alpha=100.0):
    igx, igy = gradient(img, stride=2)
    fgx, fgy = gradient(flow)
    fgxx, _ = gradient(fgx)
    _, fgyy = gradient(fgy)
    return 0.5 * ((_edge_weight(igx, alpha) * robust_l1(fgxx)).mean() + (_edge_weight(igy, alpha) * robust_l1(fgyy)).mean())


# edge_aware_smoothness_loss (models/model.py:68-91) is never called by the reference and cannot run as
# written (it adds a [B,H,W-1] tensor to a [B,H-1,W] one), so it is deliberately not restated.


# --------------------------------------------------------------------------------------------
# a-13  supervised losses     flow_model.py:173-186, occlusion_model.py:45-62, flow_occ_model.py:48-55
# --------------------------------------------------------------------------------------------


def binary_cross_entropy(p, target):
    """ATen semantics: log terms clamped at -100."""
    lp = torch.log(p).clamp(min=-100.0)
    l1p = torch.log(1.0 - p).clamp(min=-100.0)
    return -(target * lp + (1.0 - target) * l1p)


def focal_occlusion_loss(occ_pred, occ, gamma=2):
    bce = binary_cross_entropy(occ_pred, occ)
    pt = torch.exp(-bce)
    return ((1.0 - pt) ** gamma * bce).mean()


def flow_occ_supervised_loss(flow_pred, flow, occ_pred, occ):
    return (flow_pred - flow).abs().mean(), binary_cross_entropy(occ_pred, occ).mean()


# --------------------------------------------------------------------------------------------
# a-14  SSIM (metric)                                inpainting_metrics/ssim/ssim.py:7-37
# --------------------------------------------------------------------------------------------


def ssim_window(window_size, sigma=1.5, dtype=torch.float32):
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    g = g / g.sum()
    return (g[:, None] * g[None, :]).to(dtype)


def ssim(img1, img2, window_size=11, size_average=True):
    C = img1.shape[1]
    w = ssim_window(window_size, dtype=img1.dtype).expand(C, 1, window_size, window_size).contiguous()
    p = window_size // 2

    def blur(t):
        return F.conv2d(t, w, padding=p, groups=C)

    mu1, mu2 = blur(img1), blur(img2)
    s11 = blur(img1 * img1) - mu1 * mu1
    s22 = blur(img2 * img2) - mu2 * mu2
    s12 = blur(img1 * img2) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return m.mean() if size_average else m.mean(dim=(1, 2, 3))


# --------------------------------------------------------------------------------------------
# input pipeline        models/data/datasets.py:50-55,157-187 + models/lightning_datamodule.py:20-23
# --------------------------------------------------------------------------------------------


def pack_pairs(img1_u8, img2_u8, flow_hw2=None):
    """uint8 [B,H0,W0,3] frames (+ [B,H0,W0,2] flow) -> ([B,6,H,W] in [-1,1], [B,2,H,W]): centre crop to the largest
    multiple of 64 (datasets.py:148-150, StaticCenterCrop :50-55), ToTensor (/255), Normalize(0.5, 0.5)
    (lightning_datamodule.py:20-23), cat of the two frames (:179), flow transposed (:185)."""
    B, H0, W0, _ = img1_u8.shape
    th, tw = (H0 // 64) * 64, (W0 // 64) * 64
    ya, yb, xa, xb = (H0 - th) // 2, (H0 + th) // 2, (W0 - tw) // 2, (W0 + tw) // 2

    def prep(u8):
        t = u8[:, ya:yb, xa:xb].permute(0, 3, 1, 2).to(torch.float32) / 255.0
        return (t - 0.5) / 0.5

    imgs = torch.cat((prep(img1_u8), prep(img2_u8)), dim=1)
    flow = None if flow_hw2 is None else flow_hw2[:, ya:yb, xa:xb].permute(0, 3, 1, 2).contiguous()
    return imgs, flow


# --------------------------------------------------------------------------------------------
# evaluation metrics                                models/data/utils/flow_utils.py:179-310
# --------------------------------------------------------------------------------------------


def pack_occ(occ_u8):
    """FlyingChairs2 ground-truth occlusion mask, models/data/datasets.py:660-669: decoded uint8 [H0,W0] -> float32 ->
    StaticCenterCrop to a multiple of 64 -> ToTensor (a float array is not rescaled) -> occ[occ > 0.5] = 1 ; occ[occ != 1] = 0.
    occ_u8: uint8 [B,H0,W0] -> float32 [B,1,H,W]."""
    B, H0, W0 = occ_u8.shape
    th, tw = (H0 // 64) * 64, (W0 // 64) * 64
    y0, x0 = (H0 - th) // 2, (W0 - tw) // 2
    occ = occ_u8[:, y0:y0 + th, x0:x0 + tw].to(torch.float32).unsqueeze(1).clone()
    occ[occ > 0.5] = 1.0
    occ[occ != 1.0] = 0.0
    return occ


def flow_error(tu, tv, u, v):
    """Average end-point error (flow_utils.py:179-232, occ=None): ground truth above 1e7 in magnitude marks unknown
    pixels, which are zeroed in all four maps (so they contribute 0 and still count in the mean)."""
    unknown = (tu.abs() > 1e7) | (tv.abs() > 1e7)
    du = torch.where(unknown, torch.zeros_like(tu), tu - u)
    dv = torch.where(unknown, torch.zeros_like(tv), tv - v)
    return torch.sqrt(du * du + dv * dv).mean()


def flow_kitti_error(tu, tv, u, v, mask):
    """(mean EPE over mask != 0, 1 - outlier ratio) with outlier = epe > 3 and epe/(|gt|+1e-5) > 0.05 (flow_utils.py:234-271)."""
    valid = mask != 0
    epe = torch.sqrt((tu - u) ** 2 + (tv - v) ** 2)[valid]
    mag = (torch.sqrt(tu ** 2 + tv ** 2) + 1e-5)[valid]
    err = (epe > 3) & (epe / mag > 0.05)
    return epe.mean(), 1.0 - err.sum().to(epe.dtype) / valid.sum().to(epe.dtype)


# --------------------------------------------------------------------------------------------
# census (ternary) photometric term -- PARITY UNPINNED
#   BASELINE.json's north_star names a census term, but the reference defines none (SURVEY.md section 8a-14:
#   "census: absent").  This restates the published soft census / ternary loss of UnFlow (Meister et al., AAAI 2018)
#   as commonly implemented (7x7 patch, grey*255, t = d/sqrt(0.81+d^2), soft Hamming (dt^2)/(0.1+dt^2), interior mask),
#   combined with the occlusion weighting of photometric_error (models/model.py:37-46).  There is no reference output
#   to compare with: the CUDA kernel is checked against THIS function only.
# --------------------------------------------------------------------------------------------


def _grey255(img):
    if img.shape[1] == 3:
        g = img[:, 0:1] * 0.2989 + img[:, 1:2] * 0.5870 + img[:, 2:3] * 0.1140
    else:
        g = img.mean(dim=1, keepdim=True)
    return g * 255.0


def census_transform(img, max_distance=3):
    """[B,C,H,W] -> [B,(2m+1)^2,H,W]: soft sign of (neighbour - centre) of the grey image, zero padding."""
    m = int(max_distance)
    n = 2 * m + 1
    g = _grey255(img)
    B, _, H, W = g.shape
    gp = F.pad(g, (m, m, m, m))
    planes = [gp[:, :, iy:iy + H, ix:ix + W] - g for iy in range(n) for ix in range(n)]
    t = torch.cat(planes, dim=1)
    return t / torch.sqrt(0.81 + t * t)


def census_distance(img1, img2, max_distance=3):
    """mean over the patch of the soft Hamming distance, zeroed on the m-pixel border -> [B,1,H,W]."""
    m = int(max_distance)
    t1, t2 = census_transform(img1, m), census_transform(img2, m)
    d = (t1 - t2) ** 2
    dist = (d / (0.1 + d)).mean(dim=1, keepdim=True)
    valid = torch.zeros_like(dist)
    H, W = dist.shape[2:]
    if H > 2 * m and W > 2 * m:
        valid[:, :, m:H - m, m:W - m] = 1.0
    return dist * valid, valid


def census_loss(img_pred, img, occ=None, max_distance=3):
    """sum(dist * valid * (1-occ)) / (sum(valid * (1-occ)) + 1e-16); occ [B,1,H,W] as in photometric_error."""
    dist, valid = census_distance(img_pred, img, max_distance)
    w = valid if occ is None else valid * (1.0 - occ)
    return (dist * w).sum() / (w.sum() + 1e-16)


# --------------------------------------------------------------------------------------------
# a-10  FlowNetCV forward, functional over a state_dict    cost_volume_flow_net.py:22-246
# --------------------------------------------------------------------------------------------

_LEVEL_CH = {1: 16, 2: 32, 3: 64, 4: 96, 5: 128, 6: 196}
_ENC_NAMES = {1: ("conv1a", "conv1aa", "conv1b"), 2: ("conv2a", "conv2aa", "conv2b"),
              3: ("conv3a", "conv3aa", "conv3b"), 4: ("conv4a", "conv4aa", "conv4b"),
              5: ("conv5a", "conv5aa", "conv5b"), 6: ("conv6aa", "conv6a", "conv6b")}
_WARP_SCALE = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}


def _conv_lrelu(sd, name, x, stride=1, padding=1, dilation=1):
    y = F.conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], stride=stride, padding=padding, dilation=dilation)
    return F.leaky_relu(y, 0.1)


def flownetcv_forward(sd, x, displacement=4):
    """Returns (flow1 [B,2,H,W] in pixels, flow_l2 [B,2,H/4,W/4] in quarter-res pixels)."""
    pyr = {}
    for which, im in ((1, x[:, :3]), (2, x[:, 3:])):
        t = im
        for lvl in range(1, 7):
            a, b, c = _ENC_NAMES[lvl]
            t = _conv_lrelu(sd, c, _conv_lrelu(sd, b, _conv_lrelu(sd, a, t, stride=2)))
            pyr[(which, lvl)] = t
    up_flow = up_feat = None
    flow = feat = None
    for lvl in (6, 5, 4, 3, 2):
        c1, c2 = pyr[(1, lvl)], pyr[(2, lvl)]
        if lvl < 6:
            c2 = warp(c2, up_flow * _WARP_SCALE[lvl], align_corners=False)
        c1, c2 = normalize_features([c1, c2])
        corr = F.leaky_relu(cost_volume(c1, c2, displacement), 0.1)
        feat = corr if lvl == 6 else torch.cat((corr, c1, up_flow, up_feat), 1)
        for i in range(5):
            feat = torch.cat((_conv_lrelu(sd, "conv%d_%d" % (lvl, i), feat), feat), 1)
        flow = F.conv2d(feat, sd["predict_flow%d.weight" % lvl], sd["predict_flow%d.bias" % lvl], padding=1)
        if lvl > 2:
            up_flow = F.conv_transpose2d(flow, sd["deconv%d.weight" % lvl], sd["deconv%d.bias" % lvl], stride=2, padding=1)
            up_feat = F.conv_transpose2d(feat, sd["upfeat%d.weight" % lvl], sd["upfeat%d.bias" % lvl], stride=2, padding=1)
    t = feat
    for name, dil in (("dc_conv1", 1), ("dc_conv2", 2), ("dc_conv3", 4), ("dc_conv4", 8), ("dc_conv5", 16), ("dc_conv6", 1)):
        t = _conv_lrelu(sd, name, t, padding=dil, dilation=dil)
    flow2 = flow + F.conv2d(t, sd["dc_conv7.weight"], sd["dc_conv7.bias"], padding=1)
    flow1 = resize_bilinear(flow2, scale_factor=4) * 20
    return flow1, flow2 * 5.0


# --------------------------------------------------------------------------------------------
# a-12  occlusion-aware unsupervised step                 models/model.py:366-409, :424
# --------------------------------------------------------------------------------------------


def occ_aware_step(sd, batch, displacement=4):
    """(photo, smooth1, smooth2, flow_error, photo_occ, occ_error) for model=='pwc'."""
    imgs, flow_gt, occ_gt = batch
    img1, img2 = imgs[:, 0:3], imgs[:, 3:6]
    flow_pred, flow_l2 = flownetcv_forward(sd, imgs, displacement)
    img_warped = warp(img2, flow_pred, align_corners=True)
    with torch.no_grad():
        back_flow, _ = flownetcv_forward(sd, torch.cat((img2, img1), 1), displacement)
        occ_pred = occlusion_from_range_map(range_map(back_flow))
    photo = photometric_error(img_warped, img1, occ_pred)
    img1_l2 = resize_bilinear(img1, scale_factor=0.25)
    smooth1 = first_order_smoothness_loss(img1_l2, flow_l2)
    smooth2 = second_order_smoothness_loss(img1_l2, flow_l2)
    flow_error = ((flow_pred - flow_gt) ** 2).mean()
    photo_occ = photometric_error(img_warped, img1, 1.0 - occ_pred)
    occ_error = binary_cross_entropy(occ_gt, occ_pred).mean()  # arguments swapped as in models/model.py:407
    return photo, smooth1, smooth2, flow_error, photo_occ, occ_error


def total_loss(losses, photo_weight=4.0, smooth1_weight=0.5, smooth2_weight=0.0):
    """models/model.py:424 with the shipped config/unsupervised_config.yml weights as defaults."""
    return photo_weight * losses[0] + smooth1_weight * losses[1] + smooth2_weight * losses[2]


def deterministic_state_dict(shapes, seed=0, dtype=torch.float32, flow_gain=1.0):
    """Name-keyed deterministic weights (independent of module construction order) used by the golden
    fixtures and the GPU parity tests: fan-in-scaled normal weights, small normal biases.

    flow_gain scales the flow-predicting layers (predict_flow*, dc_conv7).  With gain 1 random weights
    give ~50 px flows on a 64 px image; the step's gradient is then discontinuous enough (bilinear cell
    changes, frame exits) that a 1e-6 relative weight perturbation moves parameter gradients by 1e-2
    in the reference itself -- no implementation can be compared at 1e-3 there.  The model-level
    fixtures use a small gain so the comparison is well conditioned."""
    sd = {}
    for i, (name, shape) in enumerate(sorted(shapes.items())):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if len(shape) > 1:
            fan = 1
            for s in shape[1:]:
                fan *= s
            sd[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * (1.0 / math.sqrt(fan))).to(dtype)
        else:
            sd[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * 0.05).to(dtype)
        if flow_gain != 1.0 and (name.startswith("predict_flow") or name.startswith("dc_conv7")):
            sd[name] = sd[name] * flow_gain
    return sd
