"""In-situ parity harness: record every hot-path op call made during one REAL training step (inputs, outputs, the
cotangent that reaches each output and the gradient each op sends to each input) and re-evaluate every call with the
fp64 oracle on exactly those tensors.  This separates kernel parity from the conditioning of the whole network
(tests/test_gpu_model.py explains why end-to-end parameter gradients cannot be compared at 1e-4)."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import ocflow_oracle as O


def _iso(t):
    # a fresh autograd node whose gradient is THIS op's contribution only (the tensor may have other consumers)
    return t.view_as(t) if isinstance(t, torch.Tensor) and t.requires_grad else t


@contextlib.contextmanager
def recording():
    from ocflow_b200 import ops

    calls = []
    orig = dict(cost_volume=ops.cost_volume, normalize_features=ops.normalize_features, warp=ops.warp,
                occ_photo_fused=ops.occ_photo_fused, smoothness_loss=ops.smoothness_loss, range_map=ops.range_map)

    def cv(f1, f2, max_displacement=4, leaky_slope=1.0):
        f1, f2 = _iso(f1), _iso(f2)
        out = orig["cost_volume"](f1, f2, max_displacement, leaky_slope)

        def orc(a, b):
            r = O.cost_volume(a, b, max_displacement)
            return [F.leaky_relu(r, leaky_slope) if leaky_slope != 1.0 else r]
        calls.append(("corr", [f1, f2], [out], orc))
        return out

    def nf(fl, **kw):
        fl = [_iso(t) for t in fl]
        out = orig["normalize_features"](fl, **kw)
        calls.append(("normalize", fl, list(out), lambda *a: O.normalize_features(list(a), **kw)))
        return out

    def wp(img, flow, align_corners=True, is_mask=False, occ=None, flow_scale=1.0):
        img, flow = _iso(img), _iso(flow)
        out = orig["warp"](img, flow, align_corners, is_mask, occ, flow_scale)
        if occ is None:
            calls.append(("warp ac=%d" % align_corners, [img, flow], [out], lambda a, f: [O.warp(a, f * flow_scale, align_corners, is_mask)]))
        return out

    def opf(img1, img2, flow, rmap=None, flow_gt=None, occ_gt=None, alpha=0.001):
        flow = _iso(flow)
        out = orig["occ_photo_fused"](img1, img2, flow, rmap, flow_gt, occ_gt, alpha)

        def orc(f):
            dt = f.dtype
            i1, i2 = img1.detach().to(dt).cpu(), img2.detach().to(dt).cpu()
            occ = O.occlusion_from_range_map(rmap.to(dt).cpu()) if rmap is not None else torch.zeros_like(i1[:, :1])
            w = O.warp(i2, f, True)
            res = [O.photometric_error(w, i1, occ), O.photometric_error(w, i1, 1.0 - occ)]
            if flow_gt is not None:
                res.append(((f - flow_gt.to(dt).cpu()) ** 2).mean())
            if occ_gt is not None:
                res.append(O.binary_cross_entropy(occ_gt.to(dt).cpu(), occ).mean())
            return res
        outs = [out[0], out[1]] + ([out[2]] if flow_gt is not None else []) + ([out[3]] if occ_gt is not None else [])
        # fp32 oracle here: the fp32 REFERENCE's d/dflow is itself 2e-4..5e-4 away from fp64 (the normalise /
        # un-normalise round trip of the sampling grid costs ~2e-6 px), so fp64 would test the reference, not the kernel
        calls.append(("occ_photo_fused", [flow], outs, orc, torch.float32))
        return out

    def sm(img, flow, order, alpha=100.0, alpha_rho=0.001):
        img, flow = _iso(img), _iso(flow)
        out = orig["smoothness_loss"](img, flow, order, alpha, alpha_rho)
        fn = O.first_order_smoothness_loss if order == 1 else O.second_order_smoothness_loss
        calls.append(("smooth%d" % order, [img, flow], [out], lambda a, f: [fn(a, f, alpha)]))
        return out

    def rm(flow, with_occlusion=False):
        out = orig["range_map"](flow, with_occlusion)
        calls.append(("range_map", [flow.detach()], [out[0] if with_occlusion else out], lambda f: [O.range_map(f)]))
        return out

    def lf(c1, c2, up_flow=None, up_feat=None, flow_scale=1.0, leaky_slope=0.1):
        ins = [_iso(c1), _iso(c2)] + ([_iso(up_flow), _iso(up_feat)] if up_flow is not None else [])
        out = orig["level_fused"](*ins, flow_scale=flow_scale, leaky_slope=leaky_slope) if up_flow is not None else \
            orig["level_fused"](ins[0], ins[1], None, None, flow_scale, leaky_slope)

        def orc(a, b, fl=None, ft=None):   # cost_volume_flow_net.py:186-190 as the reference composes it
            if fl is not None:
                b = O.warp(b, fl * flow_scale, False)
            an, bn = O.normalize_features([a, b])
            corr = F.leaky_relu(O.cost_volume(an, bn, 4), leaky_slope)
            return [corr if fl is None else torch.cat((corr, an, fl, ft), 1)]
        calls.append(("level_fused", ins, [out], orc))
        return out

    orig["level_fused"] = ops.level_fused
    ops.cost_volume, ops.normalize_features, ops.warp, ops.occ_photo_fused, ops.smoothness_loss, ops.range_map = cv, nf, wp, opf, sm, rm
    ops.level_fused = lf
    try:
        yield calls
    finally:
        for k, v in orig.items():
            setattr(ops, k, v)


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(calls, loss, report=None):
    """Backpropagate `loss`, then compare every recorded call with the fp64 oracle.  Returns a list of
    (index, name, shape, [out errors], [grad errors]) -- errors are max|d|/max|ref|."""
    gin, gout = {}, {}
    for ci, call in enumerate(calls):
        name, ins, outs = call[0], call[1], call[2]
        for j, t in enumerate(ins):
            if t.requires_grad:
                t.register_hook(lambda g, k=(ci, j): gin.__setitem__(k, g.detach().clone()))
        for j, t in enumerate(outs):
            if t.requires_grad:
                t.register_hook(lambda g, k=(ci, j): gout.__setitem__(k, g.detach().clone()))
    loss.backward()
    rows = []
    for ci, call in enumerate(calls):
        name, ins, outs, orc = call[:4]
        dtype = call[4] if len(call) > 4 else torch.float64
        leaves = [t.detach().to(dtype).cpu().requires_grad_(t.requires_grad) for t in ins]
        ro = orc(*leaves)
        oerr = [rel_max(o, r) for o, r in zip(outs, ro)]
        gerr = []
        cots = [(j, gout[(ci, j)]) for j in range(len(outs)) if (ci, j) in gout]
        if cots and any(l.requires_grad for l in leaves):
            tot = sum((ro[j] * g.to(dtype).cpu()).sum() for j, g in cots)
            gr = torch.autograd.grad(tot, [l for l in leaves if l.requires_grad], allow_unused=True)
            it = iter(gr)
            for j, l in enumerate(leaves):
                if l.requires_grad:
                    g = next(it)
                    if g is not None and (ci, j) in gin and float(g.abs().max()) > 0:
                        gerr.append(rel_max(gin[(ci, j)], g))
        rows.append((ci, name, tuple(ins[0].shape), oerr, gerr))
        if report is not None:
            report("%2d %-16s %-18s out %s  grads %s" % (ci, name, tuple(ins[0].shape), " ".join("%.1e" % e for e in oerr),
                                                         " ".join("%.1e" % e for e in gerr)))
    return rows
