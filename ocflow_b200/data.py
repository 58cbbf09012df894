"""On-device input pipeline: decoded uint8 frames (+ .flo flow) -> network input, on the GPU.

Mirrors what the reference does per sample on the host (models/data/datasets.py:157-187 with the transform of
models/lightning_datamodule.py:20-23): centre crop to a multiple of 64, ToTensor, Normalize(0.5, 0.5), cat of the two
frames, flow transposed to [2,H,W] -- as one kernel over the whole batch.  The frames cross PCIe as uint8.
"""
import torch

from . import _lib
from .ops import _p, _stream


def center_crop_origin(image_size, crop_size):
    """StaticCenterCrop (models/data/datasets.py:50-55): origin of img[(h-th)//2:(h+th)//2, (w-tw)//2:(w+tw)//2]."""
    (h, w), (th, tw) = image_size, crop_size
    return (h - th) // 2, (w - tw) // 2


def render_size(image_size):
    """The reference crops to the largest multiple of 64 (models/data/datasets.py:148-150)."""
    return (image_size[0] // 64) * 64, (image_size[1] // 64) * 64


def pack_pairs(img1_u8, img2_u8, flow_hw2=None, crop_size=None, origin=None):
    """img1_u8, img2_u8: CUDA uint8 [B,H0,W0,3]; flow_hw2: optional CUDA fp32 [B,H0,W0,2].
    Returns (imgs [B,6,H,W] in [-1,1], flow [B,2,H,W] or None) for the crop (default: the reference's centre crop to a
    multiple of 64)."""
    for name, t in (("img1_u8", img1_u8), ("img2_u8", img2_u8)):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise TypeError("%s must be a CUDA tensor: ocflow_b200 has no CPU path" % name)
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
            raise TypeError("%s must be uint8 [B,H,W,3] (got %s %s)" % (name, t.dtype, tuple(t.shape)))
    if img1_u8.shape != img2_u8.shape:
        raise ValueError("the two frames must have the same shape")
    img1_u8, img2_u8 = img1_u8.contiguous(), img2_u8.contiguous()
    B, H0, W0, _ = img1_u8.shape
    H, W = crop_size if crop_size is not None else render_size((H0, W0))
    y0, x0 = origin if origin is not None else center_crop_origin((H0, W0), (H, W))
    if flow_hw2 is not None:
        if not flow_hw2.is_cuda or flow_hw2.dtype != torch.float32 or tuple(flow_hw2.shape) != (B, H0, W0, 2):
            raise TypeError("flow_hw2 must be CUDA float32 [B,H0,W0,2]")
        flow_hw2 = flow_hw2.contiguous()
    imgs = torch.empty((B, 6, H, W), device=img1_u8.device, dtype=torch.float32)
    flow = torch.empty((B, 2, H, W), device=img1_u8.device, dtype=torch.float32) if flow_hw2 is not None else None
    with torch.cuda.device_of(img1_u8):
        _lib.call("ocf_pack_pairs", _p(img1_u8), _p(img2_u8), _p(flow_hw2), _p(imgs), _p(flow), B, H0, W0, H, W, int(y0), int(x0), _stream())
    return imgs, flow


def pack_occ(occ_u8, crop_size=None, origin=None):
    """occ_u8: CUDA uint8 [B,H0,W0] (the decoded FlyingChairs2 `*-occ_01.png`).  Returns the [B,1,H,W] fp32 {0,1} mask of
    models/data/datasets.py:660-669 for the crop (default: the reference's centre crop to a multiple of 64)."""
    if not isinstance(occ_u8, torch.Tensor) or not occ_u8.is_cuda:
        raise TypeError("occ_u8 must be a CUDA tensor: ocflow_b200 has no CPU path")
    if occ_u8.dtype != torch.uint8 or occ_u8.dim() != 3:
        raise TypeError("occ_u8 must be uint8 [B,H,W] (got %s %s)" % (occ_u8.dtype, tuple(occ_u8.shape)))
    occ_u8 = occ_u8.contiguous()
    B, H0, W0 = occ_u8.shape
    H, W = crop_size if crop_size is not None else render_size((H0, W0))
    y0, x0 = origin if origin is not None else center_crop_origin((H0, W0), (H, W))
    occ = torch.empty((B, 1, H, W), device=occ_u8.device, dtype=torch.float32)
    with torch.cuda.device_of(occ_u8):
        _lib.call("ocf_pack_occ", _p(occ_u8), _p(occ), B, H0, W0, H, W, int(y0), int(x0), _stream())
    return occ
