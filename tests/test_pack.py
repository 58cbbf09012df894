"""On-device input pipeline: oracle vs the fixture produced by the REAL reference classes
(StaticCenterCrop + the torchvision transform of the datamodule; tests/golden/pack.pt, oracle/make_golden_pack.py), and the
CUDA kernel vs both -- bit-exact (integer -> fp32 with torchvision's op order)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_golden
from oracle import ocflow_oracle as O

CASES = load_golden(os.path.join(GOLD, "pack.pt"))


def _regenerated_inputs():
    """The generator's draws, in order (the large Sintel-size case is stored as checksums only)."""
    rng = np.random.default_rng(3)
    out = []
    for B, H0, W0 in ((2, 100, 140), (1, 436, 1024), (3, 64, 64)):
        i1 = rng.integers(0, 256, (B, H0, W0, 3), dtype=np.uint8)
        i2 = rng.integers(0, 256, (B, H0, W0, 3), dtype=np.uint8)
        fl = (rng.standard_normal((B, H0, W0, 2)) * 4).astype(np.float32)
        out.append((torch.from_numpy(i1), torch.from_numpy(i2), torch.from_numpy(fl)))
    return out


def _check(fn, to_dev):
    inputs = _regenerated_inputs()
    for c, (i1, i2, fl) in zip(CASES, inputs):
        if "img1" in c:
            assert torch.equal(c["img1"], i1) and torch.equal(c["flow"], fl)   # the replayed draws are the stored inputs
        imgs, flow = fn(to_dev(i1), to_dev(i2), to_dev(fl))
        imgs, flow = imgs.cpu(), flow.cpu()
        if "ref_imgs" in c:
            assert torch.equal(imgs, c["ref_imgs"]), float((imgs - c["ref_imgs"]).abs().max())
            assert torch.equal(flow, c["ref_flow"])
        else:
            assert imgs.shape == (1, 6, 384, 1024) and flow.shape == (1, 2, 384, 1024)
            assert float(imgs.double().sum()) == c["ref_imgs_sum"]
            assert float(imgs.double().abs().sum()) == c["ref_imgs_abs_sum"]
            assert float(flow.double().sum()) == c["ref_flow_sum"]


def test_oracle_pack_matches_reference_fixture():
    _check(O.pack_pairs, lambda t: t)


@pytest.mark.gpu
def test_cuda_pack_matches_reference_fixture_bit_exact():
    from ocflow_b200 import data

    _check(data.pack_pairs, lambda t: t.cuda())
    # images only, explicit crop window, and argument checking
    i1 = torch.randint(0, 256, (2, 70, 90, 3), dtype=torch.uint8)
    imgs, flow = data.pack_pairs(i1.cuda(), i1.cuda(), None, crop_size=(64, 64), origin=(3, 20))
    assert flow is None
    want = ((i1[:, 3:67, 20:84].permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    assert torch.equal(imgs[:, :3].cpu(), want) and torch.equal(imgs[:, 3:].cpu(), want)
    with pytest.raises(TypeError):
        data.pack_pairs(i1, i1)
    with pytest.raises(RuntimeError):
        data.pack_pairs(i1.cuda(), i1.cuda(), None, crop_size=(64, 64), origin=(10, 40))   # window leaves the frame


def _reference_occ(occ_np):
    """The reference's own statements for one sample (models/data/datasets.py:660-669) with its StaticCenterCrop and
    torchvision's ToTensor, imported from the installed reference when it is present."""
    import importlib
    import sys

    from oracle import ref_loader
    from torchvision import transforms

    ref_loader.load()
    if not hasattr(sys.modules.get("imageio"), "imread"):      # empty stub installed by ref_loader; never called on this path
        sys.modules["imageio"].imread = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("imageio is not installed"))
    ds = importlib.import_module("models.data.datasets")

    h, w = occ_np.shape
    cropper = ds.StaticCenterCrop((h, w), [(h // 64) * 64, (w // 64) * 64])
    occ = cropper(occ_np.astype(np.float32)[:, :, None])     # StaticCenterCrop indexes three axes: a decoded mask is [H,W,1]
    occ = transforms.ToTensor()(occ)
    occ[occ > 0.5] = 1.0
    occ[occ != 1.0] = 0.0
    return occ


def test_oracle_pack_occ_matches_the_reference_statements():
    from oracle import ref_loader

    rng = np.random.default_rng(11)
    occ = (rng.random((2, 100, 140)) < 0.3).astype(np.uint8) * 255
    occ[0, 20:30, 40:50] = 1          # any non-zero decoded value is "occluded"
    mine = O.pack_occ(torch.from_numpy(occ))
    assert mine.shape == (2, 1, 64, 128) and set(mine.unique().tolist()) <= {0.0, 1.0}
    if ref_loader.available():
        for b in range(2):
            assert torch.equal(mine[b], _reference_occ(occ[b]))


@pytest.mark.gpu
def test_cuda_pack_occ_bit_exact():
    from ocflow_b200 import data

    rng = np.random.default_rng(12)
    occ = torch.from_numpy((rng.integers(0, 4, (3, 436, 1024)) == 0).astype(np.uint8) * rng.integers(1, 256, (3, 436, 1024), dtype=np.uint8))
    assert torch.equal(data.pack_occ(occ.cuda()).cpu(), O.pack_occ(occ))
    win = data.pack_occ(occ.cuda(), crop_size=(64, 128), origin=(5, 7)).cpu()
    assert torch.equal(win, (occ[:, 5:69, 7:135] > 0).float().unsqueeze(1))
    with pytest.raises(TypeError):
        data.pack_occ(occ)
