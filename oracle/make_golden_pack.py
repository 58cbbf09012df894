"""TEST INFRASTRUCTURE ONLY -- tests/golden/pack.pt: the reference's per-sample host pipeline (StaticCenterCrop of
models/data/datasets.py:50-55 + the torchvision transform of models/lightning_datamodule.py:20-23 + cat / transpose of
datasets.py:179-186) run on seeded uint8 frames.  Run in the build container:  python oracle/make_golden_pack.py
"""
import importlib
import os
import sys

import numpy as np
import torch
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    ref_loader.load()
    # imageio is absent here (ref_loader installs an empty stub) and never called on this path: give the stub the one name
    # models/data/utils/frame_utils.py imports so that models.data.datasets can be imported
    if not hasattr(sys.modules.get("imageio"), "imread"):
        def _no_imread(*a, **k):
            raise RuntimeError("imageio is not installed in the build container")
        sys.modules["imageio"].imread = _no_imread
    DS = importlib.import_module("models.data.datasets")
    transform = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])  # lightning_datamodule.py:20-23
    rng = np.random.default_rng(3)
    cases = []
    for B, H0, W0 in ((2, 100, 140), (1, 436, 1024), (3, 64, 64)):
        i1 = rng.integers(0, 256, (B, H0, W0, 3), dtype=np.uint8)
        i2 = rng.integers(0, 256, (B, H0, W0, 3), dtype=np.uint8)
        fl = (rng.standard_normal((B, H0, W0, 2)) * 4).astype(np.float32)
        th, tw = (H0 // 64) * 64, (W0 // 64) * 64                       # datasets.py:148-150
        imgs, flows = [], []
        for b in range(B):
            cropper = DS.StaticCenterCrop((H0, W0), (th, tw))            # datasets.py:164
            a, c = transform(cropper(i1[b])), transform(cropper(i2[b]))  # :165-174
            imgs.append(torch.cat((a, c)))                               # :179
            flows.append(torch.from_numpy(cropper(fl[b]).transpose(2, 0, 1).copy()))  # :182-186
        cases.append(dict(img1=torch.from_numpy(i1), img2=torch.from_numpy(i2), flow=torch.from_numpy(fl),
                          ref_imgs=torch.stack(imgs), ref_flow=torch.stack(flows)))
    # keep the fixture small: store the big case's outputs as a checksum only
    big = cases[1]
    big["ref_imgs_sum"] = float(big["ref_imgs"].double().sum())
    big["ref_imgs_abs_sum"] = float(big["ref_imgs"].double().abs().sum())
    big["ref_flow_sum"] = float(big["ref_flow"].double().sum())
    big["seed_note"] = "inputs regenerated from np.random.default_rng(3) in the same draw order"
    for k in ("img1", "img2", "flow", "ref_imgs", "ref_flow"):
        del big[k]
    out = os.path.join(ROOT, "tests", "golden", "pack.pt")
    torch.save(cases, out)
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()
