"""TEST INFRASTRUCTURE ONLY -- tests/golden/metrics.pt from the REAL reference's numpy metrics
(models/data/utils/flow_utils.py: evaluate_flow :289-296 -> flow_error :179-232, evaluate_kitti_flow :299-310 ->
flow_kitti_error :234-271).  Run in the build container:  python oracle/make_golden_metrics.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    ref_loader.load()  # puts the reference root on sys.path
    FU = importlib.import_module("models.data.utils.flow_utils")
    rng = np.random.default_rng(11)
    cases = []
    for H, W, scale in ((5, 7, 3.0), (33, 41, 8.0), (64, 96, 1.0)):
        gt = (rng.standard_normal((H, W, 2)) * scale).astype(np.float32)
        pred = (gt + rng.standard_normal((H, W, 2)) * 2.0).astype(np.float32)
        gt[0, 0, 0] = 1e9                      # an "unknown" ground-truth pixel (flow_utils.py:201-206)
        gt[H // 2, W // 3, 1] = -2e8
        mask = (rng.random((H, W)) < 0.7).astype(np.float32)
        gt3 = np.concatenate((gt, mask[:, :, None]), 2)
        gt3[0, 0, 0] = 1.5
        gt3[H // 2, W // 3, 1] = -0.5
        c = dict(gt=torch.from_numpy(gt.copy()), pred=torch.from_numpy(pred.copy()), gt3=torch.from_numpy(gt3.copy()))
        # the reference functions modify their (view) arguments in place: hand them copies
        c["ref_epe"] = float(FU.evaluate_flow(gt.copy(), pred.copy()))
        e2, a2 = FU.evaluate_kitti_flow(gt3[:, :, :2].copy(), pred.copy())
        e3, a3 = FU.evaluate_kitti_flow(gt3.copy(), pred.copy())
        c["ref_kitti2"] = (float(e2), float(a2))
        c["ref_kitti3"] = (float(e3), float(a3))
        cases.append(c)
    out = os.path.join(ROOT, "tests", "golden", "metrics.pt")
    torch.save(cases, out)
    print(out, os.path.getsize(out), [(c["ref_epe"], c["ref_kitti3"]) for c in cases])


if __name__ == "__main__":
    main()
