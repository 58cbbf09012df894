"""CPU: `.flo` reader / writer against the REAL reference's read_flow / save_flow where the reference tree is present
(files written by one are read by the other, byte-identical output), and a self-contained round trip everywhere."""
import importlib
import os

import numpy as np
import pytest

from ocflow_b200 import flow_io
from oracle import ref_loader


def test_flo_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    flow = (rng.standard_normal((13, 17, 2)) * 9).astype(np.float32)
    p = str(tmp_path / "a.flo")
    flow_io.save_flow(p, flow)
    assert os.path.getsize(p) == 12 + 13 * 17 * 2 * 4
    back = flow_io.read_flow(p)
    assert back.shape == (13, 17, 2) and back.dtype == np.float32 and np.array_equal(back, flow)
    flow_io.save_flow(p, flow[:, :, 0], flow[:, :, 1])          # separate u, v
    assert np.array_equal(flow_io.read_flow(p), flow)
    t = flow_io.read_flow_pinned(p)
    assert tuple(t.shape) == (13, 17, 2) and np.array_equal(t.numpy(), flow)
    with open(p, "wb") as f:
        f.write(b"\x00" * 40)
    assert flow_io.read_flow(p) is None                          # bad magic: None, like the reference


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_flo_files_interoperate_with_the_real_reference(tmp_path):
    ref_loader.load()
    FU = importlib.import_module("models.data.utils.flow_utils")
    rng = np.random.default_rng(1)
    flow = (rng.standard_normal((9, 21, 2)) * 4).astype(np.float32)
    ours, theirs = str(tmp_path / "ours.flo"), str(tmp_path / "theirs.flo")
    flow_io.save_flow(ours, flow)
    FU.save_flow(theirs, flow)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    assert np.array_equal(FU.read_flow(ours), flow) and np.array_equal(flow_io.read_flow(theirs), flow)
    assert np.array_equal(flow_io.read_flow(theirs), FU.read_flow(theirs))
