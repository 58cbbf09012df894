"""Middlebury `.flo` files -- host-side mirror of read_flow / save_flow (models/data/utils/flow_utils.py:10-63).

Format: float32 tag 202021.25, int32 width, int32 height, then height x width x 2 float32 (u, v interleaved), little endian.
`read_flow` returns the reference's [h, w, 2] numpy array (None with the same message on a bad tag); `read_flow_pinned`
returns the same data as a pinned torch tensor ready for a non-blocking H2D copy into `ocflow_b200.data.pack_pairs`.
"""
import numpy as np
import torch

FLO_TAG = 202021.25


def read_flow(filename):
    with open(filename, "rb") as f:
        raw = f.read()
    if len(raw) < 12 or np.frombuffer(raw, np.float32, 1, 0)[0] != np.float32(FLO_TAG):
        print("Magic number incorrect. Invalid .flo file")
        return None
    w, h = (int(v) for v in np.frombuffer(raw, np.int32, 2, 4))
    data = np.frombuffer(raw, np.float32, 2 * w * h, 12)
    return data.reshape(h, w, 2).copy()


def read_flow_pinned(filename):
    flow = read_flow(filename)
    if flow is None:
        return None
    t = torch.from_numpy(flow)
    return t.pin_memory() if torch.cuda.is_available() else t


def save_flow(filename, uv, v=None):
    """uv: [h, w, 2] when v is None, else the [h, w] u component with v given separately (flow_utils.py:31-63)."""
    if v is None:
        uv = np.asarray(uv)
        if uv.ndim != 3 or uv.shape[2] != 2:
            raise AssertionError("uv must be [h, w, 2]")
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(uv), np.asarray(v)
    if u.shape != v.shape:
        raise AssertionError("u and v must have the same shape")
    h, w = u.shape
    with open(filename, "wb") as f:
        np.array([FLO_TAG], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        np.stack((u, v), axis=2).astype(np.float32).tofile(f)
