"""TEST INFRASTRUCTURE ONLY -- loader for the real OCFlow reference (never shipped, never on the product path).

Imports the unmodified reference from ``$OCFLOW_REF``, ``baseline/_ref`` (the git-ignored install made by
``oracle/install_ref.py``; it travels to the GPU box with the snapshot) or ``/root/reference`` (build container
only) so that ``oracle/make_golden.py``, the "oracle vs real reference" pin tests, the patched-vs-unpatched GPU
tests and ``bench.py --impl reference`` can call the reference's own functions.

Two third-party modules the reference imports are absent from this image (and from the offline
wheelhouse): ``pytorch_lightning`` and ``matplotlib``.  They are replaced by 2 inert stubs
(SURVEY.md section 8c): ``LightningModule`` becomes ``torch.nn.Module`` and ``matplotlib.pyplot``
an empty module.  Neither is touched by the hot path.
"""
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.environ.get("OCFLOW_REF"), os.path.join(_ROOT, "baseline", "_ref"), "/root/reference")


def reference_root():
    for cand in _CANDIDATES:
        if cand and os.path.isfile(os.path.join(cand, "models", "networks", "correlation_layer.py")):
            return cand
    return None


def available():
    return reference_root() is not None


def _install_stubs():
    import torch.nn as nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class _LightningModule(nn.Module):
            # the reference does `self.hparams = hparams` (models/model.py:161)
            pass

        pl.LightningModule = _LightningModule
        pl.LightningDataModule = object
        pl.seed_everything = lambda seed: None
        sys.modules["pytorch_lightning"] = pl
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "imageio" not in sys.modules:
        try:
            import imageio  # noqa: F401
        except Exception:
            sys.modules["imageio"] = types.ModuleType("imageio")


class Ref:
    """Namespace with the reference's hot-path symbols."""


_cached = None


def load():
    """Return a namespace of the REAL reference symbols, or raise RuntimeError when absent."""
    global _cached
    if _cached is not None:
        return _cached
    root = reference_root()
    if root is None:
        raise RuntimeError("OCFlow reference tree not found (set $OCFLOW_REF); use tests/golden fixtures")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    ref = Ref()
    ref.root = root
    ref.correlation_layer = importlib.import_module("models.networks.correlation_layer")
    ref.cost_volume_flow_net = importlib.import_module("models.networks.cost_volume_flow_net")
    ref.pwc_net = importlib.import_module("models.networks.pwc_net")
    ref.flow_net = importlib.import_module("models.networks.flow_net")
    ref.model = importlib.import_module("models.model")
    ref.flow_model = importlib.import_module("models.flow_model")
    ref.utils = importlib.import_module("utils")
    ref.ssim = importlib.import_module("inpainting_metrics.ssim.ssim")
    _cached = ref
    return ref
