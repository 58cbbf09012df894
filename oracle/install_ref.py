"""TEST / BENCH INFRASTRUCTURE ONLY -- installs the UNMODIFIED OCFlow reference into baseline/_ref/.

    python -m oracle.install_ref            (also run by __graft_entry__.build() when /root/reference is present)

The reference is a pure-Python tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build; this recipe is the equivalent: it copies the reference's Python sources
byte for byte (no edits, verified by a SHA-256 manifest written next to them) into `baseline/_ref/`, which is
git-ignored (the reference never enters this repository's history) but NOT gpurun-ignored, so the real reference
travels to the GPU box with the snapshot.  There it serves three purposes, none of them on the product path:

  * `bench.py --impl reference`: the reference's own `FlowStageModel('pwc', occ_aware).general_step_occ_aware`
    + weighted loss + backward + Adam on the box's host cores;
  * `-m gpu` tests that run `ocflow_b200.patch.patch_reference()` on the real reference networks on CUDA and compare
    them with the unpatched reference (torch's generic CUDA kernels) on the same weights and inputs;
  * the like-for-like GPU baseline `bench.py` reports as `reference_on_gpu`.

Only `*.py` files are copied (models/, inpainting_metrics/, utils.py; the entry scripts and YAML configs are not
needed).  The two third-party modules the reference imports but this image lacks (pytorch_lightning, matplotlib)
are stubbed at import time by oracle/ref_loader.py, not here: the installed tree stays identical to upstream.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SRC_CANDIDATES = (os.environ.get("OCFLOW_REF_SRC"), "/root/reference")
KEEP_DIRS = ("models", "inpainting_metrics")
KEEP_FILES = ("utils.py",)


def source_root():
    for cand in SRC_CANDIDATES:
        if cand and os.path.isfile(os.path.join(cand, "models", "networks", "correlation_layer.py")):
            return cand
    return None


def _walk(src):
    for d in KEEP_DIRS:
        for dirpath, _, files in os.walk(os.path.join(src, d)):
            for fn in sorted(files):
                if fn.endswith(".py"):
                    yield os.path.relpath(os.path.join(dirpath, fn), src)
    for fn in KEEP_FILES:
        if os.path.isfile(os.path.join(src, fn)):
            yield fn


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def installed():
    return os.path.isfile(os.path.join(DEST, "models", "networks", "correlation_layer.py")) and \
        os.path.isfile(os.path.join(DEST, "MANIFEST.json"))


def verify():
    """True when every installed file still has the hash recorded at install time (i.e. is unmodified)."""
    if not installed():
        return False
    man = json.load(open(os.path.join(DEST, "MANIFEST.json")))
    return all(os.path.isfile(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h for rel, h in man["files"].items())


def install(force=False):
    """Copy the reference into baseline/_ref/.  Returns the destination, or None when no source tree is present
    (the GPU box: it uses the tree that travelled with the snapshot)."""
    src = source_root()
    if src is None:
        return DEST if installed() else None
    rels = list(_walk(src))
    if installed() and not force:
        man = json.load(open(os.path.join(DEST, "MANIFEST.json")))
        if man["files"] == {rel: _sha(os.path.join(src, rel)) for rel in rels} and verify():
            return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in rels:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest, "note": "unmodified copy of the reference's Python sources"}, fh, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    out = install(force="--force" in sys.argv)
    print(out if out else "no reference source tree found and nothing installed")
