"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (the installed reference on the host cores) prints
exactly one JSON line with the agreed keys, ranks > 0 stay silent, and the product arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")
TINY = ["--height", "64", "--width", "64", "--steps", "1", "--warmup", "0", "--batch", "1"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference"] + TINY)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("image pairs/sec") and d["value"] > 0 and d["steps"] == 1
    from oracle import ref_loader
    want_kind = "reference" if ref_loader.available() else "port"   # the real reference whenever baseline/_ref is installed
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["same_config_as_ours"] is True and d["config"]["global_batch"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["dtype"] == "f32"


def test_reference_arm_other_ranks_exit_silently():
    p = _run(["--impl", "reference", "--gpus", "2"] + TINY, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    p = _run(TINY)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
