"""Developer aid: builds a -DOCF_TIMELINE copy of the library, runs one kernel of bench.kernel_table and prints the per-CTA
globaltimer stamps (ns relative to the earliest CTA start).  usage: python tools/timeline.py corr_fwd_L2"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocflow_b200 import build as B
dbg = os.path.join(ROOT, "tools", "bin", "libocflow_tl.so")
os.makedirs(os.path.dirname(dbg), exist_ok=True)
if "--nobuild" not in sys.argv:
    cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-DOCF_TIMELINE"] + [a for a in sys.argv[2:] if a.startswith("-D")] + ["-o", dbg] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
    subprocess.check_call(cmd)
if "--buildonly" in sys.argv:
    sys.exit(0)
import torch
import argparse
from ocflow_b200 import _lib
_lib.LIB_PATH = dbg          # load the instrumented library in place of the product one (before the first load())
lib = _lib.load()
import bench
a = argparse.Namespace(batch=8, height=384, width=512, levels="")
name = sys.argv[1]
table = bench.kernel_table(a, torch)
fn = table[name][0]
flush = torch.empty(256 * 1024 * 1024, device="cuda")
for _ in range(3):
    flush.zero_(); fn()
torch.cuda.synchronize()
n = 1024 * 16
buf = (ctypes.c_ulonglong * n)()
lib.ocf_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.ocf_debug_timeline(buf, n)
assert rc == 0, rc
rows = [list(buf[i * 16:(i + 1) * 16]) for i in range(1024)]
rows = [r for r in rows if r[0]]
t0 = min(r[0] for r in rows)
print("ctas", len(rows), "span_us %.2f" % ((max(r[14] for r in rows) - t0) / 1e3))
def rel(v): return "%7.2f" % ((v - t0) / 1e3) if v else "      -"
print("cta sm  slots 0..9, 14")
for i, r in enumerate(rows):
    if i < 12 or i % 37 == 0 or i >= len(rows) - 4:
        print("%4d %3d" % (i, r[15]), *[rel(r[q]) for q in range(10)], rel(r[14]))
import statistics
for k, nm in [(q, "slot%d" % q) for q in range(10)] + [(14, "end")]:
    v = [(r[k] - t0) / 1e3 for r in rows if r[k]] or [0]
    print("%-10s min %.2f median %.2f max %.2f" % (nm, min(v), statistics.median(v), max(v)))
