"""Developer aid: builds a -DOCF_TIMELINE copy of the library, runs one kernel of bench.kernel_table and prints the per-CTA
globaltimer stamps (ns relative to the earliest CTA start).  usage: python tools/timeline.py corr_fwd_L2 [--nobuild]
[--env "A=1,B=2;A=0"]  (several environment settings, separated by ';', probed one after the other in this process:
the library re-reads its developer knobs on every call under OCF_KNOBS_DYNAMIC=1)"""
import ctypes, os, subprocess, sys
os.environ["OCF_KNOBS_DYNAMIC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocflow_b200 import build as B
dbg = os.path.join(ROOT, "tools", "bin", "libocflow_tl.so")
os.makedirs(os.path.dirname(dbg), exist_ok=True)
if "--nobuild" not in sys.argv:
    # corr.cu (the only instrumented source) with -DOCF_TIMELINE, linked against the product build's other objects
    B.build()
    obj = os.path.join(ROOT, "tools", "bin", "corr_tl.o")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + ["-DOCF_TIMELINE"] + [a for a in sys.argv[2:] if a.startswith("-D")] +
                          ["-c", "-o", obj, os.path.join(B.CSRC, "corr.cu")])
    others = [os.path.join(B.OBJ, src[:-3] + ".o") for src in B.sources() if src != "corr.cu"]
    subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", dbg, obj] + others)
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
import statistics
lib.ocf_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
settings = [""]
if "--env" in sys.argv:
    settings = sys.argv[sys.argv.index("--env") + 1].split(";")
n = 1024 * 16


def capture(setting):
    kv = [p.split("=") for p in setting.split(",") if p]
    for k, v in kv:
        os.environ[k] = v
    try:
        assert lib.ocf_debug_timeline(None, 0) == 0
        for _ in range(3):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * n)()
        rc = lib.ocf_debug_timeline(buf, n)
        assert rc == 0, rc
        return buf
    finally:
        for k, _ in kv:
            os.environ.pop(k, None)


for setting in settings:
  print("==== %s  [%s]" % (name, setting or "default knobs"))
  buf = capture(setting)
  rows = [list(buf[i * 16:(i + 1) * 16]) for i in range(1024)]
  rows = [r for r in rows if r[0]]
  t0 = min(r[0] for r in rows)
  print("ctas", len(rows), "span_us %.2f" % ((max(r[14] for r in rows) - t0) / 1e3))
  def rel(v): return "%7.2f" % ((v - t0) / 1e3) if v else "      -"
  print("cta sm  slots 0..9, 14")
  for i, r in enumerate(rows):
      if i < 12 or i % 37 == 0 or i >= len(rows) - 4:
          print("%4d %3d" % (i, r[15]), *[rel(r[q]) for q in range(10)], rel(r[14]))
  for k, nm in [(q, "slot%d" % q) for q in range(10)] + [(14, "end")]:
      v = [(r[k] - t0) / 1e3 for r in rows if r[k]] or [0]
      print("%-10s min %.2f median %.2f max %.2f" % (nm, min(v), statistics.median(v), max(v)))
  if name.startswith("corr_bwd"):
      # tiled backward: slot 0 entry, 1 coefficient boxes landed, 2 lifted + masked, 3 first feature stage landed, 4 first
      # 4-channel group reduced, 5 mode + 1, 14 end.  Phase lengths (us) per wave (by entry time) and mode.
      def med(v):
          return statistics.median(v) if v else float("nan")
      starts = sorted((r[0] - t0) / 1e3 for r in rows)
      print("entry times (us): " + " ".join("%.1f" % starts[i] for i in range(0, len(starts), max(1, len(starts) // 24))))
      for mode in (1, 2):
          for lo, hi, nm in ((0, 4, "wave 1 (entry < 4 us)"), (4, 1e9, "later waves")):
              sel = [r for r in rows if r[5] == mode and lo <= (r[0] - t0) / 1e3 < hi]
              if not sel:
                  continue
              d = lambda a, b: med([(r[b] - r[a]) / 1e3 for r in sel if r[a] and r[b]])  # noqa: E731
              print("mode %d %-22s n=%3d  entry->g landed %5.2f  ->lifted %5.2f  ->stage0 landed %5.2f  ->first group %5.2f  first group->end %6.2f  total %6.2f"
                    % (mode - 1, nm, len(sel), d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 14), d(0, 14)))
              if all(r[1] == 0 for r in sel):
                  print("        (no staging on this path) entry->lifted %5.2f" % d(0, 2))
