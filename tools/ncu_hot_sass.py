"""Top stalled SASS instructions of one kernel in an .ncu-rep.  usage: ncu_hot_sass.py rep kernel_regex [n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[h + 1:]):
    if len(r) <= i_ex or not r[i_s].isdigit():
        continue
    data.append((int(r[i_s]), int(r[i_ex] or 0), r[i_src].strip(), k))
tot = sum(d[0] for d in data) or 1
print("samples", tot, "warp-inst", sum(d[1] for d in data), "sass lines", len(data))
for s, e, src, k in sorted(data, reverse=True)[:n]:
    print("%6d %5.1f%%  exec %8d  [%4d] %s" % (s, 100 * s / tot, e, k, src[:110]))
