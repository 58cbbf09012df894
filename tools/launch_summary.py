"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = collections.OrderedDict()
total = 0.0
n = 0
for r in rows[h + 1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1000 if r[iu] in ("ns", "nsecond") else (v * 1000 if r[iu] in ("ms", "msecond") else v)
    name = r[ik]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
mine = ("corr_", "warp_", "range_map", "occ_photo", "norm_", "group_sums", "smooth_", "photometric", "robust_l1", "pair_loss", "gradient_kernel", "flow_to_warp", "occ_from_range", "resize_", "pack_pairs")
print("launches %d   total kernel time %.1f us" % (n, total))
mt = sum(v[1] for k, v in agg.items() if k.startswith(mine))
mc = sum(v[0] for k, v in agg.items() if k.startswith(mine))
print("ocflow_b200 kernels: %d launches, %.1f us = %.2f%% of kernel time" % (mc, mt, 100 * mt / total))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%8.1f us %5.1f%%  x%-5d %s%s" % (t, 100 * t / total, c, "* " if k.startswith(mine) else "  ", k))
