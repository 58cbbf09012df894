"""profiles/traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of each hot-path kernel, from
an `ncu --set full` capture of `tools/run_kernel.py <names> --reps 1` (launch order == the order of <names>).
usage: python tools/make_traffic.py rep.ncu-rep name1 name2 ...   (normalize_* entries span 2 launches)"""
import csv, json, os, subprocess, sys
rep, names = sys.argv[1], sys.argv[2:]
skip = 0
if names and names[0].startswith("--skip="):   # launches that precede the named ones (bench.kernel_table runs one statistics pass per level while it builds the table)
    skip, names = int(names[0][7:]), names[1:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def val(r, key):
    return float(r[idx[key]].replace(",", "")) * scale[units[idx[key]]]
launches = [(r[idx["Kernel Name"]], val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
             32.0 * float(r[idx["lts__t_sectors.sum"]].replace(",", "")) if "lts__t_sectors.sum" in idx else 0.0) for r in rows[2:]][skip:]
out, i = {}, 0
# launches per name: normalize_bwd = sums + apply, range_map / warp_bwd carry their zero-fill (a memset node is not a kernel: 1 launch)
for n in names:
    k = 2 if n.startswith("normalize_bwd") or n.startswith("normalize_fwd") else 1
    out[n] = int(sum(b for _, b, _ in launches[i:i + k]))
    out["lts:" + n] = int(sum(l for _, _, l in launches[i:i + k]))   # lts__t_sectors x 32 B: bytes through the L2 (reads + writes; gradients that stay dirty in L2 at
    i += k                                                            # kernel end do not show up in dram__bytes_write)
assert i == len(launches), (i, len(launches))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
