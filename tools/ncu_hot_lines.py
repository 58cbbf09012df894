"""Top stall instructions of one kernel from `ncu -i rep --page source --csv` output.
usage: python tools/ncu_hot_lines.py rep.ncu-rep kernel_substring [which=0] [n=40]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sel = [s for s in sections if rx in s["name"]][which]
hdr = sel["rows"][0]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in sel["rows"][1:] if len(r) == len(hdr)]
S = lambda r: int(r[idx["# Samples"]] or 0)
tot = sum(S(r) for r in data)
print(sel["name"][:140], "\ntotal samples", tot, "instructions", len(data))
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[idx[k]] or 0) for r in data) for k in stall_keys}
print("stall mix:", "  ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -S(r))[:n]:
    st = sorted(((int(r[idx[k]] or 0), k[6:]) for k in stall_keys), reverse=True)[:2]
    print("%4d %6d %4.1f%%  %-70s %s" % (order[id(r)], S(r), 100 * S(r) / max(tot, 1), r[idx["Source"]].strip()[:70], st))
