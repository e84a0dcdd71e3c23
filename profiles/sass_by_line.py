"""Attribute the per-SASS-instruction metrics of an `ncu --page source --csv` export to CUDA source lines.
    python profiles/sass_by_line.py <ncu_source.csv> <nvdisasm -g -c output> [top]
The nvdisasm listing carries `//## File "...", line N` markers; instruction order is the same in both listings."""
import csv
import re
import sys

src_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
section = sys.argv[4] if len(sys.argv) > 4 else None      # substring of the mangled kernel name (templates)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
kernel = rows[0][1].split("(")[0].split("::")[-1]
lines, cur, active = [], ("?", 0), False
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        active = (section or kernel) in l
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
print(f"kernel {kernel}: {len(data)} instructions in the ncu export, {len(lines)} in the disassembly")
n = min(len(data), len(lines))
agg = {}
for r, ln in zip(data[:n], lines[:n]):
    a = agg.setdefault(ln, [0, 0, 0, {}])
    a[0] += int(r[ci["# Samples"]])
    a[1] += int(r[ci["Instructions Executed"]])
    a[2] += 1
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h and int(r[ci[h]]) > 0:
            a[3][h[6:]] = a[3].get(h[6:], 0) + int(r[ci[h]])
tot_s = sum(a[0] for a in agg.values())
tot_i = sum(a[1] for a in agg.values())
print(f"samples {tot_s}, warp instructions executed {tot_i}")
src_cache = {}
def text(f, n):
    import glob
    if f not in src_cache:
        c = glob.glob("/root/repo/**/" + f, recursive=True)
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    t = src_cache[f]
    return t[n - 1].strip()[:80] if 0 < n <= len(t) else ""
for title, key in (("by stall samples", 0), ("by warp instructions executed", 1)):
    print("---", title)
    for ln, a in sorted(agg.items(), key=lambda x: -x[1][key])[:top]:
        st = dict(sorted(a[3].items(), key=lambda x: -x[1])[:3])
        print(f"{ln[0]}:{ln[1]:4d} samples {a[0]:5d} ({100 * a[0] / max(tot_s, 1):4.1f}%) inst {a[1]:8d} ({100 * a[1] / max(tot_i, 1):4.1f}%) static {a[2]:4d} {st}  | {text(*ln)}")
