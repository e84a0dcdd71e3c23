"""Per-kernel SASS evidence: counts of the instruction families that prove which hardware path a kernel takes
(tcgen05 MMA = UTC*MMA, TMEM loads / stores = LDTM / STTM, TMA bulk copies = UBLKCP / UTMA*, L2 prefetch = UBLKPF,
packed fp32 = FFMA2 / FMUL2 / FADD2, MUFU, LDGSTS) from `cuobjdump -sass liblgk.so`.
    python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "legged_games_gym_b200", "liblgk.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
FAMS = OrderedDict([("UTC*MMA (tcgen05.mma)", r"^UTC\w*MMA"), ("UTCBAR (tcgen05.commit)", r"^UTCBAR"), ("LDTM (tcgen05.ld)", r"^LDTM"),
                    ("STTM (tcgen05.st)", r"^STTM"), ("UBLKCP (cp.async.bulk)", r"^UBLKCP"), ("UBLKPF (bulk L2 prefetch)", r"^UBLKPF"),
                    ("UTMA* (tensor-map TMA)", r"^UTMA"), ("SYNCS (mbarrier)", r"^SYNCS"), ("LDGSTS (cp.async)", r"^LDGSTS"),
                    ("FFMA2/FMUL2/FADD2 (packed fp32)", r"^(FFMA2|FMUL2|FADD2)"), ("FFMA", r"^FFMA$"), ("MUFU", r"^MUFU"),
                    ("IMAD*", r"^IMAD"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
                    ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("ATOM/RED", r"^(ATOM|RED)")])
kern, counts, total, arch = None, {}, Counter(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        cur_arch = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        arch[kern] = cur_arch
        for fam, pat in FAMS.items():
            if re.match(pat, op):
                counts[kern][fam] += 1
print(f"# SASS instruction-family counts per kernel of {os.path.basename(lib)} (cuobjdump -sass; static counts)")
print("# arch of every cubin:", sorted(set(arch.values())))
for k in sorted(counts, key=lambda k: -total[k]):
    dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    fams = ", ".join(f"{f.split(' ')[0]} {n}" for f, n in counts[k].items() if n)
    print(f"{dem[:110]}\n    {total[k]} instructions [{arch.get(k)}]: {fams}")
