#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (evidence that the hot path is tcgen05 / TMEM / bulk-TMA code):

    python tools/sass_opcodes.py [sinnerf_b200/libsinnerf_b200.so] > profiles/rNN_sass_opcodes.txt

Counts, per kernel of `cuobjdump -sass`, the mnemonics that matter on sm_100a: UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UBLKCP / UTMALDG (bulk copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy
mma.sync -- must be absent), FFMA, MUFU, RED/ATOM, LDG/STG, LDS/STS."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "sinnerf_b200/libsinnerf_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("UTC?MMA", r"^UTC\w*MMA"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UBLKCP", r"^UBLKCP"), ("UTMALDG", r"^UTMA"),
          ("UTCBAR", r"^UTCBAR"), ("SYNCS", r"^SYNCS"), ("HMMA", r"^[HIQ]G?MMA"), ("FFMA", r"^FFMA"), ("MUFU", r"^MUFU"),
          ("F2FP", r"^F2FP"), ("RED/ATOM", r"^(RED|ATOM)"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("SHFL", r"^SHFL")]
cur, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for name, pat in GROUPS:
            if re.match(pat, op):
                counts[cur][name] += 1
                break
names = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: SASS opcode counts per kernel (cuobjdump -sass); columns: total instructions, then the groups below")
print("# " + " ".join(n for n, _ in GROUPS))
for (mangled, c), pretty in zip(counts.items(), names):
    pretty = pretty.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("snb::", "").replace("(bool)", "")
    m = re.search(r">\(", pretty)
    short = (pretty[:m.start() + 1] if m else re.sub(r"\(.*", "", pretty)).replace("void ", "")[:64]
    print(f"{short:66s} {total[mangled]:6d}  " + " ".join(f"{n}={c[n]}" for n, _ in GROUPS if c[n]))
agg = collections.Counter()
for c in counts.values():
    agg.update(c)
print("# library totals: " + " ".join(f"{n}={agg[n]}" for n, _ in GROUPS))
