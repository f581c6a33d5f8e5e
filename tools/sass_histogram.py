#!/usr/bin/env python
"""Opcode histogram of the SASS in the shipped library (cuobjdump -sass), per kernel family: evidence of what the compiled
code is made of (bulk-async copies UBLKCP / UBLKRED, mbarrier SYNCS, native shared-memory integer atomics ATOMS.ADD, ...).

    python tools/sass_histogram.py [dualip_b200/_lib/libdualip_b200.so] > profiles/r2/sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dualip_b200/_lib/libdualip_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
fam = None
hist = collections.defaultdict(collections.Counter)
size = collections.Counter()
arch = set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        for key in ("matching_slab_kernel", "matching_long_cta_kernel", "matching_long_kernel", "agd_step_peer_kernel", "agd_step_kernel",
                    "epilogue_kernel", "lp_", "fair_calc_kernel", "project_block_kernel", "fill_slabs", "cta_row_bound"):
            if key in name:
                fam = key
                break
        else:
            fam = "other (plan builders, operators, cub)"
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fam:
        op = m.group(1)
        hist[fam][op] += 1
        size[fam] += 1
print(f"{lib}: arch {sorted(arch)}")
interesting = ("UBLKCP", "UBLKRED", "SYNCS", "ATOMS", "ATOMG", "RED", "REDUX", "SHFL", "LDG", "LDS", "STS", "STG", "F2I", "F2F", "DADD", "DFMA", "MUFU", "FMNMX", "BAR")
for f in sorted(hist, key=lambda k: -size[k]):
    h = hist[f]
    print(f"\n== {f}: {size[f]} SASS instructions ({size[f] * 16 / 1024:.0f} KB)")
    groups = collections.Counter()
    for op, c in h.items():
        groups[op.split(".")[0]] += c
    print("  top opcodes: " + ", ".join(f"{op} {c}" for op, c in groups.most_common(14)))
    marks = []
    for key in interesting:
        sel = {op: c for op, c in h.items() if op.startswith(key)}
        if sel:
            top = sorted(sel.items(), key=lambda kv: -kv[1])[:4]
            marks.append(f"{key}: " + ", ".join(f"{op} {c}" for op, c in top))
    print("  " + "\n  ".join(marks))
