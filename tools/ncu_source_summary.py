#!/usr/bin/env python
"""Summarises the source page of an ncu report (ncu -i X.ncu-rep --page source --csv > src.csv): stall reasons over all
samples, samples and executed instructions by opcode, and the most-sampled instructions with their top stall reasons.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv && python tools/ncu_source_summary.py /tmp/src.csv
"""
import collections
import csv
import sys


def main(path, top_n=20):
    rows = list(csv.reader(open(path)))
    print(rows[0][0], rows[0][1] if len(rows[0]) > 1 else "")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    samp, ins, src = ix["# Samples"], ix["Instructions Executed"], ix["Source"]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(int(r[samp]) for r in data)
    tot_i = sum(int(r[ins]) for r in data)
    print(f"SASS instructions {len(data)} ({len(data) * 16 / 1024:.0f} KB), executed warp instructions {tot_i}, stall samples {tot_s}")
    agg = collections.Counter()
    for r in data:
        for h in stall:
            agg[h] += int(r[ix[h]] or 0)
    print("stall reasons, % of samples: " + ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}" for k, v in agg.most_common(12)))
    by_s, by_i = collections.Counter(), collections.Counter()
    for r in data:
        tok = r[src].strip().split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        by_s[op] += int(r[samp])
        by_i[op] += int(r[ins])
    print("opcode      samples%  executed%")
    for op, c in by_s.most_common(16):
        print(f"{op:10s} {100 * c / tot_s:8.2f} {100 * by_i[op] / tot_i:9.2f}")
    print("most-sampled instructions: samples, executed, SASS, top stall reasons")
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][samp]))[:top_n]:
        r = data[i]
        st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stall), key=lambda kv: -kv[1])[:3]
        print(f"{int(r[samp]):6d} {int(r[ins]):9d}  {r[src].strip()[:58]:58s} {dict((k, v) for k, v in st if v)}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)
