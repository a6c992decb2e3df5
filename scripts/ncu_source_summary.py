#!/usr/bin/env python
"""Source-level summary of an .ncu-rep captured with --set full --import-source on: executed
instructions and stall samples per source file, stall-reason mix, opcode mix and the divergence
picture (share of executed instructions by number of active lanes, and the low-occupancy regions).

    python scripts/ncu_source_summary.py gpurun_out/prof.ncu-rep > profiles/x_source.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    # ---- per source file (cuda,sass view: SASS rows grouped under their source line)
    rows = page(rep, "cuda,sass")
    cur, hdr, ia = None, None, None
    per = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            ia = r.index("Address")
            continue
        if hdr is None or len(r) <= ia or r[ia] == "":
            continue
        per[cur][0] += 1
        per[cur][1] += num(r[hdr["Instructions Executed"]])
        per[cur][2] += num(r[hdr["# Samples"]])
    tot = sum(v[1] for v in per.values()) or 1.0
    ts = sum(v[2] for v in per.values()) or 1.0
    print("== per source file: static SASS instructions, share of executed warp-instructions, share of stall samples")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:30s} {v[0]:6d} {100 * v[1] / tot:6.2f} % {100 * v[2] / ts:6.2f} %")

    # ---- SASS view
    rows = page(rep, "sass")
    print("\n== kernel:", rows[0][1] if rows and len(rows[0]) > 1 else "?")
    hdr = {h: i for i, h in enumerate(rows[1])}
    data = [r for r in rows[2:] if len(r) == len(rows[1])]
    tot = sum(num(r[hdr["Instructions Executed"]]) for r in data) or 1.0
    ts = sum(num(r[hdr["# Samples"]]) for r in data) or 1.0
    print(f"static SASS instructions {len(data)}, executed warp-instructions {tot:.4g}")
    stall = collections.Counter()
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] = sum(num(r[hdr[h]]) for r in data)
    s_all = sum(stall.values()) or 1.0
    print("\n== warp stall samples by reason")
    for k, v in stall.most_common(12):
        print(f"{k:28s} {100 * v / s_all:6.2f} %")
    ops = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[hdr["Source"]])
        o = m.group(2) if m else "?"
        base = o.split(".")[0]
        key = o if base in ("MUFU", "F2F", "LDG", "I2F", "F2I") else base
        ops[key] += num(r[hdr["Instructions Executed"]])
    print("\n== executed warp-instructions by opcode (top 24)")
    for k, v in ops.most_common(24):
        print(f"{k:24s} {100 * v / tot:6.2f} %")
    # ---- divergence
    buckets = collections.Counter()
    lost = 0.0
    for r in data:
        n, t = num(r[hdr["Instructions Executed"]]), num(r[hdr["Avg. Threads Executed"]])
        if n:
            buckets[min(int(t // 4) * 4, 28)] += n
            lost += n * (32 - t) / 32
    print("\n== executed warp-instructions by average number of active lanes")
    for k in sorted(buckets):
        print(f"{k:2d}-{k + 3 if k < 28 else 32:2d} lanes {100 * buckets[k] / tot:6.2f} %")
    print(f"lane-slots idle in executed instructions: {100 * lost / tot:.1f} %")
    regions, cur = [], None
    for i, r in enumerate(data):
        n, t = num(r[hdr["Instructions Executed"]]), num(r[hdr["Avg. Threads Executed"]])
        if n > 0 and t < 16:
            if cur is None:
                cur = [i, i, 0.0, 0.0]
            cur[1] = i
            cur[2] += n
            cur[3] += n * t
        elif cur is not None:
            regions.append(cur)
            cur = None
    regions.sort(key=lambda c: -c[2])
    print("\n== largest code regions executed with fewer than 16 lanes")
    for c in regions[:6]:
        print(f"SASS #{c[0]}-{c[1]} ({c[1] - c[0] + 1} instructions): {100 * c[2] / tot:5.2f} % of executed, "
              f"{c[3] / c[2]:.1f} lanes on average; starts with `{data[c[0]][hdr['Source']].strip()[:48]}`")


if __name__ == "__main__":
    main()
