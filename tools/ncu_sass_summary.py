"""Summarise `ncu --page source --csv --print-source sass` output: executed-instruction
histogram by opcode and the stall-reason totals.  Usage: ncu_sass_summary.py file.csv [top]"""
import csv
import sys
from collections import Counter


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    col = {h: i for i, h in enumerate(hdr)}
    ops, stalls = Counter(), Counter()
    total = samples = 0
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[col["Instructions Executed"]])
        except ValueError:
            continue
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[op.split(".")[0]] += n
        total += n
        samples += int(r[col["# Samples"]] or 0)
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stalls[h] += int(r[col[h]] or 0)
    print("warp instructions executed:", total, " samples:", samples)
    for op, n in ops.most_common(top):
        print("  %-12s %12d  %5.1f%%" % (op, n, 100.0 * n / total))
    print("stall samples:")
    ssum = sum(stalls.values())
    for s, n in stalls.most_common(10):
        print("  %-26s %8d  %5.1f%%" % (s, n, 100.0 * n / max(ssum, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
