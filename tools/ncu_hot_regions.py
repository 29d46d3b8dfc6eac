"""Group the SASS of one kernel in an `ncu --page source --csv --print-source sass` dump into
runs of instructions with (almost) equal execution counts and print the runs that matter:
the hot loops with their instruction count per iteration.  Usage: ncu_hot_regions.py file.csv
[kernel-index] [min-share-%]"""
import csv
import sys


def kernels(path):
    rows = list(csv.reader(open(path)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
            cur["rows"].append(r)
    return out


def main(path, index=0, min_share=1.0):
    k = kernels(path)[index]
    col = {h: i for i, h in enumerate(k["hdr"])}
    data = [(r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["# Samples"]] or 0))
            for r in k["rows"]]
    total = sum(d[1] for d in data)
    print(k["name"][:100])
    print(len(data), "SASS instructions,", total, "warp instructions executed")
    i = 0
    while i < len(data):
        n = data[i][1]
        j, s, smp = i, 0, 0
        while j < len(data) and abs(data[j][1] - n) <= 0.15 * max(n, 1):
            s += data[j][1]
            smp += data[j][2]
            j += 1
        if s > min_share / 100.0 * total:
            print("instr %4d-%4d  executed %9d x %3d instr  share %5.1f%%  samples %5d" % (
                i, j - 1, n, j - i, 100.0 * s / total, smp))
        i = j


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0,
         float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
