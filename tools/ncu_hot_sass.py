"""Top SASS instructions by warp-stall samples in an .ncu-rep (source page): python tools/ncu_hot_sass.py file.ncu-rep [N]"""
import csv
import subprocess
import sys


def main(path, n=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = rows[hdr_i + 1:]
    si, ii, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    total = sum(int(r[si] or 0) for r in body)
    print(f"{rows[0][1][:110]}\n total samples {total}")
    order = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:n]
    for k in sorted(order):
        r = body[k]
        print(f"{k:5d} {int(r[si]):7d} {100.0 * int(r[si]) / max(1, total):5.1f}%  exec {r[ei]:>9s}  {r[ii].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
