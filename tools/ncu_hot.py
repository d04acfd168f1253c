#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep (source page, SASS view): the instructions with the most warp-stall samples,
each with its dominant stall reasons and the instructions around it.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top=25] [context=2]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; ctxn = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    c_src, c_samp, c_exec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    total = sum(int(r[c_samp] or 0) for r in body)
    print("total samples %d, instructions %d" % (total, len(body)))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][c_samp] or 0))[:top]
    for i in order:
        r = body[i]; s = int(r[c_samp] or 0)
        st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:3]
        print("---- #%d  %5.1f%%  samples %d  executed %s  | %s" % (i, 100.0 * s / max(total, 1), s, r[c_exec], ", ".join("%s %d" % (n, v) for v, n in st if v)))
        for j in range(max(0, i - ctxn), min(len(body), i + ctxn + 1)):
            print("   %s %5d  %s" % (">>" if j == i else "  ", int(body[j][c_samp] or 0), body[j][c_src].strip()[:110]))

if __name__ == "__main__":
    main()
