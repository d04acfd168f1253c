#!/usr/bin/env python
"""Warp-stall samples per CUDA source line of one kernel from an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top=30] [file-substring]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30; want = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fname = ""; h = None; agg = {}
    for r in csv.reader(io.StringIO(raw)):
        if not r: continue
        if r[0] == "File Name": fname = r[1]; continue
        if r[0] == "Line No": h = r; cs = h.index("# Samples"); stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]; continue
        if h is None or len(r) != len(h) or not r[0]: continue          # SASS rows have an empty line number
        try: s = int(r[cs])
        except ValueError: continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, r[1], [0] * len(stall)])
        a[0] += s
        for j, c in enumerate(stall):
            try: a[2][j] += int(r[c])
            except ValueError: pass
    total = sum(a[0] for a in agg.values())
    print("total samples", total)
    names = [h[c] for c in stall]
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if want and want not in f: continue
        st = sorted(zip(a[2], names), reverse=True)[:2]
        print("%5.1f%% %7d  %s:%d  %-90s | %s" % (100.0 * a[0] / max(total, 1), a[0], f.split("/")[-1], ln, a[1].strip()[:90], ", ".join("%s %d" % (n, v) for v, n in st if v)))

if __name__ == "__main__":
    main()
