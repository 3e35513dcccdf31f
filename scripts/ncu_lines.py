#!/usr/bin/env python
"""Per CUDA source line: instructions executed and warp-stall samples of one kernel in an .ncu-rep captured with
`--import-source on` (kernels compiled with -lineinfo).  Read here, no GPU needed.

    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":     # a CUDA line row (its SASS rows follow)
            try:
                agg[(cur_file, int(r[0]), r[1].strip()[:90])] = (int(r[iI]), int(r[iS]))
            except ValueError:
                pass
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print("total instructions %d, stall samples %d" % (ti, ts))
    for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100.0 * s / ts, 100.0 * i / ti, f, ln, src))


if __name__ == "__main__":
    main()
