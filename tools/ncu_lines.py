#!/usr/bin/env python
"""Aggregate an .ncu-rep's warp-stall samples by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys

def main(rep, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, agg, total = None, None, {}, 0.0
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]; continue
        if len(r) == 2: continue
        if r and r[0] == 'Line No':
            hdr = {n: i for i, n in enumerate(r)}; continue
        if hdr is None or not r: continue
        if r[0] != '':      # a CUDA source line row (aggregated over its SASS)
            try:
                s = float(r[hdr['# Samples']]); ex = float(r[hdr['Instructions Executed']])
            except ValueError:
                continue
            key = (cur_file, int(r[0]))
            a = agg.setdefault(key, [0.0, 0.0, r[1].strip()[:90]])
            a[0] += s; a[1] += ex; total += s
    print(f"== {rep}: {total:.0f} samples")
    for (f, ln), (s, ex, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*s/total:6.2f}%  x{ex:14.0f}  {f}:{ln:<5d} {src}")

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
