#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_lines.py export.csv [top_n]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None; out = []
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or row[0] == "": continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], row[4:]))
    def f(k):
        try: return float(d.get(k, 0))
        except ValueError: return 0.0
    out.append((cur_file, ln, row[1].strip()[:90], f("Instructions Executed"), f("# Samples"), f("Warp Stall Sampling (All Samples)")))
ti = sum(o[3] for o in out); ts = sum(o[5] for o in out)
print(f"total inst {ti:.0f} samples {ts:.0f}")
for o in sorted(out, key=lambda o: -o[5])[:top]:
    print(f"{o[0]:16s}{o[1]:5d} inst {o[3]/ti*100:5.1f}% stall {o[5]/max(ts,1)*100:5.1f}%  {o[2]}")
