#!/usr/bin/env python
"""Instruction / stall-sample share per source line RANGE of an `ncu --page source --csv` export.
usage: ncu_phases.py export.csv file:lo-hi[:label] ...   (lines not covered are reported as 'other')"""
import csv, sys, collections
path = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    parts = a.split(":")
    lo, hi = parts[1].split("-")
    ranges.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else a))
inst = collections.Counter(); stall = collections.Counter()
cur = None; hdr = None
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or row[0] == "": continue
    try: ln = int(row[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], row[4:]))
    def f(k):
        try: return float(d.get(k, 0))
        except ValueError: return 0.0
    key = "other:" + cur
    for fn, lo, hi, label in ranges:
        if cur == fn and lo <= ln <= hi: key = label; break
    inst[key] += f("Instructions Executed"); stall[key] += f("Warp Stall Sampling (All Samples)")
ti, ts = sum(inst.values()), sum(stall.values())
for k, v in inst.most_common():
    print(f"{k:40s} inst {v/ti*100:5.1f}%  stall {stall[k]/max(ts,1)*100:5.1f}%")
