#!/usr/bin/env python
"""Stall-reason totals (and per hot line) from an `ncu --page source --csv --print-source cuda,sass` export."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hdr = None; cur = None; tot = collections.Counter(); per = {}
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": curf = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None: continue
    if row[0] != "":
        try: cur = (curf, int(row[0]), row[1].strip()[:70])
        except ValueError: cur = None
        if cur:
            d = dict(zip(hdr[4:], row[4:]))
            c = collections.Counter()
            for k, v in d.items():
                if k.startswith("stall_") and "Not Issued" not in k:
                    try: c[k] += float(v)
                    except ValueError: pass
            per[cur] = c; tot.update(c)
s = sum(tot.values())
print("TOTAL", " ".join(f"{k[6:]}={v/s*100:.1f}%" for k, v in tot.most_common(10)))
for key, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    t = sum(c.values())
    print(f"{key[0]}:{key[1]} {t/s*100:.1f}% | " + " ".join(f"{k[6:]}={v/t*100:.0f}%" for k, v in c.most_common(4)) + " | " + key[2])
