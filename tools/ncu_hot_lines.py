"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export, by warp instructions executed."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tables, cur, fname, hdr = [], None, None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr, cur = r, []
        tables.append((fname, hdr, cur))
        continue
    if cur is not None and len(r) == len(hdr):
        cur.append(r)
items, tot, tot_s = [], 0, 0
for fname, hdr, cur in tables:
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    for r in cur:
        if not r[0]:
            continue            # SASS rows
        try:
            n, s = int(r[ii]), int(r[si])
        except ValueError:
            continue
        tot += n
        tot_s += s
        items.append((n, s, fname.split('/')[-1], r[0], r[1][:100]))
print("warp instructions", tot, "samples", tot_s)
for n, s, f, l, src in sorted(items, reverse=True)[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% smp  {f}:{l}  {src}")
