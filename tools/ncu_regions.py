"""Hot regions of a kernel from an ncu report: contiguous SASS ranges with their share of the issued warp-instructions and
the active lanes per instruction.   ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_regions.py x.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ie, it = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isamp = hdr.index("# Samples")
data = [(r[isrc].strip(), int(r[ie]), int(r[it]), int(r[isamp])) for r in rows[2:] if len(r) > it]
tot = sum(d[1] for d in data); tott = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print("total warp-inst", tot, "avg lanes", tott / tot, "samples", tots)
# regions: split where executed count changes by > 25%
regions = []; start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[i - 1][1]) > 0.25 * max(data[i][1], data[i - 1][1], 1) or data[i - 1][0].split()[0] in ("BRA", "EXIT", "BSYNC", "@P0", "@!P0") and False:
        regions.append((start, i)); start = i
out = []
for a, b in regions:
    e = sum(d[1] for d in data[a:b]); t = sum(d[2] for d in data[a:b]); s = sum(d[3] for d in data[a:b])
    if e == 0: continue
    out.append((a, b, e, t, s))
for a, b, e, t, s in out:
    if e / tot < 0.004: continue
    ops = " ".join(d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1] for d in data[a:min(b, a + 14)])
    print(f"[{a:4d},{b:4d}) n={b-a:3d} exec/inst={e/(b-a):.3e} share={100*e/tot:5.1f}% lanes={t/e:5.1f} samp={100*s/tots:5.1f}%  {ops[:150]}")
