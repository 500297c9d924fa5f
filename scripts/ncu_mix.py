"""Dynamic SASS opcode mix of one kernel from an ncu report's source page (development aid).
usage: ncu -i rep --page source --csv --kernel-name regex:... --launch-skip N --launch-count 1 > src.csv; python scripts/ncu_mix.py src.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
data = rows[rows.index(hdr) + 1:]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, smp, n = collections.Counter(), collections.Counter(), 0
first = None
for r in data:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    op = r[isrc].strip().split()
    if not op:
        continue
    o = op[0] if not op[0].startswith('@') else op[1]
    o = o.split('.')[0]
    c = int(r[iex])
    first = first or c
    tot[o] += c; smp[o] += int(r[ismp]); n += c
print(f"warps {first}; total warp-insts {n}; per warp {n/first:.1f}")
for o, c in tot.most_common(24):
    print(f"{o:10s} {c/first:8.1f} per warp   stall samples {smp[o]}")
