"""Hottest SASS instructions (warp-stall samples) of one kernel of an ncu report:
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python profiles/hot_sass.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) < 10 or r[0] == "Address" or r[0] == "Kernel Name":
        if data: break
        continue
    data.append(r)
isamp, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
tot = sum(int(r[isamp]) for r in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {int(r[isamp]):7d} {100.0*int(r[isamp])/tot:5.1f}% exec {r[iex]:>9s}  {r[isrc].strip()[:100]}")
