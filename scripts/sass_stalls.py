"""Stall-reason breakdown + hottest instructions from an ncu source-page CSV (first kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for r in rows[2:]:
    if r[0] in ('Kernel Name', 'Address'):
        break
    data.append(r)
top = sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]
for r in top:
    st = {c: int(r[hdr.index(c)]) for c in cols if int(r[hdr.index(c)]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[isamp], r[ie], r[ia].strip()[:70], st)
tot = {c: 0 for c in cols}
for r in data:
    for c in cols:
        tot[c] += int(r[hdr.index(c)])
s = sum(tot.values())
print({k: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
