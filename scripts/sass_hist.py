"""Dynamic opcode histogram of one kernel from `ncu --page source --csv --print-source sass` output."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = 0
hdr = None
tot = 0; by = collections.Counter(); samp = collections.Counter(); tots = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kern += 1
        continue
    if r and r[0] == "Address":
        hdr = r
        ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
        continue
    if kern != 1 or hdr is None or len(r) <= ie:
        continue
    src = r[ia].strip(); n = int(r[ie]); s = int(r[isamp])
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)', src)
    op = m.group(2) if m else src
    by[op] += n; samp[op] += s; tot += n; tots += s
print("total warp-inst", tot, "samples", tots)
for op, n in by.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(tots,1):5.1f}%")
