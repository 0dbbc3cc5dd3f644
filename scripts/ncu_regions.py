"""Bucket the SASS of a kernel in an .ncu-rep (source page) into regions of equal execution count:
instruction share and stall-sample share per region, first instruction shown, opcode mix on request."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
idx = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
s0 = idx[which]; e0 = idx[which + 1] if which + 1 < len(idx) else len(rows)
print(rows[s0][1])
hdr = rows[s0 + 1]; data = [r for r in rows[s0 + 2:e0] if len(r) == len(hdr)]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("instructions", tot, "samples", tots, "sass lines", len(data))
cur = None; start = 0; regions = []
for i, r in enumerate(data):
    e = int(r[iex])
    if cur is None or abs(e - cur) > 0.03 * max(cur, 1):
        if cur is not None: regions.append((start, i - 1, cur))
        cur = e; start = i
regions.append((start, len(data) - 1, cur))
for s, e, c in regions:
    n = e - s + 1
    smp = sum(int(data[k][ismp]) for k in range(s, e + 1))
    if c * n > 0.004 * tot or smp > 0.004 * tots:
        print("%5d-%5d n=%4d exec=%10d inst%%=%5.1f samples%%=%5.1f  %s" % (s, e, n, c, c * n / tot * 100, smp / tots * 100, data[s][isrc].strip()[:60]))
if len(sys.argv) > 4:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for k in range(a, b + 1): print("%6s %s" % (data[k][ismp], data[k][isrc]))
