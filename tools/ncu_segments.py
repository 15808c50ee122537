"""Bucket the warp-stall samples of an `ncu --page source --csv` dump into code segments split at barriers."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total", tot)
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
acc = 0; start = 0; desc = []; reasons = {}
def flush(i):
    global acc, start, desc, reasons
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
    print(f"[{start:4d}-{i:4d}] {acc:7d} ({100 * acc / tot:4.1f}%)  {' '.join(desc[:7]):60s} {top}")
    acc = 0; start = i + 1; desc = []; reasons = {}
for i, r in enumerate(data):
    src = r[ix['Source']].strip()
    n = int(r[ix['# Samples']]); acc += n
    for h in st:
        v = int(r[ix[h]])
        if v: reasons[h] = reasons.get(h, 0) + v
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    if any(k in src for k in ('DMMA', 'UBLKCP', 'UTMA', 'LDGSTS', 'STS', 'DEPBAR', 'FENCE', 'LDG', 'STG', 'DADD', 'DFMA', 'LDS')) and (not desc or desc[-1] != op):
        desc.append(op)
    if 'BAR.SYNC' in src or 'EXIT' in src:
        flush(i)
flush(len(data))
