"""Per-kernel stall summary of an ncu report's source page (several kernels per report).
usage: python tools/ncu_src.py report.ncu-rep [ntop] [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30; want = sys.argv[3] if len(sys.argv) > 3 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, data=[]); blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks:
    if want not in b["name"]:
        continue
    hdr, data = b["hdr"], b["data"]; ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print("KERNEL", b["name"][:90], "samples", tot, "instructions", len(data))
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in st}
    print(" stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:ntop]
    for i in sorted(top):
        r = data[i]; s = {h: int(r[ix[h]]) for h in st if int(r[ix[h]]) > 0}; s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
        print(" ", i, r[ix['Source']].strip()[:70].ljust(70), r[ix['# Samples']], r[ix['Instructions Executed']], s)
