"""Summarise an ncu report: key raw metrics + the top stalled SASS instructions.
usage: python tools/ncu_top.py report.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct', 'sm__warps_active.avg.pct',
        'launch__registers_per_thread', 'sm__pipe_fp64_cycles_active.avg.pct', 'smsp__issue_active.avg.pct', 'sm__throughput.avg.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'launch__occupancy_limit', 'lts__t_sector_hit_rate', 'sm__inst_executed_pipe_fp64',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'smsp__average_warps_issue_stalled', 'launch__grid_size', 'launch__shared_mem_per_block']
for r in rows[2:]:
    print("KERNEL", r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want) and 'not_issued' not in h:
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot)
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in st}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:ntop]
for i in sorted(top):
    r = data[i]
    s = {h: int(r[ix[h]]) for h in st if int(r[ix[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(i, r[ix['Source']].strip()[:64].ljust(64), r[ix['# Samples']], r[ix['Instructions Executed']], s)
