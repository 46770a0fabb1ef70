"""Summarise an .ncu-rep (read on the CPU box): key raw metrics, stall mix, hottest SASS lines.
Usage: python tools/ncu_summary.py report.ncu-rep [n_px] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print("==", r[idx['Kernel Name']][:100])
    for w in want:
        if w in idx:
            print(f"   {w} = {r[idx[w]]} {rows[1][idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page holds one table per kernel, separated by a "Kernel Name" line
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    print("==", lines[0][1][:100])
    h = lines[1]; ix = {k: i for i, k in enumerate(h)}
    data = [r for r in lines[2:] if len(r) == len(h)]
    tot = sum(int(r[ix['Instructions Executed']]) for r in data)
    ts = sum(int(r[ix['# Samples']]) for r in data)
    print(f"   warp-instructions {tot}" + (f" = {tot / npx:.2f} per px" if npx else "") + f", samples {ts}")
    st = {k: sum(int(r[ix[k]]) for r in data) for k in h if k.startswith('stall_') and 'Not Issued' not in k}
    print("   stalls:", ", ".join(f"{k[6:]} {v / ts:.0%}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / ts > 0.01))
    ops = collections.Counter()
    for r in data:
        t = r[ix['Source']].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[ix['Instructions Executed']])
    print("   opcodes:", ", ".join(f"{k} {v / tot:.0%}" for k, v in ops.most_common(14)))
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
        print(f"   {r[ix['# Samples']]:>5} samp {r[ix['Instructions Executed']]:>9} exec  {r[ix['Source']][:100]}")
