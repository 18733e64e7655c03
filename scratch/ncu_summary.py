"""usage: python scratch/ncu_summary.py <report.ncu-rep> [pairs]  - key metrics + dynamic opcode mix + top stall lines"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
units_per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__waves_per_multiprocessor'] + [
    f'smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio' for x in
    ('long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'not_selected',
     'dispatch_stall', 'no_instruction', 'branch_resolving', 'sleeping', 'membar')]
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    print("==", d['Kernel Name'][:70], d['Grid Size'], d['Block Size'])
    for k in keys:
        if k in d and d[k] != '':
            print(f'  {k:88s} {d[k]:>16s} {u[k]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[h]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, data, tot = collections.Counter(), collections.Counter(), [], 0
for i, r in enumerate(rows[h + 1:]):
    try:
        n, s = int(r[ie]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    text = r[ia].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', text)
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += n
    samp[op] += s
    tot += n
    data.append((s, n, i, text))
ts = sum(d[0] for d in data)
print(f"-- {tot} warp instructions ({tot / units_per:.1f} per unit), {ts} stall samples")
print("   " + ", ".join(f"{op} {n / units_per:.1f}" for op, n in ops.most_common(28)))
for s, n, i, text in sorted(data, reverse=True)[:16]:
    print(f'   {100 * s / max(ts, 1):5.1f}%  exec {n:9d}  #{i:5d}  {text[:90]}')
