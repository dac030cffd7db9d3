#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + SASS opcode mix + hot regions (run where ncu is installed)."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 120
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
for r in rows[2:]:
    print('===', r[hdr.index('Kernel Name')][:70])
    for wn in want:
        if wn in hdr:
            i = hdr.index(wn)
            print(f'  {wn:82s} {r[i]:>16s} {units[i]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be concatenated; take the first block
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[start]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[start + 1:]:
    if len(r) <= iex or r[0] == 'Address' or r[0] == 'Kernel Name':
        if r and r[0] == 'Kernel Name':
            break
        continue
    try:
        data.append((r[isrc], int(r[iex]), int(r[ismp]), [int(r[i] or 0) for i in stall_cols]))
    except ValueError:
        pass
tot = sum(d[1] for d in data)
tots = max(sum(d[2] for d in data), 1)
print(f'\nSASS instrs {len(data)}; warp instr executed {tot}; samples {tots}')


def opcode(s):
    t = s.split()
    op = t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0]


mix, smp = collections.Counter(), collections.Counter()
for s, e, sm, _ in data:
    mix[opcode(s)] += e
    smp[opcode(s)] += sm
print('opcode mix: ' + ', '.join(f'{k} {v / tot * 100:.1f}%' for k, v in mix.most_common(22)))
names = [hdr[i] for i in stall_cols]
agg = [sum(d[3][j] for d in data) for j in range(len(stall_cols))]
print('stall samples: ' + ', '.join(f'{n} {a / tots * 100:.1f}%' for n, a in sorted(zip(names, agg), key=lambda t: -t[1])[:10]))
print()
for i in range(0, len(data), seg):
    chunk = data[i:i + seg]
    e = sum(c[1] for c in chunk)
    s = sum(c[2] for c in chunk)
    ops = collections.Counter()
    for c in chunk:
        ops[opcode(c[0])] += c[1]
    st = [sum(c[3][j] for c in chunk) for j in range(len(stall_cols))]
    top_st = ', '.join(f'{n[6:]} {a * 100 // max(s, 1)}%' for n, a in sorted(zip(names, st), key=lambda t: -t[1])[:3])
    top = ', '.join(f'{k}:{v * 100 // max(e, 1)}%' for k, v in ops.most_common(5))
    print(f'sass[{i:4d}:{i + seg:4d}] instr {e / tot * 100:5.1f}%  samples {s / tots * 100:5.1f}%   {top}   | {top_st}')
