#!/usr/bin/env python
"""Summarise an .ncu-rep: per launch key metrics + (optional) per-opcode executed counts for one launch.
usage: ncu_summary.py report.ncu-rep [launch_index_for_opcode_breakdown] [voxels]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
short = ['ms', 'rdMB', 'wrMB', 'inst', 'issue%', 'alu%', 'fma%', 'warps%', 'regs', 'dram%', 'barr', 'math', 'wait', 'ssb', 'lsb',
         'disp', 'nsel', 'brch', 'noin', 'mio', 'lg']
print(f"{'kernel':34s} " + " ".join(f"{s:>8s}" for s in short))
ki = H.index('Kernel Name')
for r in data:
    name = r[ki].replace('void vpb::<unnamed>::', '').replace('(vpb::<unnamed>::PassArgs)', '').replace('(int)', '').replace('(bool)', '')[:34]
    vals = []
    for w in want:
        try:
            v = float(r[H.index(w)].replace(',', ''))
            u = units[H.index(w)]
            if w.startswith('gpu__time'):
                v = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else v)
            if 'bytes' in w:
                v = v / 1e6 if u == 'byte' else (v / 1e3 if u == 'Kbyte' else (v * 1e3 if u == 'Gbyte' else v))
            vals.append(f"{v:8.3g}" if v < 1e5 else f"{v:8.2e}")
        except Exception:
            vals.append(f"{'-':>8s}")
    print(f"{name:34s} " + " ".join(vals))

if len(sys.argv) > 2:
    li = int(sys.argv[2])
    vox = float(sys.argv[3]) if len(sys.argv) > 3 else None
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    H = rows[hi]
    ie, isrc, ist = H.index('Instructions Executed'), H.index('Source'), H.index('Warp Stall Sampling (All Samples)')
    byop, stall, tot, seen = collections.Counter(), collections.Counter(), 0, set()
    for r in rows[hi + 1:]:
        if len(r) <= ie or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            n = int(r[ie])
        except ValueError:
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        byop[op] += n
        tot += n
        try:
            stall[op] += int(r[ist])
        except ValueError:
            pass
    print(rows[0][1][:100] if rows[0] else '')
    print('total warp instr', tot, ' thread-instr per voxel', tot * 32 / vox if vox else '')
    ts = sum(stall.values())
    for op, n in byop.most_common(28):
        print(f"{op:10s} {n:12d} {n / tot * 100:5.1f}%  per voxel {n * 32 / vox if vox else 0:6.1f}   stall {stall[op] / ts * 100:5.1f}%")
