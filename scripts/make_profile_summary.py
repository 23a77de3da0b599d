"""Turn an .ncu-rep (ncu --set full --import-source on) into a small text summary for profiles/.
usage: make_profile_summary.py <report.ncu-rep> <out.txt> [kernel-regex]"""
import csv, io, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__cycles_elapsed.avg',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
    'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
lines = [f"# summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)"]
ki = H.index('Kernel Name')
for r in rows[2:]:
    lines.append(f"\n## launch: {r[ki]}")
    for m in METRICS:
        if m in H:
            i = H.index(m)
            lines.append(f"{m:88s} {r[i]:>18s} {rows[1][i]}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-count', '1'],
                     capture_output=True, text=True).stdout
cur = None; agg = collections.Counter(); aggx = collections.Counter(); ai = ei = None
reasons = collections.defaultdict(collections.Counter); ridx = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        ai = r.index('Warp Stall Sampling (All Samples)'); ei = r.index('Instructions Executed') if 'Instructions Executed' in r else None
        ridx = {i: c for i, c in enumerate(r) if c.startswith('stall_') and 'Not Issued' not in c}; continue
    if r[0].isdigit() and ai is not None:
        def I(v):
            try: return int(v)
            except ValueError: return 0
        k = (cur, int(r[0]), r[1].strip()[:100]); agg[k] += I(r[ai]); aggx[k] += I(r[ei]) if ei is not None else 0
        for i, c in ridx.items():
            if i < len(r): reasons[k][c[6:]] += I(r[i])
tot = sum(agg.values()) or 1; totx = sum(aggx.values()) or 1
lines.append("\n## first launch: warp-stall samples / executed warp instructions by source file")
byf = collections.Counter(); byfx = collections.Counter()
for k, v in agg.items(): byf[k[0]] += v; byfx[k[0]] += aggx[k]
for f, v in byf.most_common(): lines.append(f"{f:28s} samples {100*v/tot:5.1f}%  exec {100*byfx[f]/totx:5.1f}%")
lines.append("\n## hottest CUDA source lines (share of stall samples, share of executed instructions)")
for k, v in agg.most_common(25):
    top = ' '.join(f'{n}:{100*c/max(1,sum(reasons[k].values())):.0f}%' for n, c in reasons[k].most_common(3))
    lines.append(f"{100*v/tot:5.1f}%  {100*aggx[k]/totx:5.1f}%  {k[0]}:{k[1]}  {k[2]}   [{top}]")
open(out, 'w').write("\n".join(lines) + "\n")
print("wrote", out)
