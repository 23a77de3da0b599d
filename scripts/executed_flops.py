"""EXECUTED FP64 FLOP per unit of every kernel bench.py reports a `frac_executed` for, from the committed ncu summaries
(profiles/<tag>_*_ncu_full.txt: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum.per_cycle_elapsed x
smsp__cycles_elapsed.avg of one launch) and the number
of units that launch processed (the one-launch scripts under scripts/ fix them).  Writes profiles/<tag>_executed_flops.txt;
the resulting numbers are hard-coded in bench.py (NCU_EXEC_FLOP).
usage: python scripts/executed_flops.py <tag> <evals of the profiled K5 k=3 launch>"""
import re, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
k5_evals = float(sys.argv[2]) if len(sys.argv) > 2 else 64.97e6


def launches(path):
    out, cur = [], None
    for line in open(path):
        if line.startswith("## launch:"):
            cur = {"name": line[10:].strip()}
            out.append(cur)
        elif cur is not None:
            m = re.match(r"(\S+)\s+([\d.,]+)\s", line)
            if m:
                cur[m.group(1)] = float(m.group(2).replace(",", ""))
    return out


def flop(l):
    # the full set reports these as thread instructions per elapsed cycle, summed over the SM sub-partitions
    cyc = l.get("smsp__cycles_elapsed.avg", 0.0)
    g = lambda k: l.get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed", 0.0) * cyc
    return 2 * g("dfma") + g("dmul") + g("dadd"), g("dfma"), g("dmul"), g("dadd")


ROWS = [  # key, summary file, launch index, units of that launch, what
    ("k5_eval_k3", "lbfgs_k3", 0, k5_evals, "lbfgs_kernel, k = 3 launch of the sweep: per loss+grad evaluation incl. the L-BFGS bookkeeping"),
    ("k2_lossgrad_k3", "loss_grad", 0, 2 ** 22, "loss_grad_kernel<2 lanes>, sqCNOT k = 3, per row (scripts/k2_one.py)"),
    ("k2_lossgrad_k6", "loss_grad", 1, 2 ** 22, "loss_grad_kernel<4 lanes>, sqCNOT k = 6, per row"),
    ("k3_weyl", "weyl_traj", 0, 2 ** 21, "weyl_kernel, per Haar matrix (scripts/weyl_traj_one.py)"),
    ("k4b_traj_point", "weyl_traj", 1, 2 ** 18 * 10 * 5, "trajectory_kernel, per trajectory point (N = 10 slices x R = 5 sub-times)"),
    ("k6_smush_sqcnot_k3", "coverage", 0, 4e6, "coverage_kernel, parallel-drive template (6 slice exponentials), per sample (scripts/cov_one.py)"),
    ("k6_plain_sqcnot_k3", "coverage", 1, 4e6, "coverage_kernel, plain template, per sample"),
    ("k2_smush_lossgrad", "smush_adj", 0, 2 ** 20, "smush_loss_grad_kernel<grad>, sqrt(iSWAP) k = 3 T = 2 (P = 30), per row (scripts/smush_one.py)"),
    ("k2_smush_loss", "smush_adj", 1, 2 ** 20, "smush_loss_grad_kernel<loss only>, per row"),
]
lines = [f"# executed FP64 FLOP per unit = (2 dfma + dmul + dadd thread instructions, predicated on) of one launch / units of that launch",
         f"# source: profiles/{tag}_*_ncu_full.txt (ncu --set full --clock-control none); K5 k = 3 launch: {k5_evals:.4g} evaluations", ""]
table = {}
for key, f, idx, units, what in ROWS:
    path = f"profiles/{tag}_{f}_ncu_full.txt"
    try:
        l = launches(path)[idx]
    except (OSError, IndexError):
        lines.append(f"{key}: {path} launch {idx} missing")
        continue
    fl, a, b, c = flop(l)
    table[key] = fl / units
    lines.append(f"{key:22s} {fl / units:10.1f} FLOP/unit   = (2 x {a:.4g} + {b:.4g} + {c:.4g}) / {units:.6g}   [{l['name'][:60]}; {what}]")
lines += ["", "NCU_EXEC_FLOP = {"] + [f'    "{k}": {v:.1f},' for k, v in table.items()] + ["}"]
open(f"profiles/{tag}_executed_flops.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
