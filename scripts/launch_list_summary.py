"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> profiles/<tag>_bench_launch_list.txt: launches / time / share per
kernel, the share of lbfgs_kernel in the sweep steps, and every launch in order.
usage: python scripts/launch_list_summary.py gpurun_out/<tag>_launches.csv profiles/<tag>_bench_launch_list.txt"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hi]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg, lines = collections.OrderedDict(), []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name, v, u = r[ki].split('(')[0], float(r[vi].replace(',', '')), r[ui]
    v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
    lines.append((name, v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out = ["# launch list of `python bench.py --steps 1 --warmup 1 --targets 100000 --cpu-seconds 0 --no-micro` under",
       "# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are serialised and cold-cache: shares, not absolutes)",
       f"# {len(lines)} launches, {tot:.2f} ms in kernels", "", "## by kernel (launches, total ms, share)"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{c:5d} {t:10.3f} ms {100 * t / tot:6.2f} %  {n}")
sweep = [(n, t) for n, t in lines if 'coverage_kernel' not in n and 'dfma' not in n]
st, lb = sum(t for _, t in sweep), sum(t for n, t in sweep if 'lbfgs_kernel' in n)
out += ["", "## the sweep steps only (coverage_1e9 block and the DFMA peak probe excluded; 4 sweeps: warm-up + timed, resident and e2e)",
        f"lbfgs_kernel launches {lb:.3f} ms of {st:.3f} ms in kernels = {100 * lb / st:.2f} % (bench.py kernel_share_of_step is the live, warm-cache figure)",
        "", "## launches in order (ms)"] + [f"{v:10.4f}  {n}" for n, v in lines]
open(sys.argv[2], 'w').write("\n".join(out) + "\n")
print("\n".join(out[:16]))
