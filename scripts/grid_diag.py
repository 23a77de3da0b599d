"""Why a duration of the configs[3] training grid is slow: per-problem stop reasons, iteration and evaluation counts of the
K5c adjoint solve (SlamOptOpts.diag) for the iSwap-smush template at the given durations.
usage: python scripts/grid_diag.py [t ...]   (default 0.5 0.75 1.0)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from scripts.smush_training_grid import make_basis

REASONS = {0: "none/skipped", 1: "f_stop", 2: "gtol", 3: "gtol_far", 4: "max_iter", 5: "non-finite", 6: "solved elsewhere",
           7: "no feasible descent", 8: "line search exhausted", 9: "stalled"}


def main():
    dev = engine.require_cuda()
    ts = [float(a) for a in sys.argv[1:]] or [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]
    n_t, R = 4096, 8
    np.random.seed(0)
    rng = np.random.default_rng(0)  # one stream through the durations, as bench.py's configs[3] block draws its targets
    for t in ts:
        basis, T = make_basis(t)
        P = basis.desc.n_params
        lo, hi = basis.x0_bound_arrays()
        half = n_t // 2
        own = engine.template_eval(basis.desc, torch.as_tensor(rng.uniform(lo, hi, (half, P)), device=dev)).cpu().numpy()
        V = torch.as_tensor(np.concatenate([own, bench.haar_targets(n_t - half, 7)]), device=dev)
        for diag in (2,):
            # exactly bench.py's sequence (seeded once; a 64-target warm-up call, then the timed call), with the stop reason
            # and the evaluation count of every restart packed into the iteration table
            opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=R)
            opt.approximate_targets(V[:64].cpu().numpy(), range(1, 2))
            opts = engine.opt_defaults()
            opts.diag = diag
            torch.cuda.synchronize(); t0 = time.perf_counter()
            opt.approximate_targets(V.cpu().numpy(), range(1, 2), opts=opts)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            it = opt._ws["iters"].cpu().numpy().astype(np.int64)
            loss = opt._ws["loss"].cpu().numpy()
            reason, cnt = it >> 24, it & 0xFFFFFF
            what = "evaluations" if diag == 2 else "iterations"
            print(f"t={t} T={T} P={P} diag={diag}: {dt * 1e3:.1f} ms, total evals {opt.last_stats['evals']}")
            for r in np.unique(reason):
                sel = reason == r
                c = cnt[sel]
                print(f"   reason {r} ({REASONS.get(int(r), '?')}): {sel.sum():6d} problems, {what} mean {c.mean():9.1f} "
                      f"p50 {np.percentile(c, 50):7.0f} p99 {np.percentile(c, 99):8.0f} max {c.max():8d}; "
                      f"loss median {np.median(loss[sel]):.3e}")
            top = np.argsort(cnt.ravel())[-5:]
            print("   longest:", [(int(cnt.ravel()[i]), int(reason.ravel()[i]), float(loss.ravel()[i])) for i in top], flush=True)


if __name__ == "__main__":
    main()
