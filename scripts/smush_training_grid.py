"""BASELINE configs[3]: parallel-drive (iSWAP-smush) training grid -- the flow of scripts/local_smush_test.ipynb in the
reference, batched.  Template: 1Q layer, ConversionGainSmushGate(0, 0, pi/2, 0, gx[T], gy[T], t), 1Q layer (k = 1), drive
amplitudes bounded to +-2 pi; for every duration t of the grid (T = t / 0.25 time slices) `--targets` targets are trained with
`--restarts` restarts each through TemplateOptimizer.approximate_targets (host buffers in and out).  Half of the targets are
instances of the template itself at t (reachable), half are Haar random (mostly not reachable at k = 1).
Solver: K5c with the analytic adjoint through the slice exponentials (default) or central differences (--fd)."""
import argparse, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from slam_decomposition_b200 import engine
from slam_decomposition_b200.basisv2 import CircuitTemplateV2
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import ConversionGainSmushGate


def make_basis(t, duration_1q=0.25, bound=2 * np.pi):
    T = round(t / duration_1q)

    def pp2(*vargs):
        return ConversionGainSmushGate(0, 0, np.pi / 2, 0, vargs[:T], vargs[T:], t_el=t)

    basis = CircuitTemplateV2(n_qubits=2, base_gates=[pp2], edge_params=[[(0, 1)]], vz_only=False, param_vec_expand=[0, T, T])
    basis.build(1)
    basis.spanning_range = range(1, 2)
    for el in basis.circuit.parameters:
        if "Q" in str(el):
            basis.add_bound(str(el), bound, -bound)
    return basis, T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=4096)
    ap.add_argument("--restarts", type=int, default=8)
    ap.add_argument("--fd", action="store_true")
    ap.add_argument("--history", type=int, default=0, help="L-BFGS pairs kept (0 = automatic)")
    ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    grid = (0.25, 0.5, 0.75, 1.0, 1.25, 1.5)
    if args.build_only:
        for t in grid:
            b, T = make_basis(t)
            print(t, T, b.desc.n_params, [p.name for p in b.circuit.parameters][-2 * T:])
        return
    dev = engine.require_cuda()
    rng = np.random.default_rng(0)
    np.random.seed(0)
    total = 0.0
    for t in grid:
        basis, T = make_basis(t)
        P = basis.desc.n_params
        lo, hi = basis.x0_bound_arrays()
        half = args.targets // 2
        X = rng.uniform(lo, hi, (half, P))
        own = engine.template_eval(basis.desc, torch.as_tensor(X, device=dev)).cpu().numpy()
        V = np.concatenate([own, bench.haar_targets(args.targets - half, 7)])
        opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=args.restarts)
        opt.smush_adjoint = not args.fd
        if args.history:
            opt.tune = {"history": args.history}
        opt.approximate_targets(V[:64], range(1, 2))  # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = opt.approximate_targets(V, range(1, 2))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        total += dt
        ok = out["loss"] <= 1e-9
        print(f"t={t:4.2f} T={T} P={P:2d}: {args.targets} targets x {args.restarts} restarts in {dt * 1e3:8.1f} ms "
              f"({opt.last_stats['evals'] / dt / 1e6:6.1f} M evals/s); solved <= 1e-9: template instances {ok[:half].mean():.3f}, "
              f"Haar {ok[half:].mean():.3f}; max |amplitude| {np.abs(out['Xk'][:, -2 * T:]).max():.3f}", flush=True)
    print(f"grid total {total:.2f} s ({'central differences' if args.fd else 'adjoint'})")


if __name__ == "__main__":
    main()
