"""``TemplateOptimizer`` (reference: src/slam/optimizer.py:24-313) over the batched device L-BFGS.

The reference runs, per target: for k in spanning range -> for restart in range(R) ->
``scipy.optimize.minimize(BFGS, finite-difference gradient)`` with early exits.  Here all targets x
restarts of one template size k are solved concurrently by ``slam_lbfgs_solve`` (analytic gradients,
state in shared memory); the k-loop stays on the host and carries a per-target ``active`` mask, which
reproduces "smallest k that reaches the success threshold" (optimizer.py:233, 297-303).

Return contract kept (optimizer.py:186): ``(training_loss, coordinate_list, [DataDictEntry])``.
"""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, engine
from .basis import CircuitTemplate, HamiltonianTemplate, _CircuitTemplateBase
from .basis_abc import DataDictEntry, VariationalTemplate
from .basisv2 import CircuitTemplateV2
from .cost_function import UnitaryCostFunction
from .sampler import SampleFunction
from .weyl import c1c2c3, c1c2c3_batch

SUCCESS_THRESHOLD = 1e-10
TRAINING_RESTARTS = 5


class TemplateOptimizer:
    def __init__(self, basis: VariationalTemplate, objective: UnitaryCostFunction, use_callback=False,
                 override_fail=False, success_threshold=None, training_restarts=None, override_method=None):
        self.basis = basis
        self.objective = objective
        self.preseeding = self.basis.preseeded
        self.use_callback = use_callback
        self.training_loss = []
        self.coordinate_list = []
        self.best_cycle_list = []
        self.override_fail = override_fail
        self.override_method = override_method
        self.success_threshold = SUCCESS_THRESHOLD if success_threshold is None else success_threshold
        self.training_restarts = TRAINING_RESTARTS if training_restarts is None else training_restarts
        assert not (self.preseeding and self.override_fail)
        assert not (self.preseeding and self.basis.n_qubits != 2)
        self.last_stats = {}
        self._ws = None
        self.fd_central = True  # K5c without an adjoint: central differences (False = scipy's forward differences, step 1.49e-8)
        self.smush_adjoint = True  # K5c on parameter-bound smush templates: analytic adjoint gradient (False = differences)
        # K5 launches of consecutive template sizes are chained on two streams (SlamOptOpts.solved_in / solved_out) so the
        # next size fills the SMs the tail of the current one leaves idle and no host round trip separates them
        self.pipeline = True  # set False to run one launch per size with a host round trip in between (A/B)
        self._pipe = None  # streams + per-size workspaces of the chained sweep
        self._host_x = None  # pinned staging buffer of approximate_targets()
        self._desc_cache = {}
        self.tune = {}  # e.g. {"tune_hist_min": 4}: SlamOptOpts.tune_* fields applied to the default options
        self.launch_evals = []  # (k, loss+grad evaluations) per slam_lbfgs_solve launch while engine.LBFGS_EVENTS is on

    # ------------------------------------------------------------------------------------------
    # reference entry points
    # ------------------------------------------------------------------------------------------
    def approximate_target_U(self, target_U):
        """Atomic training function for one target (optimizer.py:65-119)."""
        V = torch.as_tensor(np.asarray(target_U, dtype=np.complex128)[None], device=engine.require_cuda())
        return self._approximate_batch(V)[0]

    def approximate_from_distribution(self, sampler: SampleFunction):
        """All targets of the sampler at once (optimizer.py:180-186 runs them one by one)."""
        if hasattr(sampler, "batch"):
            V = sampler.batch(engine.require_cuda())
        else:
            V = torch.as_tensor(np.stack([np.asarray(u, dtype=np.complex128) for u in sampler]),
                                device=engine.require_cuda())
        target_data = self._approximate_batch(V)
        return self.training_loss, self.coordinate_list, target_data

    def _run(self, target_u, target_spanning_range):
        """(best_result, best_Xk, best_cycles) for one target (optimizer.py:188-313)."""
        V = torch.as_tensor(np.asarray(target_u, dtype=np.complex128)[None], device=engine.require_cuda())
        res = self._run_batch(V, target_spanning_range)
        P = int(res["best_P"][0])
        best_k = int(res["best_k"][0])
        if best_k > 0 and isinstance(self.basis, _CircuitTemplateBase):
            # the reference's loop ends built at the size it breaks on (optimizer.py:297-303), so basis.eval(best_Xk) works
            # right after _run; the batched sweep ends at max(k), hence the rebuild (gate cycles restored: not a new build
            # in the reference's sequence)
            state = self.basis.cycle_state()
            self.basis.build(n_repetitions=best_k)
            self.basis.set_cycle_state(state)
        return float(res["best_loss"][0]), res["best_x"][0, :P].cpu().numpy(), best_k

    # ------------------------------------------------------------------------------------------
    # batched core
    # ------------------------------------------------------------------------------------------
    def _cost_kind(self) -> int:
        if not isinstance(self.objective, UnitaryCostFunction):
            raise ValueError("Unrecognized Cost Function")  # optimizer.py:211
        ck = self.objective.cost_kind
        if ck is None:
            raise NotImplementedError(
                f"{type(self.objective).__name__} has no device optimiser yet (unitary_fidelity() is available)")
        return ck

    def _solver(self, desc, ck: int) -> str:
        """Solver choice.  The reference uses scipy BFGS with finite-difference gradients unless told otherwise
        (optimizer.py:255-268).  Here:
          "lbfgs"  analytic-gradient L-BFGS (K5) whenever the functional is BasicCost / SquareCost and every gate has a
                   closed-form derivative;
          "fd"     thread-per-problem L-BFGS over the generic forward evaluator (K5c) for parameter-bound smush gates and
                   BasicCostInverse: analytic adjoint gradients through the slice exponentials for smush templates
                   (`smush_adjoint`), finite differences -- the reference's own algorithm class -- otherwise;
          "nm"     the derivative-free Nelder-Mead kernel (K5b) for the coordinate-based functionals (Makhlin / Weyl /
                   reduced: piecewise constant after the 8-dp rounding), and when override_method asks for it."""
        if getattr(self.basis, "using_constraints", False):
            # the reference switches scipy to SLSQP (optimizer.py:259-264); here: augmented Lagrangian around K5c
            if ck not in (_lib.COST_BASIC, _lib.COST_SQUARE) or self.override_method == "Nelder-Mead":
                raise NotImplementedError("cost-constrained templates run with BasicCost / SquareCost and the gradient solver")
            return "con"
        if self.override_method == "Nelder-Mead":
            return "nm"
        if ck not in (_lib.COST_BASIC, _lib.COST_SQUARE, _lib.COST_BASIC_INVERSE):
            return "nm"
        if ck == _lib.COST_BASIC_INVERSE:
            return "fd"
        if desc.gate_kind in (_lib.GATE_SMUSH, _lib.GATE_SMUSH_1QPHASE):
            if any(desc.slot_param[g][s] >= 0 for g in range(desc.k) for s in range(desc.n_slots)):
                return "fd"
        return "lbfgs"

    def _x0(self, Nt: int, device) -> tuple:
        """(x0 tensor | None, lo, hi): initial points.  Uniform-box templates use the kernel's Philox stream."""
        b = self.basis
        if isinstance(b, CircuitTemplateV2):
            lo, hi = b.x0_bound_arrays()
            if np.all(lo == lo[0]) and np.all(hi == hi[0]):
                return None, float(lo[0]), float(hi[0])
            gen = torch.Generator(device=device).manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
            u = torch.rand((Nt, self.training_restarts, lo.size), dtype=torch.float64, device=device, generator=gen)
            lo_t, hi_t = torch.as_tensor(lo, device=device), torch.as_tensor(hi, device=device)
            return lo_t + (hi_t - lo_t) * u, 0.0, 1.0
        if isinstance(b, HamiltonianTemplate):
            return None, 0.0, 1.0  # np.random.random(p_len) (basis.py:48)
        return None, 0.0, 2 * np.pi  # CircuitTemplate: np.random.random(P) * 2 pi (basis.py:111)

    def _run_batch(self, V: torch.Tensor, k_range: Sequence[int], opts: Optional[_lib.SlamOptOpts] = None,
                   keep_history: bool = False) -> dict:
        b = self.basis
        if not isinstance(b, (_CircuitTemplateBase, HamiltonianTemplate)):
            raise NotImplementedError("the device optimizer runs CircuitTemplate / CircuitTemplateV2 / HamiltonianTemplate")
        if self.override_method not in (None, "BFGS", "L-BFGS-B", "Nelder-Mead"):
            raise NotImplementedError(f"override_method={self.override_method}")
        ck = self._cost_kind()
        device = V.device
        Nt = V.shape[0]
        R = int(self.training_restarts)
        if opts is None:
            opts = engine.opt_defaults()
            for name, val in self.tune.items():  # explicit launch tuning of K5 (SlamOptOpts.tune_*), for A/B measurements
                setattr(opts, name, int(val))
        opts.cost_kind = ck if ck in (_lib.COST_BASIC, _lib.COST_SQUARE) else _lib.COST_BASIC
        opts.success_threshold = float(self.success_threshold)
        opts.f_stop = min(opts.f_stop, 1e-3 * float(self.success_threshold))
        k_list = list(k_range)
        if not k_list:
            raise ValueError("empty spanning range")
        if (self.pipeline and not keep_history and opts.early_exit and len(k_list) > 1
                and not getattr(b, "using_bounds", False) and not getattr(b, "using_constraints", False)
                and self._all_lbfgs(k_list, ck)):
            return self._run_chained(V, k_list, opts)
        # persistent workspace: output tables are reused across k and across calls (no allocator churn per sweep)
        ws = self._ws
        if ws is None or ws["key"] != (Nt, R, str(device)):
            ws = {"key": (Nt, R, str(device)), "xcap": 0, "xbuf": None,
                  "loss": torch.empty((Nt, R), dtype=torch.float64, device=device),
                  "iters": torch.empty((Nt, R), dtype=torch.int32, device=device),
                  "ar": torch.arange(Nt, device=device)}
            self._ws = ws
        best_loss = torch.full((Nt,), float("inf"), dtype=torch.float64, device=device)
        best_k = torch.full((Nt,), -1, dtype=torch.int32, device=device)
        best_P = torch.zeros((Nt,), dtype=torch.int32, device=device)
        # the parameter table is sized for the largest template of the range up front, so that its shape does not depend
        # on where this rank's targets happen to be solved (ranks gather their tables with one fixed-shape collective)
        state = b.cycle_state() if hasattr(b, "cycle_state") else None
        b.build(n_repetitions=max(k_list))
        if state is not None:
            b.set_cycle_state(state)  # (a sizing build the reference does not do: its gate cycles must not advance)
        best_x = torch.zeros((Nt, b.desc.n_params), dtype=torch.float64, device=device)
        active = torch.ones((Nt,), dtype=torch.int32, device=device)
        evals = torch.zeros(1, dtype=torch.int64, device=device)
        per_k = []
        ar = ws["ar"]
        timing = engine.LBFGS_EVENTS is not None
        marks = []
        for k in k_list:
            logging.info(f"Starting opt on template size {k}")
            b.build(n_repetitions=k)
            desc = b.desc
            P = desc.n_params
            if P > ws["xcap"]:
                ws["xbuf"] = None
                ws["xbuf"] = torch.empty(Nt * R * P, dtype=torch.float64, device=device)
                ws["xcap"] = P
            x = ws["xbuf"][: Nt * R * P].view(Nt, R, P)
            x0, lo, hi = self._x0(Nt, device)
            opts.x0_lo, opts.x0_hi = lo, hi
            bound_t = None
            if getattr(b, "using_bounds", False):
                # box bounds in API order (basisv2.py:162-164).  Once any bound is set, the reference's parameter_guess
                # gives EVERY parameter a bound -- its own, or the default (-4 pi, 4 pi) (basisv2.py:156-166) -- and passes
                # the list to scipy's L-BFGS-B (optimizer.py:257-258), so unlisted parameters are boxed too; a None side
                # of an explicit bound is unbounded.  The device optimiser projects onto the box.
                lo_b, hi_b = [], []
                for prm in b.circuit.parameters:
                    bd = b.bounds.get(prm.name, None) or b.default_bound
                    l_, h_ = bd[0], bd[1]
                    l_ = -np.inf if l_ is None else float(l_)
                    h_ = np.inf if h_ is None else float(h_)
                    lo_b.append(min(l_, h_))
                    hi_b.append(max(l_, h_))
                bound_t = (torch.as_tensor(lo_b, dtype=torch.float64, device=device),
                           torch.as_tensor(hi_b, dtype=torch.float64, device=device))
                opts.lower, opts.upper = bound_t[0].data_ptr(), bound_t[1].data_ptr()
            else:
                opts.lower, opts.upper = None, None
            seed = int(np.random.randint(0, 2 ** 62))
            trace_loss = trace_x = None
            if keep_history:
                # per-iteration trace buffers (bounded: ~256 MiB); iterations beyond the cap are not recorded
                cap = int(max(8, min(opts.max_iter, (1 << 28) // max(1, Nt * R * (P + 1) * 8))))
                trace_loss = torch.zeros((Nt * R, cap), dtype=torch.float64, device=device)
                trace_x = torch.zeros((Nt * R, cap, P), dtype=torch.float64, device=device)
                opts.trace_cap, opts.trace_loss, opts.trace_x = cap, trace_loss.data_ptr(), trace_x.data_ptr()
            else:
                opts.trace_cap, opts.trace_loss, opts.trace_x = 0, None, None
            solver = self._solver(desc, ck)
            if solver == "nm":
                if getattr(b, "using_bounds", False):
                    raise NotImplementedError("box bounds with the Nelder-Mead kernel")
                nm = engine.nm_defaults()
                nm.cost_kind, nm.max_iter, nm.early_exit = ck, opts.max_iter, opts.early_exit
                nm.success_threshold, nm.x0_lo, nm.x0_hi = float(self.success_threshold), lo, hi
                loss, x, iters = engine.nm_solve(desc, V, R, nm, x0=x0, seed=seed, active=active, evals=evals,
                                                 out=(ws["loss"], x, ws["iters"]))
            elif solver == "con":
                loss, x, iters = self._solve_constrained(desc, V, R, opts, ck, x0, seed, active, evals,
                                                         (ws["loss"], x, ws["iters"]))
            elif solver == "fd":
                # scipy's BFGS stops at |g| < 1e-5, which on these near-singular landscapes is a loss of 1e-6 .. 1e-8
                # (the band where the reference's own runs end, scripts/cost_function_comparison.ipynb:118-121).  Here
                # that tolerance only ends restarts sitting at a clearly non-zero minimum (loss > 1e-4) or making < 3 %
                # progress per 32 steps; the others run on to opts.gtol / f_stop with central differences.
                saved = (opts.cost_kind, opts.f_far)
                opts.cost_kind, opts.f_far = ck, max(opts.f_far, 1e-4)
                is_smush = (desc.gate_kind in (_lib.GATE_SMUSH, _lib.GATE_SMUSH_1QPHASE)
                            and any(desc.slot_param[g][s] >= 0 for g in range(desc.k) for s in range(desc.n_slots)))
                mode = 2 if (is_smush and self.smush_adjoint) else int(bool(self.fd_central))
                loss, x, iters = engine.fd_lbfgs_solve(desc, V, R, opts, x0=x0, seed=seed, active=active, evals=evals,
                                                       out=(ws["loss"], x, ws["iters"]), central=mode)
                opts.cost_kind, opts.f_far = saved
            else:
                loss, x, iters = engine.lbfgs_solve(desc, V, R, opts, x0=x0, seed=seed, active=active, evals=evals,
                                                    out=(ws["loss"], x, ws["iters"]))
            if timing:
                marks.append((k, evals.clone()))  # async snapshot; converted to per-launch deltas after the sweep
            # a restart whose objective went non-finite reports NaN (DBL_MAX = skipped): neither may win the reduction
            lmin, rmin = torch.nan_to_num(loss, nan=float("inf")).min(dim=1)
            improved = (active != 0) & (lmin < best_loss)
            xsel = x[ar, rmin]
            if best_x is None or best_x.shape[1] < P:
                nb = torch.zeros((Nt, P), dtype=torch.float64, device=device)
                if best_x is not None:
                    nb[:, : best_x.shape[1]] = best_x
                best_x = nb
            best_x[:, :P] = torch.where(improved[:, None], xsel, best_x[:, :P])
            if best_x.shape[1] > P:
                best_x[:, P:] = torch.where(improved[:, None], torch.zeros_like(best_x[:, P:]), best_x[:, P:])
            best_loss = torch.where(improved, lmin, best_loss)
            best_k = torch.where(improved, torch.full_like(best_k, k), best_k)
            best_P = torch.where(improved, torch.full_like(best_P, P), best_P)
            if keep_history:  # per-restart tables are only copied out when the caller wants histories
                per_k.append({"k": k, "loss": loss.clone(), "x": x.clone(), "iters": iters.clone(),
                              "active": active.clone(), "desc": desc, "trace_loss": trace_loss, "trace_x": trace_x})
            active = ((active != 0) & ~(best_loss < self.success_threshold)).to(torch.int32)
            n_left = int(active.sum().item())
            logging.info(f"Cycle (k ={k}), solved {Nt - n_left}/{Nt}")
            if n_left == 0:
                logging.info(f"Break on cycle {k}")
                break
        self.last_stats = {"evals": int(evals.item())}
        prev = 0
        for k_, snap in marks:
            cur_ = int(snap.item())
            self.launch_evals.append((k_, cur_ - prev))
            prev = cur_
        return {"best_loss": best_loss.cpu().numpy(), "best_k": best_k.cpu().numpy(), "best_P": best_P.cpu().numpy(),
                "best_x": best_x, "per_k": per_k, "best_loss_dev": best_loss, "best_k_dev": best_k}

    # ------------------------------------------------------------------------------------------
    # circuit-cost constraint (CircuitTemplateV2.set_constraint, basisv2.py:192-203; SLSQP in the reference)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _circuit_cost_batch(desc, x: torch.Tensor) -> torch.Tensor:
        """circuit_cost (basisv2.py:98-127) of every row of x [..., P], on the device, from the descriptor's slot tables."""
        def slot(g, s_):
            p = desc.slot_param[g][s_]
            return x[..., p] if p >= 0 else torch.full(x.shape[:-1], float(desc.slot_const[g][s_]), dtype=x.dtype,
                                                        device=x.device)
        c = torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)
        for g in range(desc.k):
            if desc.gate_kind == _lib.GATE_RISWAP:
                c = c + slot(g, 0)
            elif desc.gate_kind in (_lib.GATE_CG, _lib.GATE_SMUSH):
                c = c + (slot(g, 2).abs() + slot(g, 3).abs()) * slot(g, desc.n_slots - 1) / (np.pi / 2)
        return c

    def _solve_constrained(self, desc, V, R, opts, ck, x0, seed, active, evals, out):
        """min cost(U(x), V)  s.t.  circuit_cost(x) <= basis.constraint_max, per (target, restart): augmented Lagrangian
        with one multiplier per problem; every inner solve is one K5c launch (central differences -- the constraint term is
        differenced together with the objective, as scipy's SLSQP differences both), warm-started from the previous one.
        Returns the PURE loss of every restart (+inf where the constraint is violated by more than 1e-7)."""
        Nt = V.shape[0]
        P = desc.n_params
        cmax = float(self.basis.constraint_max)
        lam = torch.zeros(Nt * R, dtype=torch.float64, device=V.device)
        mu = 10.0
        saved = (opts.cost_kind, opts.early_exit, opts.con_mu, opts.con_max, opts.con_lambda, opts.f_far)
        opts.cost_kind, opts.early_exit, opts.f_far = ck, 0, max(opts.f_far, 1e-4)
        live = None if active is None else (active != 0)
        x_start = x0
        try:
            for outer in range(8):
                opts.con_mu, opts.con_max, opts.con_lambda = mu, cmax, lam.data_ptr()
                loss, x, iters = engine.fd_lbfgs_solve(desc, V, R, opts, x0=x_start, seed=seed, active=active, evals=evals,
                                                       out=out, central=1)
                viol = self._circuit_cost_batch(desc, x) - cmax  # [Nt, R]
                if live is not None:
                    viol = torch.where(live[:, None], viol, torch.zeros_like(viol))
                lam = torch.clamp(lam + mu * viol.reshape(-1), min=0.0)
                if float(viol.clamp(min=0.0).max().item()) <= 1e-9 and outer > 0:
                    break
                x_start = x.clone()
                mu = min(mu * 4.0, 1e7)
        finally:
            opts.cost_kind, opts.early_exit, opts.con_mu, opts.con_max, opts.con_lambda, opts.f_far = saved
        tgt = torch.arange(Nt, dtype=torch.int32, device=V.device).repeat_interleave(R)
        pure, _, _ = engine.loss_grad(desc, x.reshape(Nt * R, P), V, tgt_idx=tgt, cost_kind=ck, want_grad=False)
        pure = pure.reshape(Nt, R)
        pure = torch.where(viol <= 1e-7, pure, torch.full_like(pure, float("inf")))
        if live is not None:
            pure = torch.where(live[:, None], pure, torch.full_like(pure, float("inf")))
        loss.copy_(pure)
        return loss, x, iters

    # ------------------------------------------------------------------------------------------
    # chained sweep: one K5 launch per template size, alternating between two streams
    # ------------------------------------------------------------------------------------------
    def _all_lbfgs(self, k_list, ck) -> bool:
        """Probe every size of the range for the solver it needs.  The reference builds each size once, in order, and its
        gate cycles advance across builds (basis.py:68-72); the probing builds are undone so that the sweep's own builds see
        the sequence the reference's k-loop sees."""
        b = self.basis
        state = b.cycle_state() if hasattr(b, "cycle_state") else None
        try:
            for k in k_list:
                b.build(n_repetitions=k)
                if self._solver(b.desc, ck) != "lbfgs":
                    return False
            return True
        finally:
            if state is not None:
                b.set_cycle_state(state)

    @staticmethod
    def _cycle_phase(b):
        """Where the basis' gate / edge cycles stand, modulo their lengths (what the next builds will see)."""
        try:
            base, edges = b.gate_2q_base, b.gate_2q_edges
            return (base.pos % len(base.items), edges.pos % len(edges.items),
                    tuple(c.pos % len(c.items) for c in edges.items))
        except (AttributeError, ZeroDivisionError):
            return None

    def _run_chained(self, V: torch.Tensor, k_list, opts, _verify_descs=None) -> dict:
        """The k-loop of optimizer.py:233-303 without host round trips and without host-side merging.  Every size gets its
        own per-restart tables; the launch for size k_{i+1} reads the (live) solved flags of size k_i and skips the targets
        already below the threshold, which is the reference's early exit; launches alternate between two streams, so the CTAs
        of the next size start on the SMs the draining launch frees.  Every retired restart also feeds one packed 64-bit
        atomic minimum per target (SlamOptOpts.best_key: the smallest size that succeeded, else the lowest loss seen --
        exactly what the sequential loop keeps), and one small kernel (slam_best_gather) copies the winners out of the
        tables afterwards: no eager tensor ops between or after the launches."""
        b = self.basis
        device = V.device
        Nt = V.shape[0]
        R = int(self.training_restarts)
        main = torch.cuda.current_stream(device)
        # The six template builds cost ~1.2 ms of host time, during which nothing is queued on the GPU.  When this basis has
        # been through the same sizes from the same gate-cycle position before, the launches are queued from the cached
        # descriptors first and the builds run afterwards, overlapped with the kernels (they are still done, in the same
        # order: the reference's k-loop leaves the basis built at the last size with its gate cycles advanced).  Each fresh
        # descriptor is compared byte for byte with the cached one; on a mismatch (the basis was modified in between) the
        # cache is dropped and the sweep repeated.
        ckey = (id(b), tuple(k_list), self._cycle_phase(b))
        cached = self._desc_cache.get(ckey) if _verify_descs is None else None
        if cached is not None:
            descs = cached
        else:
            descs = []
            for k in k_list:
                b.build(n_repetitions=k)
                descs.append((k, b.desc, b.desc.n_params))
            if _verify_descs is None:
                self._desc_cache = {ckey: descs}  # (one entry: the sweep this optimizer is used for)
        Pmax = max(P for _, _, P in descs)
        key = (Nt, R, str(device), tuple((k, P) for k, _, P in descs))
        pipe = self._pipe
        if pipe is None or pipe["key"] != key:
            pipe = {"key": key, "streams": (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)),
                    "flags": torch.empty((len(descs), Nt), dtype=torch.int32, device=device),
                    "evals": torch.empty(len(descs), dtype=torch.int64, device=device),
                    "best_key": torch.empty(Nt, dtype=torch.int64, device=device),
                    "tab": [{"loss": torch.empty((Nt, R), dtype=torch.float64, device=device),
                             "iters": torch.empty((Nt, R), dtype=torch.int32, device=device),
                             "x": torch.empty((Nt, R, P), dtype=torch.float64, device=device)} for _, _, P in descs]}
            self._pipe = pipe
        flags, evals, best_key = pipe["flags"], pipe["evals"], pipe["best_key"]
        flags.zero_()
        evals.zero_()
        best_key.fill_(-1)
        timing = engine.LBFGS_EVENTS is not None
        span = None
        if engine.LBFGS_SPANS is not None:
            span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            span[0].record(main)
        ready = torch.cuda.Event()
        ready.record(main)
        for i, (k, desc, P) in enumerate(descs):
            logging.info(f"Starting opt on template size {k}")
            st = pipe["streams"][i % 2]
            st.wait_event(ready)
            tab = pipe["tab"][i]
            with torch.cuda.stream(st):
                x0, lo, hi = self._x0(Nt, device)
                o = _lib.SlamOptOpts.from_buffer_copy(opts)
                o.x0_lo, o.x0_hi = lo, hi
                o.lower, o.upper = None, None
                o.trace_cap, o.trace_loss, o.trace_x = 0, None, None
                o.solved_in = flags[i - 1].data_ptr() if i > 0 else None
                o.solved_out = flags[i].data_ptr()
                seed = int(np.random.randint(0, 2 ** 62))
                engine.lbfgs_solve(desc, V, R, o, x0=x0, seed=seed, active=None, evals=evals[i:i + 1],
                                   out=(tab["loss"], tab["x"], tab["iters"]), best_key=best_key)
        if cached is not None:  # the deferred builds (side effects on the basis) and the check of the cached descriptors
            for k, desc, _ in descs:
                b.build(n_repetitions=k)
                if bytes(b.desc) != bytes(desc):
                    self._desc_cache = {}
                    for st in pipe["streams"]:
                        main.wait_stream(st)
                    torch.cuda.current_stream(device).synchronize()
                    return self._run_chained(V, k_list, opts, _verify_descs=False)
        for st in pipe["streams"]:
            main.wait_stream(st)
        if span is not None:
            span[1].record(main)
            engine.LBFGS_SPANS.append(span)
        best_loss, best_k, best_P, best_x = engine.best_gather(
            best_key, [(k, pipe["tab"][i]["loss"], pipe["tab"][i]["x"]) for i, (k, _, _) in enumerate(descs)], R, Pmax)
        ev_host = evals.cpu().numpy()
        self.last_stats = {"evals": int(ev_host.sum())}
        if timing:
            for (k, _, _), n in zip(descs, ev_host):
                self.launch_evals.append((k, int(n)))
        return {"best_loss": best_loss.cpu().numpy(), "best_k": best_k.cpu().numpy(), "best_P": best_P.cpu().numpy(),
                "best_x": best_x, "per_k": [], "best_loss_dev": best_loss, "best_k_dev": best_k}

    def approximate_targets(self, targets, k_range: Optional[Sequence[int]] = None, opts=None,
                            reuse_host_buffers: bool = True) -> dict:
        """Host-buffer batch API: ``targets`` complex128 [Nt,4,4] (numpy, or a pinned/CPU torch tensor) ->
        dict of numpy arrays ``loss [Nt]``, ``cycles [Nt]``, ``success [Nt]``, ``Xk [Nt,Pmax]``, ``n_params [Nt]``.
        Host->device and device->host copies happen inside this call (this is what bench.py's `e2e` times).

        With ``reuse_host_buffers`` (default) the parameter table ``Xk`` is a view of a pinned host buffer owned by this
        optimizer and is overwritten by the next call; pass False to get a private (pageable) copy instead -- a fresh
        34 MB allocation per 1e5-target sweep costs ~10 ms of page faults, as much as a tenth of the whole sweep."""
        dev = engine.require_cuda()
        if isinstance(targets, torch.Tensor):
            V = targets.to(dev, non_blocking=True)
        else:
            V = torch.as_tensor(np.ascontiguousarray(targets, dtype=np.complex128)).to(dev, non_blocking=True)
        if k_range is None:
            k_range = self.basis.get_spanning_range(None)
        res = self._run_batch(V, k_range, opts)
        bx = res["best_x"]
        if reuse_host_buffers:
            hb = self._host_x
            if hb is None or hb.shape != bx.shape:
                hb = self._host_x = torch.empty(bx.shape, dtype=bx.dtype, pin_memory=True)
            hb.copy_(bx, non_blocking=True)
            torch.cuda.current_stream(bx.device).synchronize()
            xk = hb.numpy()
        else:
            xk = bx.cpu().numpy()
        return {"loss": res["best_loss"], "cycles": res["best_k"], "n_params": res["best_P"],
                "success": (res["best_loss"] <= self.success_threshold).astype(np.int32), "Xk": xk}

    def _approximate_batch(self, V: torch.Tensor) -> List[DataDictEntry]:
        b = self.basis
        Nt = V.shape[0]
        target_coords = c1c2c3_batch(V, round8=True).cpu().numpy() if b.n_qubits == 2 else None
        b.assign_seed(None)
        span = b.get_spanning_range(None)
        res = self._run_batch(V, span, keep_history=self.use_callback)
        best_x_host = res["best_x"].cpu().numpy()
        out: List[DataDictEntry] = []
        failures = []
        for i in range(Nt):
            logging.info(f"Starting sample iter {i}")
            tc = tuple(target_coords[i]) if target_coords is not None else None
            logging.info(f"Begin search: {tc}")
            best_result = float(res["best_loss"][i])
            best_cycles = int(res["best_k"][i])
            best_Xk = best_x_host[i, : int(res["best_P"][i])].copy()
            logging.info(f"Overall Best Loss={best_result}")
            success = best_result <= self.success_threshold
            if success:
                logging.info(f"Success: {tc}")
            else:
                failures.append(i)
                logging.info(f"Fail: {tc}")
            # history lists in the reference's flag format [-1, k, l0, l1, ...] (optimizer.py:238; visualize.py:90-117);
            # entries are the final loss of every restart of every k tried for this target
            if self.use_callback:
                if success or self.override_fail:
                    tl: list = []
                    coords: list = []
                    for rec in res["per_k"]:
                        if int(rec["active"][i].item()) == 0:
                            continue
                        tl.extend([-1, rec["k"]])
                        coords = []  # coordinate_list is reset per k (optimizer.py:235), training_loss is not
                        li = rec["loss"][i].tolist()
                        its = rec["iters"][i].tolist()
                        cap = rec["trace_loss"].shape[1]
                        R = len(li)
                        xs = []
                        for r in range(R):  # restarts in the reference's sequential order, up to the first success
                            if li[r] >= 1e300:
                                continue
                            n = min(int(its[r]), cap)
                            tl.extend(rec["trace_loss"][i * R + r, :n].tolist())
                            if n:
                                xs.append(rec["trace_x"][i * R + r, :n])
                            if li[r] < self.success_threshold:
                                break
                        if xs:
                            U = engine.template_eval(rec["desc"], torch.cat(xs).contiguous())
                            coords = [tuple(c) for c in c1c2c3_batch(U, round8=True).tolist()]
                    self.training_loss.append(tl)
                    self.coordinate_list.append(coords)
            else:
                self.training_loss.append(best_result)
            self.best_cycle_list.append(best_cycles)
            out.append(DataDictEntry(int(success), best_result, best_Xk, best_cycles))
        if failures and not self.override_fail:
            raise ValueError(
                "Failed to converge within error threshold. Try increasing restart attempts or increasing temperature "
                f"scaling on preseed. (targets {failures[:8]}{'...' if len(failures) > 8 else ''})")
        # leave the template at the size of the last target's best result, as the reference does on failure
        if out and out[-1].cycles > 0 and isinstance(b, _CircuitTemplateBase):
            b.build(n_repetitions=out[-1].cycles)
        return out
