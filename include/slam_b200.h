/*
 * slam_b200.h -- C ABI of libslam_b200.so, the B200 (sm_100a) engine for the SLAM
 * template-evaluation hot path.
 *
 * The reference (Pitt-JonesLab/slam_decomposition) is pure Python and has no FFI; the seam this
 * library replaces is the set of duck-typed Python calls TemplateOptimizer makes on its basis /
 * objective objects.  Each entry point below names the reference call
 * site(s) it replaces (paths relative to the reference root).  The Python package
 * `slam_decomposition_b200` binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every pointer marked [dev] is a CUDA device pointer to caller-owned, contiguous memory
 *     (e.g. torch.Tensor.data_ptr()); the library never frees or retains it.
 *   - `desc` pointers are HOST pointers to a POD descriptor, read during the call only.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous on that stream unless stated otherwise and re-entrant.  The only state the library keeps is one private,
 *     lazily created cudaMemPool_t per device for stream-ordered scratch (work counters, optimiser workspaces; at most 1 GiB
 *     of freed blocks is retained, larger workspaces return to the driver); the device's default pool is not touched.
 *   - complex128 values are (re, im) pairs of doubles; 4x4 matrices are row-major, 32 doubles.
 *   - return value: 0 = ok, negative = SlamStatus error (no exception crosses the ABI).
 *   - parameter vectors (`x`, `grad`) are in the reference's API order: the order of
 *     `QuantumCircuit.parameters`, i.e. sorted by name (basis.py:113-116); the descriptor carries
 *     the index tables, so kernels read API-ordered vectors directly.
 */
#ifndef SLAM_B200_H
#define SLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAM_ABI_VERSION 4
#define SLAM_MAX_K 16     /* max 2Q-gate applications per template (k <= 6 in the sweeps; 15 pulse slices) */
#define SLAM_MAX_SLOTS 40 /* max scalar slots of one 2Q gate (smush1q: 8 + 2T + 1, T <= 15)    */
#define SLAM_MAX_PARAMS 256

typedef enum SlamStatus {
  SLAM_OK = 0,
  SLAM_ERR_INVALID = -1,     /* bad descriptor / argument (reference: ValueError)              */
  SLAM_ERR_UNSUPPORTED = -2, /* valid but not implemented for this gate kind                   */
  SLAM_ERR_CUDA = -3,        /* CUDA runtime error; see slam_last_cuda_error()                 */
  SLAM_ERR_NO_DEVICE = -4
} SlamStatus;

/* 2Q basis-gate families (src/slam/utils/gates/custom_gates.py) */
typedef enum SlamGateKind {
  SLAM_GATE_RISWAP = 0,        /* RiSwapGate(alpha)                         custom_gates.py:582-595 */
  SLAM_GATE_CG = 1,            /* ConversionGainGate(phi_c,phi_g,gc,gg,t)   custom_gates.py:163-184 */
  SLAM_GATE_SMUSH = 2,         /* ConversionGainSmushGate(phi_c,phi_g,gc,gg,gx[T],gy[T],t) :215-250 */
  SLAM_GATE_SMUSH_1QPHASE = 3, /* ConversionGainSmush1QPhaseGate(phi_a,phi_b,phi_c,phi_g,gc,gg,gz1,gz2,gx[T],gy[T],t) :260-306 */
  SLAM_GATE_FIXED = 4          /* constant 4x4 (CanonicalGate, BerkeleyGate, any UnitaryGate)       */
} SlamGateKind;

/* cost functionals (src/slam/cost_function.py) evaluated from T = Tr(V^dag U) */
typedef enum SlamCostKind {
  SLAM_COST_BASIC = 0,         /* BasicCost        1 - |T|/4           cost_function.py:140-145 */
  SLAM_COST_SQUARE = 1,        /* SquareCost       1 - (|T|^2+4)/20    cost_function.py:169-173 */
  SLAM_COST_BASIC_INVERSE = 2, /* BasicCostInverse |T|/4               cost_function.py:133-137 */
  /* coordinate-based functionals (forward evaluation only: slam_nm_solve), on 8-dp rounded invariants as the reference */
  SLAM_COST_MAKHLIN_FUNCTIONAL = 3, /* MakhlinFunctionalCost  sum |g_i(V)-g_i(U)|^2   cost_function.py:219-221 */
  SLAM_COST_MAKHLIN_EUCLIDEAN = 4,  /* MakhlinEuclideanCost   ||g(V)-g(U)||_2         cost_function.py:209-216 */
  SLAM_COST_WEYL_EUCLIDEAN = 5,     /* WeylEuclideanCost      ||c(V)-c(U)||_2         cost_function.py:199-206 */
  SLAM_COST_BASIC_REDUCED = 6,      /* BasicReducedCost  Basic on canonical_gate(c(.)) cost_function.py:176-182 */
  SLAM_COST_SQUARE_REDUCED = 7      /* SquareReducedCost                               cost_function.py:185-189 */
} SlamCostKind;

/*
 * Template structure after `basis.build(k)`: 1Q layer 0, then k x (2Q gate, 1Q layer).
 * Replaces CircuitTemplate.__init__/build/_build_cycle (basis.py:52-93,124-169) and the V2 twin
 * (basisv2.py:27-299), which hold the same information as a qiskit QuantumCircuit.
 */
typedef struct SlamTemplateDesc {
  int32_t gate_kind;      /* SlamGateKind                                                        */
  int32_t k;              /* number of 2Q gate applications, 1..SLAM_MAX_K                        */
  int32_t T;              /* time slices of a smush gate (len(gx)); 0 otherwise                   */
  int32_t n_slots;        /* scalar slots per 2Q gate (1, 5, 5+2T, 9+2T, 0)                       */
  int32_t n_params;       /* P = len(Xk)                                                          */
  int32_t no_exterior_1q; /* informational; absent layers are marked by p1q = -1                  */
  int32_t vz_only;        /* 1: 1Q gates are RZ(lam) (basisv2.py:267-272); p1q[i][0], p1q[i][3]   */
  int32_t reserved;
  /* p1q[i][0..2] = Xk index of (theta,phi,lam) of the U gate on qubit 0 in layer i,
     p1q[i][3..5] = same for qubit 1; -1 in [i][0] and [i][3] = that 1Q gate is absent.          */
  int32_t p1q[SLAM_MAX_K + 1][6];
  /* slot_param[g][s] = Xk index bound to slot s of the g-th 2Q gate, or -1 if the slot is the
     constant slot_const[g][s].                                                                  */
  int32_t slot_param[SLAM_MAX_K][SLAM_MAX_SLOTS];
  double slot_const[SLAM_MAX_K][SLAM_MAX_SLOTS];
  double fixed_gate[32];  /* SLAM_GATE_FIXED: the matrix, row-major (re,im)                       */
} SlamTemplateDesc;

/* ---- housekeeping ------------------------------------------------------------------------- */
int slam_abi_version(void);
const char* slam_status_string(int status);
const char* slam_last_cuda_error(void); /* thread-local text of the last CUDA failure           */
int slam_device_count(void);
int slam_set_device(int device); /* cudaSetDevice for the calling thread (one process per GPU)   */

/*
 * K1  U[b] = template(x[b])                                  [B,4,4] complex128
 * Replaces CircuitTemplate.eval / assign_Xk + qiskit Operator(circuit).data (basis.py:102-116),
 * CircuitTemplateV2.eval (basisv2.py:143-145), and every gate __array__ they reach.
 *   x   [dev] double[B, ldx]  (ldx >= n_params, row b at x + b*ldx)
 *   U   [dev] double[B, 32]
 */
int slam_template_eval(const SlamTemplateDesc* desc, const double* x, int64_t ldx, double* U, int64_t B,
                       void* stream);

/*
 * K2  loss[b], grad[b,:] of cost(template(x[b]), V[tgt[b]])  -- the unit of the headline metric.
 * Replaces objective_func (optimizer.py:191-214) = basis.eval + objective.unitary_fidelity
 * (cost_function.py:133-173) and scipy's (P+1)-evaluation finite-difference gradient
 * (opt.minimize(jac=None), optimizer.py:270-278) by one analytic adjoint pass.
 *   V       [dev] double[Nt, 32]   targets
 *   tgt_idx [dev] int32[B] or NULL (NULL: target of row b is b % Nt)
 *   loss    [dev] double[B]
 *   grad    [dev] double[B, ldg] or NULL (loss only)
 *   trace   [dev] double[B, 2]  or NULL: T = Tr(V^dag U)
 * Templates with parameter-bound ConversionGainSmush / ConversionGainSmush1QPhase gates (hamiltonian.py:114-182) are
 * differentiated through every time slice exp(-i dt H): Hermitian eigen-decomposition of H per slice and the
 * Daleckii-Krein divided differences give d loss / d (amplitudes, phases, couplings, Z terms, duration) in one
 * backward pass (one thread per row).
 */
int slam_loss_grad(const SlamTemplateDesc* desc, const double* x, int64_t ldx, const double* V, int64_t Nt,
                   const int32_t* tgt_idx, int32_t cost_kind, double* loss, double* grad, int64_t ldg,
                   double* trace, int64_t B, void* stream);
/* as slam_loss_grad with an explicit team width: lanes per row, 1, 2 or 4 (0 = automatic: 2 up to k = 4, 4 beyond);
   ignored for templates with parameter-bound smush gates (one thread per row) */
int slam_loss_grad_lanes(const SlamTemplateDesc* desc, const double* x, int64_t ldx, const double* V, int64_t Nt,
                         const int32_t* tgt_idx, int32_t cost_kind, double* loss, double* grad, int64_t ldg,
                         double* trace, int64_t B, int32_t lanes, void* stream);

/*
 * K3  Weyl-chamber coordinates and Makhlin invariants of a batch of 4x4 unitaries.
 * Replaces weylchamber.c1c2c3 / g1g2g3 as called from basis_abc.py:80-84, optimizer.py:85,103,224,
 * cost_function.py:199-221, parallel_drive_volume.py:225, pd_playground.py:199.
 *   U [dev] double[B,32];  c [dev] double[B,3] or NULL;  g [dev] double[B,3] or NULL
 */
#define SLAM_WEYL_FOLD 1   /* c1 > 1/2 -> 1 - c1 (pd_playground.py:199-202)                      */
#define SLAM_WEYL_ROUND8 2 /* round to 8 decimals as weylchamber does                            */
int slam_weyl(const double* U, int64_t B, double* c, double* g, int32_t flags, void* stream);

/*
 * K5  Batched, device-resident L-BFGS over (target, restart) pairs for ONE template size k.
 * Replaces the restart loop around opt.minimize(method="BFGS") (optimizer.py:253-295); the caller
 * (TemplateOptimizer._run shim) iterates k ascending and carries `best_loss` across calls, which
 * reproduces the k-loop early exit (optimizer.py:233,297-303).
 */
typedef struct SlamOptOpts {
  int32_t max_iter;      /* per restart; reference: options={"maxiter": 2500}                    */
  int32_t history;       /* L-BFGS pairs kept (<= 8); 0 = auto (slam_lbfgs_solve: from the shared-memory budget;
                            slam_fd_lbfgs_solve adjoint mode: 8 for P <= 20, 6 for P <= 32, else 5; its finite-difference modes always keep 8)           */
  int32_t cost_kind;     /* SlamCostKind                                                         */
  int32_t early_exit;    /* 1: other restarts of a target stop once one is < success_threshold   */
  double success_threshold; /* reference SUCCESS_THRESHOLD = 1e-10 (optimizer.py:18)             */
  double f_stop;         /* stop a restart once loss < f_stop (polish below the threshold)       */
  double gtol;           /* stop when max|g| < gtol ...                                          */
  double gtol_far;       /* ... or max|g| < gtol_far while loss > f_far (a non-zero local min)   */
  double f_far;
  double x0_lo, x0_hi;   /* when x0 == NULL: x0 ~ U[x0_lo, x0_hi) from Philox(seed)              */
  /* optional per-iteration trace (replaces callbackF, optimizer.py:217-224): after accepted iteration i (1-based,
     i <= trace_cap) of problem p = target*restarts + restart the kernel stores the loss in
     trace_loss[p*trace_cap + i-1] and the parameters in trace_x[(p*trace_cap + i-1)*P ...]; out_iters gives the
     number of iterations.  [dev] pointers, NULL / 0 = off.                                                      */
  int32_t trace_cap;
  int32_t diag;          /* slam_fd_lbfgs_solve diagnostics, 0 = off: 1 = out_iters carries the stop reason in bits 24-31
                            (1 f_stop, 2 gtol, 3 gtol_far, 4 max_iter, 5 non-finite, 6 target solved elsewhere, 7 no feasible
                            descent, 8 line search exhausted, 9 stalled: < 3 % progress over 32 iterations at > 6 evaluations each); 2 = as 1 with the problem's evaluation count
                            (instead of its iterations) in bits 0-23                                                        */
  double* trace_loss;
  double* trace_x;
  /* optional box bounds, [dev] double[P] each: projected L-BFGS, replacing the reference's switch to scipy L-BFGS-B
     when basis.using_bounds (optimizer.py:257-258, basisv2.py:174-190).  Both arrays or neither (SLAM_ERR_INVALID
     otherwise); a side without a bound is +-inf.  Initial points are clamped into the box.                           */
  const double* lower;
  const double* upper;
  /* optional chaining of launches over ascending template sizes k (slam_lbfgs_solve only), so that the launch for k+1 can
     be enqueued on a second stream while the launch for k drains and its CTAs fill the SMs the tail of k leaves idle:
       solved_out [dev] int32[Nt], zeroed by the caller: flag t is set when a restart of target t ends below
                  success_threshold in THIS launch, or when t was skipped because of solved_in (flags are cumulative);
                  NULL = internal scratch.  Requires early_exit = 1.
       solved_in  [dev] int32[Nt] or NULL: the solved_out array of the launch for the previous size; it may still be
                  written while this launch runs.  Targets flagged there when one of their restarts is fetched are skipped
                  (out_loss = DBL_MAX), which reproduces the k-loop early exit of optimizer.py:297-303 without a host
                  round trip.  A target whose last restarts at size k succeed after size k+1 has started on it costs
                  wasted work only: the caller keeps the smallest successful k.                                       */
  const int32_t* solved_in;
  int32_t* solved_out;
  /* optional circuit-cost constraint (slam_fd_lbfgs_solve only, finite-difference modes), replacing the reference's switch to
     scipy SLSQP when basis.using_constraints (optimizer.py:259-264; CircuitTemplateV2.set_constraint, basisv2.py:192-203):
       circuit_cost(x) = sum over 2Q gates of  alpha                      (RiSwap,            custom_gates.py:568-572)
                                              (|gc| + |gg|) t / (pi/2)    (ConversionGain and its smush form, :208-212, :252-257)
     enters the objective as the augmented-Lagrangian term (con_mu / 2) max(0, circuit_cost(x) - con_max + lambda / con_mu)^2
     with one multiplier per (target, restart) problem; the caller runs the outer multiplier iteration.
       con_mu     penalty weight; 0 = no constraint
       con_lambda [dev] double[Nt * restarts] multipliers, or NULL (all zero)                                              */
  double con_max;
  double con_mu;
  const double* con_lambda;
  /* optional per-TARGET reduction over restarts AND over the chained launches of ascending template sizes (slam_lbfgs_solve
     only) -- the merge the reference does on the host, optimizer.py:283-303: "the smallest k that reached the threshold, else
     the lowest loss seen".  Every retired restart issues one fire-and-forget 64-bit atomicMin on best_key[target] with
         bit 63 = loss >= success_threshold | bits 62-59 = k if below the threshold, else 0 | bits 58-12 = top 47 bits of
         the loss (order preserving) | bits 11-8 = k | bits 7-0 = restart
     so the minimum identifies the winning (k, restart); slam_best_gather() then copies the winner's loss and parameters out
     of the per-restart tables.  Non-finite losses never enter.  Needs restarts <= 256 and k <= 15.
       best_key [dev] uint64[Nt], caller initialises to all ones (0xFFFF...F = no result yet)                            */
  unsigned long long* best_key;
  /* launch tuning of slam_lbfgs_solve, 0 = automatic (explicit fields, so A/B measurements need no environment variables):
       tune_lanes       lanes per problem, 2 or 4
       tune_sm_threads  register class for canonical templates with P <= 24: 512 (16 warps/SM, 128 registers) or 384
       tune_hist_min    smallest history length the automatic choice may shrink to in order to fit more teams
       tune_max_teams   cap on the teams (problems in flight) per SM                                                     */
  int32_t tune_lanes;
  int32_t tune_sm_threads;
  int32_t tune_hist_min;
  int32_t tune_max_teams;
} SlamOptOpts;

void slam_opt_defaults(SlamOptOpts* o);

/*
 *   V          [dev] double[Nt,32] targets
 *   x0         [dev] double[Nt, restarts, ldx0] or NULL (then Philox4x32-10 keyed by seed)
 *   active     [dev] int32[Nt] or NULL: targets with active[t] == 0 are skipped (already solved)
 *   out_loss   [dev] double[Nt, restarts]   final loss of every restart (DBL_MAX = skipped: target inactive / already
 *                                           solved; NaN = the objective went non-finite)
 *   out_x      [dev] double[Nt, restarts, P] final parameters of every restart
 *   out_iters  [dev] int32 [Nt, restarts]   L-BFGS iterations used

 *   out_evals  [dev] int64 [1] or NULL      += number of loss+grad evaluations performed
 */
int slam_lbfgs_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts,
                     const double* x0, int64_t ldx0, uint64_t seed, const int32_t* active,
                     const SlamOptOpts* opts, double* out_loss, double* out_x, int32_t* out_iters,
                     unsigned long long* out_evals, void* stream);

/*
 * Winners of the per-target reduction (SlamOptOpts.best_key) copied out of the per-restart tables of the chained launches:
 * for target t with key (k, r): best_loss[t] = loss_k[t, r], best_k[t] = k, best_P[t] = P_k, best_x[t, :P_k] = x_k[t, r, :]
 * (zero padded to ldx); targets without a finite result get (+inf, -1, 0, zeros).  Replaces the host-side bookkeeping of
 * optimizer.py:283-303 (`best_result`, `best_Xk`, `best_cycles`).
 *   k_of_size, P_of_size  HOST int32[n_sizes]: template size k and parameter count of every launch of the chain
 *   loss_of_size, x_of_size  HOST arrays of n_sizes [dev] pointers: out_loss [Nt, restarts] and out_x [Nt, restarts, P] of
 *                            that launch
 */
int slam_best_gather(const unsigned long long* best_key, int64_t Nt, int32_t restarts, int32_t n_sizes,
                     const int32_t* k_of_size, const int32_t* P_of_size, const double* const* loss_of_size,
                     const double* const* x_of_size, double* best_loss, int32_t* best_k, int32_t* best_P, double* best_x,
                     int64_t ldx, void* stream);

/*
 * K5c Batched L-BFGS with FINITE-DIFFERENCE gradients over the generic forward objective: the reference's own algorithm
 * class (scipy BFGS with jac=None: P forward differences of step 1.49e-8 per gradient, optimizer.py:270-278) for the
 * templates whose gates have no closed-form derivative here -- parameter-bound ConversionGainSmush /
 * ConversionGainSmush1QPhase gates (hamiltonian.py:114-182) -- and for BasicCostInverse x circuit_fidelity
 * (optimizer.py:200-201).  cost_kind must be trace based (BASIC, SQUARE, BASIC_INVERSE).
 *   central = 0: forward differences (scipy's jac=None);  1: central differences (2P evaluations per gradient, step 6e-6);
 *   central = 2: ANALYTIC adjoint gradient through the smush slices (as slam_loss_grad; smush templates only, else
 *                SLAM_ERR_UNSUPPORTED): one backward pass (~3 forward evaluations of work) instead of P + 1 evaluations.
 * Box bounds (opts->lower/upper) by projection (adjoint mode: the quasi-Newton direction is computed in the free subspace of
 * the current active set, with a projected steepest-descent fallback; a restart ends on a bound only at a KKT point); the
 * trace fields of opts are ignored; opts->history and opts->diag apply (see SlamOptOpts).  Other arguments as
 * slam_lbfgs_solve; out_evals counts forward evaluations (one per gradient when central = 2).
 */
int slam_fd_lbfgs_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts,
                        const double* x0, int64_t ldx0, uint64_t seed, const int32_t* active,
                        const SlamOptOpts* opts, int32_t central, double* out_loss, double* out_x,
                        int32_t* out_iters, unsigned long long* out_evals, void* stream);

/*
 * K5b Batched Nelder-Mead with a generic forward objective: any template (incl. parameter-bound smush gates) and any
 * SlamCostKind.  Replaces opt.minimize(method="Nelder-Mead") reached through TemplateOptimizer(override_method=...)
 * (optimizer.py:266-278).  Simplex rules, initial simplex and the xatol/fatol termination follow scipy.
 * Arguments as slam_lbfgs_solve.
 */
typedef struct SlamNmOpts {
  int32_t max_iter;   /* reference passes options={"maxiter": 2500}                                              */
  int32_t cost_kind;  /* any SlamCostKind                                                                        */
  int32_t early_exit;
  int32_t reserved;
  double success_threshold;
  double xatol, fatol; /* scipy defaults 1e-4, 1e-4                                                              */
  double x0_lo, x0_hi;
} SlamNmOpts;

void slam_nm_defaults(SlamNmOpts* o);

int slam_nm_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts, const double* x0,
                  int64_t ldx0, uint64_t seed, const int32_t* active, const SlamNmOpts* opts, double* out_loss,
                  double* out_x, int32_t* out_iters, unsigned long long* out_evals, void* stream);

/*
 * K6  Fused coverage-set Monte-Carlo: Philox -> params -> template -> c1c2c3 -> fold -> bin.
 * Replaces the N-iteration loop of parallel_drive_volume.py:209-225 and the mirror fold :292-307.
 *   params[j] = lo + (hi - lo) * u53(seed, sample, j)   (j in API order; see oracle.philox_uniform)
 *   hist   [dev] int64[nbins^3]  (+= counts; caller zeroes)   bins over folded [0,1/2]^3
 *   coords [dev] double[n_samples,3] or NULL: un-rounded folded coordinates of every sample
 */
int slam_coverage_mc(const SlamTemplateDesc* desc, uint64_t seed, int64_t first_sample, int64_t n_samples,
                     double lo, double hi, int32_t nbins, unsigned long long* hist, double* coords,
                     void* stream);

/*
 * K4b Weyl trajectory of a parallel-driven gate: N slices of ConversionGainSmush1QPhase, R
 * sub-times each, as a prefix product.  Replaces ParallelDrivenGateWidget.iterate_time / solve_end
 * (pd_playground.py:169-208).  One trajectory per batch row.
 *   gate   [dev] double[B, 8]  (phi_a,phi_b,phi_c,phi_g,gc,gg,gz1,gz2)
 *   gx,gy  [dev] double[B, N]
 *   coords [dev] double[B, N, R, 3] (folded; flags as slam_weyl) or NULL
 *   Ufinal [dev] double[B, 32] or NULL
 */
int slam_pd_trajectory(const double* gate, const double* gx, const double* gy, int32_t N, int32_t R, double dt,
                       int32_t flags, double* coords, double* Ufinal, int64_t B, void* stream);
/* Multi-segment pulses: as slam_pd_trajectory with one gate row PER SLICE, gate [dev] double[B, N, 8] -- the drive phases
   and couplings may change from slice to slice.  Replaces the composed widgets `pdgw + pdgw2 (+ pdgw3)` of
   ParallelDrivenGateWidget.__add__ (pd_playground.py:46-58) as used by scripts/parallel_drive_swap cells 6-13.          */
int slam_pd_trajectory_slices(const double* gate, const double* gx, const double* gy, int32_t N, int32_t R, double dt,
                              int32_t flags, double* coords, double* Ufinal, int64_t B, void* stream);

/*
 * Diagnostic: register-resident DFMA loop; writes achieved FP64 FLOP/s of the current device to
 * *flops (host pointer).  Synchronous.  Used as the roofline denominator (MEASURED_PEAKS.json has
 * no FP64 figure).
 */
int slam_fp64_peak(int32_t iters, double* flops, double* ms);

/*
 * Diagnostic: HOST evaluation of the (cos, sin) routine the kernels use for the U3 angles (Cody-Waite reduction +
 * fdlibm kernel polynomials, coefficients in the constant bank on the device), so that its accuracy can be checked
 * against libm without a GPU.  x, s, c are host pointers to n doubles.  No device work.
 */
int slam_selftest_sincos(const double* x, int64_t n, double* s, double* c);

#ifdef __cplusplus
}
#endif
#endif /* SLAM_B200_H */
