// slam_lbfgs.cu -- K5 host side: launch configuration and the C-ABI entry point of the batched L-BFGS
// (kernel template in slam_lbfgs.cuh, instantiated per gate mode in slam_lbfgs_{sym,sym_hi32,block,dense}.cu).
#include <algorithm>
#include <cstdlib>

#include "slam_lbfgs.cuh"

namespace slam {

// Bank tiling: in one shared-memory access each team touches LPP consecutive elements, so the teams that share a
// 128-byte wavefront are conflict-free when the team stride is an odd multiple of LPP elements (mod 128 B).
static int tile_stride(int n, int elem_bytes, int lpp) {
  const int W = 128 / elem_bytes;
  while ((n % W) % (2 * lpp) != lpp) ++n;
  return n;
}

// doubles per team: vectors, alp (padded to an even count: the trig cache is read as double2), (cos, sin) cache
static int team_doubles(const KTemplate& kt, int Pp, int m, int lpp) {
  return tile_stride(4 * Pp + m + (m & 1) + 2 * kt.n_trig, 8, lpp);
}

// history elements per team (second region, own stride): S, Y and rho
static int team_hist_elems(int Pp, int m, int hist_bytes, int lpp) {
  return tile_stride(2 * m * Pp + m, hist_bytes, lpp);
}

static size_t team_bytes(const KTemplate& kt, int Pp, int m, int hist_bytes, int lpp) {
  return (size_t)team_doubles(kt, Pp, m, lpp) * 8 + (size_t)team_hist_elems(Pp, m, hist_bytes, lpp) * hist_bytes;
}

// canonical template: full U3 layers around one constant symmetric gate, P = 6(k+1)
static bool is_canonical(const KTemplate& kt) {
  if (kt.gmode != GM_SYM || kt.vz_only || kt.P != 6 * (kt.k + 1)) return false;
  for (int i = 0; i <= kt.k; ++i)
    for (int s = 0; s < 6; ++s)
      if (kt.p1q[i][s] < 0) return false;
  for (int g = 1; g < kt.k; ++g)
    for (int c = 0; c < 4; ++c)
      if (kt.gsym[g][c] != kt.gsym[0][c]) return false;
  return true;
}

bool lbfgs_has_exact(int lpp, int npl) {
  if (lpp == 4) return npl == 3 || npl == 5 || npl == 6 || npl == 8 || npl == 9 || npl == 11;
  if (lpp == 2) return npl == 6 || npl == 9 || npl == 12 || npl == 15 || npl == 18 || npl == 21;
  return false;
}

}  // namespace slam

using namespace slam;

extern "C" void slam_opt_defaults(SlamOptOpts* o) {
  if (!o) return;
  o->max_iter = 2500;            // optimizer.py:274 options={"maxiter": 2500}
  o->history = 0;
  o->cost_kind = SLAM_COST_BASIC;
  o->early_exit = 1;
  o->success_threshold = 1e-10;  // optimizer.py:18
  o->f_stop = 1e-13;
  o->gtol = 1e-9;
  o->gtol_far = 1e-5;            // scipy BFGS default gtol, applied to restarts stuck at a non-zero local minimum
  o->f_far = 1e-6;
  o->x0_lo = 0.0;                // basis.py:111: np.random.random(P) * 2 pi
  o->x0_hi = 6.283185307179586;
  o->trace_cap = 0;
  o->diag = 0;
  o->trace_loss = nullptr;
  o->trace_x = nullptr;
  o->lower = nullptr;
  o->upper = nullptr;
  o->solved_in = nullptr;
  o->solved_out = nullptr;
  o->con_max = 0.0;
  o->con_mu = 0.0;
  o->con_lambda = nullptr;
  o->best_key = nullptr;
  o->tune_lanes = 0;
  o->tune_sm_threads = 0;
  o->tune_hist_min = 0;
  o->tune_max_teams = 0;
}

extern "C" int slam_lbfgs_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts,
                                const double* x0, int64_t ldx0, uint64_t seed, const int32_t* active,
                                const SlamOptOpts* opts, double* out_loss, double* out_x, int32_t* out_iters,
                                unsigned long long* out_evals, void* stream) {
  if (!desc || !V || !opts || !out_loss || !out_x || !out_iters || Nt < 0 || restarts < 1) return SLAM_ERR_INVALID;
  if (x0 && ldx0 < desc->n_params) return SLAM_ERR_INVALID;
  if (opts->cost_kind != SLAM_COST_BASIC && opts->cost_kind != SLAM_COST_SQUARE) return SLAM_ERR_UNSUPPORTED;
  if (opts->max_iter < 1 || opts->history < 0 || opts->history > kMaxHist) return SLAM_ERR_INVALID;
  if (desc->n_params < 1) return SLAM_ERR_INVALID;
  if (opts->con_mu != 0.0) return SLAM_ERR_UNSUPPORTED;  // cost-constrained runs go through slam_fd_lbfgs_solve
  if ((opts->solved_in || opts->solved_out) && !opts->early_exit) return SLAM_ERR_INVALID;
  if (opts->solved_in && !opts->solved_out) return SLAM_ERR_INVALID;  // the chain needs somewhere to propagate to
  if ((opts->lower == nullptr) != (opts->upper == nullptr)) return SLAM_ERR_INVALID;  // box = both arrays (+-inf allowed)
  if (opts->best_key && (restarts > 256 || desc->k > 15)) return SLAM_ERR_UNSUPPORTED;  // key packs 8 + 4 bits
  if (opts->tune_lanes != 0 && opts->tune_lanes != 2 && opts->tune_lanes != 4) return SLAM_ERR_INVALID;
  if (opts->tune_sm_threads != 0 && opts->tune_sm_threads != 512 && opts->tune_sm_threads != 384) return SLAM_ERR_INVALID;
  if (opts->tune_hist_min < 0 || opts->tune_max_teams < 0) return SLAM_ERR_INVALID;
  if (Nt == 0) return SLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/false);
  if (rc != SLAM_OK) return rc;
  if (kt.P > 96) return SLAM_ERR_UNSUPPORTED;
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  int dev = 0, sms = 0, max_smem = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // ---- launch configuration -------------------------------------------------------------------
  // lanes per problem: 4 (one matrix column per lane, 12 warps/SM) or 2 (two columns per lane: twice the ILP and
  // half the replicated scalar work per problem, 8 warps/SM).  SlamOptOpts.tune_* override the automatic choices (A/B runs).
  const bool extras = opts->lower || (opts->trace_loss && opts->trace_cap > 0);
  int lpp = opts->tune_lanes ? opts->tune_lanes : 4;
  if (lpp == 2 && kt.P > 56) lpp = 4;
  const int hb = 4;  // history element: upper half of the double (HistHi32)
  LbfgsCfg cfg;
  cfg.lpp = lpp;
  cfg.extras = extras ? 1 : 0;
  cfg.npl = (kt.P + lpp - 1) / lpp;
  cfg.exact = (is_canonical(kt) && !extras && lbfgs_has_exact(lpp, cfg.npl)) ? 1 : 0;
  int Pp = lpp * cfg.npl;
  if (!cfg.exact) {
    Pp = (kt.P + 3) & ~3;
    cfg.npl = Pp / lpp;
  }
  // history length and teams per CTA from the shared-memory budget (one persistent CTA per SM)
  const int tpw = 32 / lpp;  // teams per warp
  cfg.maxt = lpp == 4 ? kMaxT4 : kMaxT2;
  // 16 warps/SM (512 threads, 128 registers) where the exact kernels exist and 128 teams fit in shared memory with a
  // history of >= 3 pairs (P <= 24); measured per k against 12 warps with m = 5..6: k=1 19.9 -> 17.5 ms, k=2 42.5 -> 38.3,
  // k=3 38.3 -> 36.2 per 1e5-target launch (the shorter history costs 9-14 % more evaluations and still wins)
  if (lpp == 4 && cfg.exact && cfg.npl <= 6 && opts->tune_sm_threads != 384) cfg.maxt = kMaxT4x;
  int max_teams = cfg.maxt / lpp;
  {
    const int cap = opts->tune_max_teams;
    if (cap >= tpw && cap < max_teams) max_teams = cap / tpw * tpw;
  }
  const int m_min = std::max(1, std::min(opts->tune_hist_min ? opts->tune_hist_min : (cfg.maxt == kMaxT4x ? 3 : 4), 6));
  int m = opts->history ? opts->history : 6;
  int teams = max_teams;
  if (!opts->history)  // prefer a full complement of teams (occupancy) over a longer history, down to m_min
    while (m > m_min && team_bytes(kt, Pp, m, hb, lpp) * max_teams > (size_t)max_smem) --m;
  while (teams > tpw && team_bytes(kt, Pp, m, hb, lpp) * teams > (size_t)max_smem) teams -= tpw;
  const int RS = team_doubles(kt, Pp, m, lpp);
  const int HS = team_hist_elems(Pp, m, hb, lpp);
  const size_t smem = team_bytes(kt, Pp, m, hb, lpp) * teams;
  if (smem > (size_t)max_smem) return SLAM_ERR_UNSUPPORTED;

  // stream-ordered scratch: work counter + per-target early-exit flags
  Scratch scratch(st);
  unsigned long long* next = nullptr;
  int32_t* solved = opts->solved_out;  // caller-owned (and caller-zeroed) flags of a chained launch, else scratch
  if ((rc = scratch.alloc(&next, sizeof(unsigned long long), true)) != SLAM_OK) return rc;
  if (!solved && (rc = scratch.alloc(&solved, sizeof(int32_t) * (size_t)Nt, true)) != SLAM_OK) return rc;

  LbfgsArgs A;
  A.V = V; A.x0 = x0; A.ldx0 = ldx0; A.seed = seed; A.active = active; A.Nt = Nt; A.restarts = restarts;
  A.m = m; A.RS = RS; A.HS = HS; A.Pp = Pp; A.max_iter = opts->max_iter; A.cost_kind = opts->cost_kind; A.early_exit = opts->early_exit;
  A.success_threshold = opts->success_threshold; A.f_stop = opts->f_stop; A.gtol = opts->gtol;
  A.gtol_far = opts->gtol_far; A.f_far = opts->f_far; A.x0_lo = opts->x0_lo; A.x0_span = opts->x0_hi - opts->x0_lo;
  A.trace_cap = (opts->trace_loss && opts->trace_cap > 0) ? opts->trace_cap : 0;
  A.trace_loss = opts->trace_loss; A.trace_x = opts->trace_x;
  A.lower = opts->lower;
  A.upper = opts->upper;
  A.out_loss = out_loss; A.out_x = out_x; A.out_iters = out_iters; A.out_evals = out_evals;
  A.next = next; A.solved = solved; A.solved_in = opts->solved_in; A.best_key = opts->best_key;

  const int64_t total = Nt * (int64_t)restarts;
  cfg.grid = (int)std::min<int64_t>((int64_t)sms, (total + teams - 1) / teams);
  cfg.threads = teams * lpp;
  cfg.smem = smem;
  switch (kt.gmode) {
    case GM_SYM: rc = lbfgs_launch_sym(kt, A, cfg, st); break;
    case GM_BLOCK: rc = lbfgs_launch_block(kt, A, cfg, st); break;
    case GM_DENSE: rc = lbfgs_launch_dense(kt, A, cfg, st); break;
    default: rc = SLAM_ERR_UNSUPPORTED;
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// slam_best_gather: read the winners of the packed per-target reduction back out of the per-restart tables
// ------------------------------------------------------------------------------------------------------------------
namespace slam {

struct GatherArgs {
  const unsigned long long* key;
  const double* loss[16];  // per template size k (index k): [Nt, R] tables, or null
  const double* x[16];     //                                [Nt, R, P[k]]
  int P[16];
  int64_t Nt;
  int R;
  double* best_loss;
  int32_t* best_k;
  int32_t* best_P;
  double* best_x;
  int64_t ldx;
};

__global__ void __launch_bounds__(256) best_gather_kernel(const __grid_constant__ GatherArgs G) {
  // one warp per target: lanes copy the winner's parameter row (coalesced), zero padded to ldx
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= G.Nt) return;
  const unsigned long long key = G.key[t];
  int k = -1, P = 0;
  double f = __longlong_as_double(0x7ff0000000000000LL);  // +inf: no finite result for this target
  const double* row = nullptr;
  if (key != ~0ULL) {
    k = (int)((key >> 8) & 15);
    const int r = (int)(key & 255);
    if (G.loss[k] && r < G.R) {
      f = G.loss[k][t * G.R + r];
      P = G.P[k];
      row = G.x[k] + (t * G.R + r) * (int64_t)P;
    } else {
      k = -1;
    }
  }
  for (int64_t j = lane; j < G.ldx; j += 32) G.best_x[t * G.ldx + j] = (j < P) ? row[j] : 0.0;
  if (lane == 0) {
    G.best_loss[t] = f;
    G.best_k[t] = k;
    G.best_P[t] = P;
  }
}

}  // namespace slam

extern "C" int slam_best_gather(const unsigned long long* best_key, int64_t Nt, int32_t restarts, int32_t n_sizes,
                                const int32_t* k_of_size, const int32_t* P_of_size, const double* const* loss_of_size,
                                const double* const* x_of_size, double* best_loss, int32_t* best_k, int32_t* best_P,
                                double* best_x, int64_t ldx, void* stream) {
  if (!best_key || !k_of_size || !P_of_size || !loss_of_size || !x_of_size || !best_loss || !best_k || !best_P || !best_x)
    return SLAM_ERR_INVALID;
  if (Nt < 0 || restarts < 1 || restarts > 256 || n_sizes < 1 || n_sizes > 16) return SLAM_ERR_INVALID;
  GatherArgs G;
  memset(&G, 0, sizeof(G));
  for (int i = 0; i < n_sizes; ++i) {
    const int k = k_of_size[i];
    if (k < 1 || k > 15 || P_of_size[i] < 1 || P_of_size[i] > ldx || !loss_of_size[i] || !x_of_size[i]) return SLAM_ERR_INVALID;
    G.loss[k] = loss_of_size[i];
    G.x[k] = x_of_size[i];
    G.P[k] = P_of_size[i];
  }
  if (Nt == 0) return SLAM_OK;
  G.key = best_key; G.Nt = Nt; G.R = restarts;
  G.best_loss = best_loss; G.best_k = best_k; G.best_P = best_P; G.best_x = best_x; G.ldx = ldx;
  const int64_t blocks = (Nt * 32 + 255) / 256;
  best_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(G);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}
