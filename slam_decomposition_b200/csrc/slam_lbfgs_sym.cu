// K5 instantiations for constant symmetric gates (GM_SYM), guarded vector length: non-canonical templates, box bounds,
// per-iteration trace.
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_sym_generic(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.extras) return dispatch_generic<GM_SYM, HistHi32, true>(kt, A, c, st);
  return dispatch_generic<GM_SYM, HistHi32, false>(kt, A, c, st);
}

}  // namespace slam
