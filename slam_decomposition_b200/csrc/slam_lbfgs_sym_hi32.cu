// K5 instantiations for GM_SYM with the (s, y) history stored as the upper half of the double (HistHi32).
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_sym_hi32(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.exact) return dispatch_exact<GM_SYM, HistHi32>(kt, A, c, st);
  return dispatch_generic<GM_SYM, HistHi32, false>(kt, A, c, st);
}

}  // namespace slam
