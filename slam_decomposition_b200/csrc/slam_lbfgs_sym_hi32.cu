// K5 instantiations for constant symmetric gates (GM_SYM: RiSwap, ConversionGain with zero phases) -- the headline path;
// (s, y) history stored as the upper half of the double (HistHi32).
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_sym_hi32(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.extras) return dispatch_generic<GM_SYM, HistHi32, true>(kt, A, c, st);
  if (c.exact) return dispatch_exact<GM_SYM, HistHi32>(kt, A, c, st);
  return dispatch_generic<GM_SYM, HistHi32, false>(kt, A, c, st);
}

}  // namespace slam
