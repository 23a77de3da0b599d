// K5 instantiations for constant symmetric gates (GM_SYM: RiSwap, ConversionGain with zero phases) -- the headline path:
// exact-length kernels for the canonical templates P = 6(k+1).
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_sym_generic(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st);

int lbfgs_launch_sym(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.exact && !c.extras) return dispatch_exact<GM_SYM, HistHi32>(kt, A, c, st);
  return lbfgs_launch_sym_generic(kt, A, c, st);
}

}  // namespace slam
