// K5 instantiations for constant symmetric gates (GM_SYM: RiSwap, ConversionGain with zero phases) with the (s, y)
// history stored as float (A/B alternative to the default HistHi32 kernels in slam_lbfgs_sym_hi32.cu).
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_sym_hi32(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st);

int lbfgs_launch_sym(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, int hist_kind, cudaStream_t st) {
  if (hist_kind == 1) return lbfgs_launch_sym_hi32(kt, A, c, st);
  if (c.exact) return dispatch_exact<GM_SYM, HistF32>(kt, A, c, st);
  return dispatch_generic<GM_SYM, HistF32, false>(kt, A, c, st);
}

}  // namespace slam
