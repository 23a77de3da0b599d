// K5 instantiations for GM_BLOCK templates (slam_core.cuh).
#include "slam_lbfgs.cuh"

namespace slam {

int lbfgs_launch_block(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.extras) return dispatch_generic<GM_BLOCK, HistHi32, true>(kt, A, c, st);
  return dispatch_generic<GM_BLOCK, HistHi32, false>(kt, A, c, st);
}

}  // namespace slam
