// placeholder until K4 lands
#include "slam_host.h"
namespace slam {
int smush_eval_launch(const KTemplate&, const double*, int64_t, double*, int64_t, cudaStream_t) { return SLAM_ERR_UNSUPPORTED; }
int lower_const_smush(const SlamTemplateDesc*, KTemplate*, cudaStream_t) { return SLAM_ERR_UNSUPPORTED; }
}
// temporary stubs (replaced as the kernels land)
extern "C" {
int slam_coverage_mc(const SlamTemplateDesc*, uint64_t, int64_t, int64_t, double, double, int32_t, unsigned long long*, double*, void*) { return SLAM_ERR_UNSUPPORTED; }
int slam_pd_trajectory(const double*, const double*, const double*, int32_t, int32_t, double, int32_t, double*, double*, int64_t, void*) { return SLAM_ERR_UNSUPPORTED; }
}
