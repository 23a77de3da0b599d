// slam_api.cu -- housekeeping entry points and descriptor lowering for libslam_b200.so
#include <mutex>

#include "slam_host.h"

namespace slam {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int scratch_pool(int device, cudaMemPool_t* pool) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  static bool made[64] = {};
  if (device < 0 || device >= 64) return SLAM_ERR_INVALID;
  std::lock_guard<std::mutex> lock(mu);
  if (!made[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    SLAM_CUDA_CHECK(cudaMemPoolCreate(&pools[device], &props));
    unsigned long long keep = 1ULL << 30;
    SLAM_CUDA_CHECK(cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep));
    made[device] = true;
  }
  *pool = pools[device];
  return SLAM_OK;
}

static int expected_slots(int kind, int T) {
  switch (kind) {
    case SLAM_GATE_RISWAP: return 1;
    case SLAM_GATE_CG: return 5;
    case SLAM_GATE_SMUSH: return 5 + 2 * T;
    case SLAM_GATE_SMUSH_1QPHASE: return 9 + 2 * T;
    case SLAM_GATE_FIXED: return 0;
    default: return -1;
  }
}

int compile_template(const SlamTemplateDesc* d, KTemplate* kt, bool allow_bound_smush, bool allow_ties) {
  if (!d || !kt) return SLAM_ERR_INVALID;
  if (d->k < 1 || d->k > SLAM_MAX_K) return SLAM_ERR_INVALID;  // basis.py:127-128 raises ValueError for k <= 0
  if (d->n_params < 0 || d->n_params > SLAM_MAX_PARAMS) return SLAM_ERR_INVALID;
  const bool smush = d->gate_kind == SLAM_GATE_SMUSH || d->gate_kind == SLAM_GATE_SMUSH_1QPHASE;
  if (smush && (d->T < 1 || expected_slots(d->gate_kind, d->T) > SLAM_MAX_SLOTS)) return SLAM_ERR_INVALID;
  const int ns = expected_slots(d->gate_kind, smush ? d->T : 0);
  if (ns < 0 || ns != d->n_slots) return SLAM_ERR_INVALID;

  memset(kt, 0, sizeof(*kt));
  kt->k = d->k;
  kt->P = d->n_params;
  kt->gate_kind = d->gate_kind;
  kt->T = smush ? d->T : 0;
  kt->n_slots = ns;
  kt->vz_only = d->vz_only ? 1 : 0;
  for (int i = 0; i <= d->k; ++i)
    for (int s = 0; s < 6; ++s) {
      const int p = d->p1q[i][s];
      if (p >= d->n_params || p < -1) return SLAM_ERR_INVALID;
      kt->p1q[i][s] = (short)p;
    }
  bool any_bound = false;
  for (int g = 0; g < d->k; ++g) {
    bool bound = false;
    for (int s = 0; s < ns; ++s) {
      const int p = d->slot_param[g][s];
      if (p >= d->n_params || p < -1) return SLAM_ERR_INVALID;
      kt->slot_param[g][s] = (short)p;
      kt->slot_const[g][s] = d->slot_const[g][s];
      if (p >= 0) bound = true;
      else if (!std::isfinite(d->slot_const[g][s])) return SLAM_ERR_INVALID;
    }
    kt->gate_bound[g] = bound ? 1 : 0;
    any_bound |= bound;
  }
  kt->n_trig = 6 * (d->k + 1);
  if (!allow_ties) {  // each Xk entry may be bound to at most one slot (qiskit Parameters are created once per slot, basis.py:136-169);
     // the gradient kernels store, rather than accumulate, each partial derivative
    unsigned char used[SLAM_MAX_PARAMS] = {0};
    for (int i = 0; i <= d->k; ++i)
      for (int s = 0; s < 6; ++s)
        if (kt->p1q[i][s] >= 0 && used[kt->p1q[i][s]]++) return SLAM_ERR_UNSUPPORTED;
    for (int g = 0; g < d->k; ++g)
      for (int s = 0; s < ns; ++s)
        if (kt->slot_param[g][s] >= 0 && used[kt->slot_param[g][s]]++) return SLAM_ERR_UNSUPPORTED;
  }

  if (d->gate_kind == SLAM_GATE_FIXED) {
    kt->gmode = GM_DENSE;
    for (int g = 0; g < d->k; ++g) memcpy(kt->dense[g], d->fixed_gate, sizeof(double) * 32);
    return SLAM_OK;
  }
  if (smush) {
    if (any_bound) {
      if (!allow_bound_smush) return SLAM_ERR_UNSUPPORTED;
      kt->gmode = GM_SMUSH;
      return SLAM_OK;
    }
    kt->gmode = GM_DENSE;  // caller must run lower_const_smush() to fill kt->dense
    return SLAM_OK;
  }
  // RiSwap / ConversionGain: closed form
  bool sym = !any_bound && !kt->vz_only;  // the GM_SYM kernels have the RZ-layer code compiled out
  for (int g = 0; g < d->k; ++g) {
    double* c = kt->gblk[g];
    if (d->gate_kind == SLAM_GATE_RISWAP) {
      const double a = 1.5707963267948966 * kt->slot_const[g][0];
      c[0] = -1.0; c[1] = 0.0;  // phi_c = pi: -i e^{-i pi} = +i  (custom_gates.py:582-595 has +i sin)
      c[2] = 1.0;  c[3] = 0.0;
      c[4] = std::cos(a); c[5] = std::sin(a);
      c[6] = 1.0;  c[7] = 0.0;
      kt->gsym[g][0] = 1.0; kt->gsym[g][1] = 0.0; kt->gsym[g][2] = c[4]; kt->gsym[g][3] = c[5];
    } else {
      const double pc = kt->slot_const[g][0], pg = kt->slot_const[g][1];
      const double ac = kt->slot_const[g][2] * kt->slot_const[g][4], ag = kt->slot_const[g][3] * kt->slot_const[g][4];
      c[0] = std::cos(pc); c[1] = std::sin(pc);
      c[2] = std::cos(pg); c[3] = std::sin(pg);
      c[4] = std::cos(ac); c[5] = std::sin(ac);
      c[6] = std::cos(ag); c[7] = std::sin(ag);
      if (pc != 0.0 || pg != 0.0) sym = false;
      // zero phases: off-diagonals are -i sin(a) (hamiltonian.py:84-111 closed form)
      kt->gsym[g][0] = c[6]; kt->gsym[g][1] = -c[7]; kt->gsym[g][2] = c[4]; kt->gsym[g][3] = -c[5];
    }
  }
  kt->gmode = sym ? GM_SYM : GM_BLOCK;
  if (any_bound) kt->n_trig += 4 * d->k;
  return SLAM_OK;
}

}  // namespace slam

extern "C" {

int slam_abi_version(void) { return SLAM_ABI_VERSION; }

const char* slam_status_string(int status) {
  switch (status) {
    case SLAM_OK: return "ok";
    case SLAM_ERR_INVALID: return "invalid descriptor or argument";
    case SLAM_ERR_UNSUPPORTED: return "unsupported gate kind for this entry point";
    case SLAM_ERR_CUDA: return "CUDA runtime error";
    case SLAM_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

const char* slam_last_cuda_error(void) { return slam::g_cuda_err; }

int slam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int slam_set_device(int device) {
  SLAM_CUDA_CHECK(cudaSetDevice(device));
  return SLAM_OK;
}

}  // extern "C"
