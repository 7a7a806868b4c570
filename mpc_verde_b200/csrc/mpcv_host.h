// mpcv_host.h — host-side handle and error plumbing shared by the ABI translation unit and the
// per-model kernel instantiation units (one .cu object per model, built in parallel).
#pragma once

#include <cuda_runtime.h>

#include <cerrno>
#include <cstdlib>
#include <cstdint>
#include <string>

#include "mpcv_driver.cuh"
#include "mpcv_params.h"

int mpcv_set_error(int code, const std::string& msg);

#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return mpcv_set_error(-EIO, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)

// Per-problem arrays of a host-pointer solve (mpcv_solve_host).  The phase pipeline copies them share by share on
// the pipes' own streams, so that the transfers of one share run under the solve of another.
struct mpcv_host_xfer {
  struct Arr { const char* host_src; char* host_dst; char* dev; size_t row_bytes; };
  Arr in[2];     // x0, p                                   (host_src -> dev)
  Arr out[7];    // x, f, g, lam_g, lam_x, status, iters    (dev -> host_dst)
};

// tuning overrides from the environment (they exist for the A/B measurements quoted in DESIGN.md and for the tests);
// read ONCE, when a handle is created, never on the launch path.  -1 = the compiled-in default.
struct mpcv_knobs {
  int tail_cap = -1, tail_shift = -1, pipes = -1, warp_staged = -1, res_tail = -1;
  int grid_pct = 100;   // fixed grids of the phase kernels as a percentage of "fills the GPU once"
  long pipe_min = -1, resident_below = -1;
  bool hostloop = false;
};
inline mpcv_knobs mpcv_knobs_from_env() {
  mpcv_knobs k;
  if (const char* env = getenv("MPCV_TAIL_BELOW")) k.tail_cap = atoi(env);
  if (const char* env = getenv("MPCV_TAIL_SHIFT")) k.tail_shift = atoi(env);
  if (const char* env = getenv("MPCV_PHASE_PIPES")) { const int v = atoi(env); if (v >= 1) k.pipes = v; }
  if (const char* env = getenv("MPCV_PHASE_PIPE_MIN")) { const long v = atol(env); if (v >= 32) k.pipe_min = v; }
  if (const char* env = getenv("MPCV_PHASE_HOSTLOOP")) k.hostloop = env[0] == '1';
  if (const char* env = getenv("MPCV_WARP_STAGED")) k.warp_staged = atoi(env) & 3;
  if (const char* env = getenv("MPCV_GRID_PCT")) { const int v = atoi(env); if (v >= 10 && v <= 400) k.grid_pct = v; }
  if (const char* env = getenv("MPCV_RES_TAIL")) k.res_tail = atoi(env);                 // 0: ph_tail_kernel
  if (const char* env = getenv("MPCV_RESIDENT_BELOW")) k.resident_below = atol(env);     // AUTO: resident below this batch
  return k;
}

struct mpcv_handle {
  mpcv_spec spec;
  mpcv::Params P;
  mpcv::Layout L;
  int nx, nu, n_var, n_g, n_p, npg, nps;
  bool single;
  int layout;            // resolved MPCV_LAYOUT_* (AUTO for multiple shooting: PHASED here, decided per call by batch size)
  bool layout_auto = false;
  int device;
  int sm_count;
  size_t max_smem_optin;
  size_t smem_per_sm;
  mpcv_knobs knobs;
  bool has_uprev = false;             // the model carries u_prev in its state (move blocking needs it)
  cudaEvent_t last_done = nullptr;    // end of the handle's previous call: the next one waits for it on the device
  double* slab = nullptr;   // thread-layout workspace
  size_t slab_doubles = 0;
  long slab_stride = 0;
  // host-variant staging
  void* hpin = nullptr; size_t hpin_bytes = 0;
  void* dstage = nullptr; size_t dstage_bytes = 0;
  cudaStream_t own_stream = nullptr;
  long long* latency_ns = nullptr;   // optional device buffer [B]
  int64_t launches = 0;
  struct mpcv_phase_state* phase = nullptr;   // phase-kernel pipeline resources (mpcv_phase_inst.cu)
  int64_t phase_graph_launches = 0;
  const mpcv_host_xfer* host_xfer = nullptr;   // set by mpcv_solve_host around its mpcv_solve call
  bool host_xfer_done = false;                 // the launcher took care of the per-problem copies
};

// per-model entry points (defined once per model in mpcv_inst.cu, -DMPCV_INST_MODEL=<id>)
struct mpcv_model_vtable {
  int (*dims)(const mpcv_spec*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*);
  void (*fill)(mpcv_handle*);
  int (*solve)(mpcv_handle*, const mpcv::SolveIO&, long, cudaStream_t);
  int (*loop)(mpcv_handle*, const mpcv::LoopIO&, long, cudaStream_t);
  int (*rollout)(mpcv_handle*, const double*, const double*, double*, double*, long, cudaStream_t);
  int (*derivs)(mpcv_handle*, const double*, const double*, const double*, double*, double*, double*, double*,
                double*, double*, long, cudaStream_t);
};
// phase-kernel pipeline of one model (mpcv_phase_inst.cu, -DMPCV_INST_MODEL=<id>)
struct mpcv_phase_vtable {
  int (*solve)(mpcv_handle*, const mpcv::SolveIO&, long, cudaStream_t);
  void (*release)(struct mpcv_phase_state*);
  int (*sweeps)(mpcv_handle*, cudaStream_t, int*, int*);
  int (*loop)(mpcv_handle*, const mpcv::LoopIO&, long, cudaStream_t);
  int (*res_slots_per_sm)(const mpcv_handle*);   // problems an SM holds in the resident layout (0: does not fit)
};
const mpcv_phase_vtable* mpcv_phase_vtable_of(int model);
#define MPCV_DECLARE_MODEL(id) \
  extern const mpcv_model_vtable mpcv_model_vtable_##id; \
  extern const mpcv_phase_vtable mpcv_phase_vtable_##id;
MPCV_DECLARE_MODEL(0) MPCV_DECLARE_MODEL(1) MPCV_DECLARE_MODEL(2) MPCV_DECLARE_MODEL(3)
MPCV_DECLARE_MODEL(4) MPCV_DECLARE_MODEL(5) MPCV_DECLARE_MODEL(6) MPCV_DECLARE_MODEL(7)
