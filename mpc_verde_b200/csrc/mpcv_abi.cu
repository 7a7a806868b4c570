// mpcv_abi.cu — the C ABI of include/mpcv.h (model-independent part).
// There is NO CPU fallback: every compute entry point fails with -ENODEV when no CUDA device is
// usable; the per-model kernels live in mpcv_inst.cu.
#include <cstdio>
#include <cstring>

#include "mpcv_host.h"
#include "mpcv_c2d.cuh"

using namespace mpcv;

static thread_local std::string g_last_error;
int mpcv_set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static const mpcv_model_vtable* vtable_of(int model) {
  switch (model) {
    case 0: return &mpcv_model_vtable_0;
    case 1: return &mpcv_model_vtable_1;
    case 2: return &mpcv_model_vtable_2;
    case 3: return &mpcv_model_vtable_3;
    case 4: return &mpcv_model_vtable_4;
    case 5: return &mpcv_model_vtable_5;
    case 6: return &mpcv_model_vtable_6;
    case 7: return &mpcv_model_vtable_7;
    default: return nullptr;
  }
}

const mpcv_phase_vtable* mpcv_phase_vtable_of(int model) {
  switch (model) {
    case 0: return &mpcv_phase_vtable_0;
    case 1: return &mpcv_phase_vtable_1;
    case 2: return &mpcv_phase_vtable_2;
    case 3: return &mpcv_phase_vtable_3;
    case 4: return &mpcv_phase_vtable_4;
    case 5: return &mpcv_phase_vtable_5;
    case 6: return &mpcv_phase_vtable_6;
    case 7: return &mpcv_phase_vtable_7;
    default: return nullptr;
  }
}

// FP64 FMA peak: 8 independent register-resident DFMA chains per thread
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}


template <int S>
__global__ void c2d_kernel(int n, int nu, double dt, const double* Ac, const double* Bc, double* A, double* Bd, long B) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double M[S * S], E[S * S];
#pragma unroll
  for (int i = 0; i < S * S; ++i) M[i] = 0.0;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M[i * S + j] = Ac[(b * n + i) * n + j] * dt;
    for (int j = 0; j < nu; ++j) M[i * S + n + j] = Bc[(b * n + i) * nu + j] * dt;
  }
  mpcv::expm_small<S>(M, E);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) A[(b * n + i) * n + j] = E[i * S + j];
    for (int j = 0; j < nu; ++j) Bd[(b * n + i) * nu + j] = E[i * S + n + j];
  }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int mpcv_c2d(int32_t n, int32_t nu, double dt, const double* Ac, const double* Bc, double* A, double* Bd, int64_t B,
             void* stream) {
  if (!Ac || !Bc || !A || !Bd) return mpcv_set_error(-EINVAL, "mpcv_c2d: null argument");
  if (n < 1 || nu < 1 || n + nu > 6) return mpcv_set_error(-EINVAL, "mpcv_c2d: n + nu must be in 2..6");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mpcv_set_error(-ENODEV, "no CUDA device");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((B + 127) / 128);
  switch (n + nu) {
    case 2: c2d_kernel<2><<<grid, 128, 0, st>>>(n, nu, dt, Ac, Bc, A, Bd, (long)B); break;
    case 3: c2d_kernel<3><<<grid, 128, 0, st>>>(n, nu, dt, Ac, Bc, A, Bd, (long)B); break;
    case 4: c2d_kernel<4><<<grid, 128, 0, st>>>(n, nu, dt, Ac, Bc, A, Bd, (long)B); break;
    case 5: c2d_kernel<5><<<grid, 128, 0, st>>>(n, nu, dt, Ac, Bc, A, Bd, (long)B); break;
    default: c2d_kernel<6><<<grid, 128, 0, st>>>(n, nu, dt, Ac, Bc, A, Bd, (long)B); break;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}


const char* mpcv_last_error(void) { return g_last_error.c_str(); }

void mpcv_spec_defaults(mpcv_spec* s) {
  std::memset(s, 0, sizeof(*s));
  s->model = MPCV_MODEL_UNICYCLE_RK4_QUAD;   // multiple_shooting_casadi.py:29-45,75-82
  s->shooting = MPCV_SHOOTING_MULTIPLE;
  s->N = 10; s->M = 4; s->T = 0.2;
  s->Q[0] = 1.0; s->Q[1] = 5.0; s->Q[2] = 0.1;
  s->R[0] = 0.5; s->R[1] = 0.05;
  s->tol = 1e-8; s->max_iter = 3000; s->max_soc = 4; s->mu_init = 0.1;
  s->bound_push = 1e-2; s->bound_frac = 1e-2; s->bound_relax_factor = 1e-8;
  s->nlp_scaling_max_gradient = 100.0;
  s->dual_inf_tol = 1.0; s->constr_viol_tol = 1e-4; s->compl_inf_tol = 1e-4;
  s->acceptable_tol = 1e-6; s->acceptable_iter = 15; s->acceptable_obj_change_tol = 1e20;
}

int mpcv_dims(const mpcv_spec* s, int32_t* nx, int32_t* nu, int32_t* n_var, int32_t* n_g, int32_t* n_p,
              int32_t* npg, int32_t* nps) {
  if (!s) return mpcv_set_error(-EINVAL, "null spec");
  const mpcv_model_vtable* vt = vtable_of(s->model);
  if (!vt) return mpcv_set_error(-EINVAL, "unsupported model id");
  return vt->dims(s, nx, nu, n_var, n_g, n_p, npg, nps);
}

static int create_impl(mpcv_handle* h) {
  const mpcv_model_vtable* vt = vtable_of(h->spec.model);
  if (!vt) return mpcv_set_error(-EINVAL, "unsupported model id");
  vt->fill(h);
  return 0;
}

mpcv_handle* mpcv_create(const mpcv_spec* s) {
  if (!s || s->N < 1 || s->N > 4096) { mpcv_set_error(-EINVAL, "bad spec"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    mpcv_set_error(-ENODEV, "no CUDA device: mpc_verde_b200 has no CPU path");
    return nullptr;
  }
  mpcv_handle* h = new mpcv_handle();
  h->spec = *s;
  h->single = s->shooting == MPCV_SHOOTING_SINGLE;
  h->P = params_from_spec(*s);
  if (create_impl(h) != 0) { delete h; return nullptr; }
  if (s->ntu > 0 && !h->has_uprev) {
    // move blocking pins u_k = u_{k-1}: only the models that carry u_prev in their state can do that
    mpcv_set_error(-EINVAL, "ntu > 0 needs a model with u_prev in its state (MPCV_MODEL_LINEAR3_DU / LINEAR4_DU / FRENET_BICYCLE)");
    delete h;
    return nullptr;
  }
  h->knobs = mpcv_knobs_from_env();
  cudaGetDevice(&h->device);
  // Keep the context's local-memory pool at its high-water mark.  The one-kernel layouts carry 0.7 - 1.2 KB stack
  // frames per thread; without this flag the driver shrinks and regrows the pool around launches of kernels with
  // smaller frames once such a kernel has run, and a closed loop (hundreds of short launches) that follows a
  // single-shooting solve in the same process ran 50 % slower (C4 after C1: 153 ms against 100 ms).
  {
    unsigned flags = 0;
    if (cudaGetDeviceFlags(&flags) == cudaSuccess && !(flags & cudaDeviceLmemResizeToMax))
      cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax);
    cudaGetLastError();
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { mpcv_set_error(-EIO, "cudaGetDeviceProperties"); delete h; return nullptr; }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = prop.sharedMemPerBlockOptin;
  h->smem_per_sm = prop.sharedMemPerMultiprocessor;
  h->layout = s->layout;
  // AUTO: one thread per problem for single shooting; multiple shooting: the CTA-resident kernel for batches below
  // the crossover (it is latency-bound: 26 problems per SM), the slab pipeline above it (profiles/r2b_*).
  const int res_slots = mpcv_phase_vtable_of(s->model)->res_slots_per_sm(h);
  if (h->layout == MPCV_LAYOUT_AUTO) {
    h->layout = h->single ? MPCV_LAYOUT_THREAD : MPCV_LAYOUT_PHASED;
    h->layout_auto = !h->single;       // per call: the resident kernel below the crossover batch, the pipeline above
  }
  if (h->layout == MPCV_LAYOUT_RESIDENT && res_slots < 1) h->layout = MPCV_LAYOUT_PHASED;
  if (h->single) h->layout = MPCV_LAYOUT_THREAD;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming) != cudaSuccess) {
    mpcv_set_error(-EIO, "cudaStreamCreate"); delete h; return nullptr;
  }
  // diagnostic counters the kernels bump (Params::diag): [0] filter overflows
  if (cudaMalloc(&h->P.diag, 4 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(h->P.diag, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
    mpcv_set_error(-EIO, "cudaMalloc (diagnostic counters)"); mpcv_destroy(h); return nullptr;
  }
  return h;
}

void mpcv_destroy(mpcv_handle* h) {
  if (!h) return;
  if (h->phase) mpcv_phase_vtable_of(h->spec.model)->release(h->phase);
  if (h->slab) cudaFree(h->slab);
  if (h->hpin) cudaFreeHost(h->hpin);
  if (h->dstage) cudaFree(h->dstage);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->last_done) cudaEventDestroy(h->last_done);
  if (h->P.diag) cudaFree(h->P.diag);
  delete h;
}

int mpcv_set_knob(mpcv_handle* h, const char* name, int64_t value) {
  if (!h || !name) return mpcv_set_error(-EINVAL, "mpcv_set_knob: null argument");
  if (h->phase) return mpcv_set_error(-EBUSY, "mpcv_set_knob: the handle has laid out its pipes already (set knobs before the first solve)");
  const std::string n(name);
  if (n == "phase_pipes") { if (value < 1 || value > 8) return mpcv_set_error(-EINVAL, "phase_pipes: 1..8"); h->knobs.pipes = (int)value; }
  else if (n == "phase_pipe_min") { if (value < 32) return mpcv_set_error(-EINVAL, "phase_pipe_min: >= 32"); h->knobs.pipe_min = (long)value; }
  else if (n == "tail_below") h->knobs.tail_cap = (int)value;
  else if (n == "tail_shift") h->knobs.tail_shift = (int)value;
  else if (n == "resident_below") h->knobs.resident_below = (long)value;
  else return mpcv_set_error(-EINVAL, "mpcv_set_knob: unknown knob " + n);
  return 0;
}

int mpcv_diag(mpcv_handle* h, uint64_t* counters4) {
  if (!h || !counters4) return mpcv_set_error(-EINVAL, "mpcv_diag: null argument");
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(counters4, h->P.diag, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return mpcv_set_error(-EIO, "mpcv_diag: copy failed");
  return 0;
}

int64_t mpcv_launch_count(const mpcv_handle* h) { return h ? h->launches : 0; }

int mpcv_phase_sweeps(mpcv_handle* h, int32_t* sweeps, int64_t* kernel_nodes, void* stream) {
  if (!h) return mpcv_set_error(-EINVAL, "mpcv_phase_sweeps: null argument");
  int n = 0, cum = 0;
  if (int rc = mpcv_phase_vtable_of(h->spec.model)->sweeps(h, (cudaStream_t)stream, &n, &cum)) return rc;
  if (sweeps) *sweeps = n;
  // launches issued from the host + 12 kernel nodes per graph-driven sweep
  if (kernel_nodes) *kernel_nodes = h->launches + (h->phase_graph_launches > 0 ? 12 * (int64_t)cum : 0);
  return 0;
}

int mpcv_set_latency_buffer(mpcv_handle* h, long long* dev_ns) {
  if (!h) return mpcv_set_error(-EINVAL, "null handle");
  h->latency_ns = dev_ns;
  return 0;
}

// Every call reuses device state owned by the handle (workspace slabs, lists, control blocks, graphs, staging):
// a call first makes its stream wait for the end of the handle's previous call, and records its own end.
// The handle belongs to the device it was created on.
static int call_begin(mpcv_handle* h, cudaStream_t st, const char* what) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != h->device)
    return mpcv_set_error(-EINVAL, std::string(what) + ": the handle was created on device " + std::to_string(h->device) +
                                       ", the current device is " + std::to_string(dev));
  CUDA_OK(cudaStreamWaitEvent(st, h->last_done, 0));
  return 0;
}
static int call_end(mpcv_handle* h, cudaStream_t st, int rc) {
  // (recorded on failure too: whatever was enqueued before the error still uses the handle's buffers)
  if (cudaEventRecord(h->last_done, st) != cudaSuccess && rc == 0) return mpcv_set_error(-EIO, "cudaEventRecord");
  return rc;
}

static int solve_impl(mpcv_handle* h, const SolveIO& io, int64_t B, cudaStream_t st, const char* what) {
  if (int rc = call_begin(h, st, what)) return rc;
  return call_end(h, st, vtable_of(h->spec.model)->solve(h, io, (long)B, st));
}

int mpcv_solve(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx, const double* p,
               double* x, double* f, double* g, double* lam_g, double* lam_x, int32_t* status, int32_t* iters,
               int64_t B, void* stream) {
  if (!h || !lbx || !ubx || !p) return mpcv_set_error(-EINVAL, "mpcv_solve: null argument");
  SolveIO io{x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, h->latency_ns};
  return solve_impl(h, io, B, (cudaStream_t)stream, "mpcv_solve");
}

int mpcv_solve_bounds(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx, const double* p,
                      double* x, double* f, double* g, double* lam_g, double* lam_x, int32_t* status,
                      int32_t* iters, int64_t B, void* stream) {
  if (!h || !lbx || !ubx || !p) return mpcv_set_error(-EINVAL, "mpcv_solve_bounds: null argument");
  SolveIO io{x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, h->latency_ns};
  io.bstride = h->n_var;
  return solve_impl(h, io, B, (cudaStream_t)stream, "mpcv_solve_bounds");
}

int mpcv_rollout(mpcv_handle* h, const double* p, const double* U, double* X, double* q, int64_t B, void* stream) {
  if (!h || !p || !U || !X) return mpcv_set_error(-EINVAL, "mpcv_rollout: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = call_begin(h, st, "mpcv_rollout")) return rc;
  return call_end(h, st, vtable_of(h->spec.model)->rollout(h, p, U, X, q, (long)B, st));
}

int mpcv_stage_derivs(mpcv_handle* h, const double* z, const double* pstage, const double* lam, double* xn,
                      double* A, double* Bm, double* q, double* grad, double* H, int64_t B, void* stream) {
  if (!h || !z || !lam) return mpcv_set_error(-EINVAL, "mpcv_stage_derivs: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = call_begin(h, st, "mpcv_stage_derivs")) return rc;
  return call_end(h, st, vtable_of(h->spec.model)->derivs(h, z, pstage, lam, xn, A, Bm, q, grad, H, (long)B, st));
}

int mpcv_closed_loop_ex(mpcv_handle* h, const mpcv_loop_args* a, int64_t B, void* stream) {
  if (!h || !a || !a->x_init || !a->lbx || !a->ubx || !a->out_states || !a->out_controls)
    return mpcv_set_error(-EINVAL, "mpcv_closed_loop: null argument");
  if (h->nps > 0 && !a->ptraj) return mpcv_set_error(-EINVAL, "mpcv_closed_loop: ptraj required for this model");
  if (h->npg > 0 && !a->pglob && !a->pglob_traj)
    return mpcv_set_error(-EINVAL, "mpcv_closed_loop: pglob or pglob_traj required for this model");
  if (a->n_steps < 0) return mpcv_set_error(-EINVAL, "mpcv_closed_loop: n_steps < 0");
  LoopIO io{a->x_init, a->pglob, a->ptraj, a->lbx, a->ubx, a->out_states, a->out_controls, a->out_steps, a->out_iters,
            a->out_status, a->n_steps, a->warm_mode, a->stop_radius};
  io.pglob_traj = a->pglob_traj;
  io.out_horizons = a->out_horizons;
  io.out_step_ns = a->out_step_ns;
  io.flags = a->flags;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = call_begin(h, st, "mpcv_closed_loop")) return rc;
  return call_end(h, st, vtable_of(h->spec.model)->loop(h, io, (long)B, st));
}

int mpcv_closed_loop(mpcv_handle* h, const double* x_init, const double* pglob, const double* ptraj,
                     const double* lbx, const double* ubx, int32_t n_steps, int32_t warm_mode, double stop_radius,
                     double* out_states, double* out_controls, int32_t* out_steps, int32_t* out_iters,
                     int32_t* out_status, int64_t B, void* stream) {
  mpcv_loop_args a = {};
  a.x_init = x_init; a.pglob = pglob; a.ptraj = ptraj; a.lbx = lbx; a.ubx = ubx;
  a.n_steps = n_steps; a.warm_mode = warm_mode; a.stop_radius = stop_radius;
  a.out_states = out_states; a.out_controls = out_controls; a.out_steps = out_steps; a.out_iters = out_iters;
  a.out_status = out_status;
  return mpcv_closed_loop_ex(h, &a, B, stream);
}

// ---- host-pointer variant: H2D from pinned staging, solve, D2H, synchronise ---------------
static int ensure_staging(mpcv_handle* h, size_t bytes) {
  if (bytes > h->hpin_bytes) {
    if (h->hpin) cudaFreeHost(h->hpin);
    h->hpin = nullptr; h->hpin_bytes = 0;
    CUDA_OK(cudaMallocHost(&h->hpin, bytes));
    h->hpin_bytes = bytes;
  }
  if (bytes > h->dstage_bytes) {
    if (h->dstage) cudaFree(h->dstage);
    h->dstage = nullptr; h->dstage_bytes = 0;
    CUDA_OK(cudaMalloc(&h->dstage, bytes));
    h->dstage_bytes = bytes;
  }
  return 0;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) can be the source or
// the destination of an asynchronous copy directly; pageable memory goes through the pinned staging block
static bool is_pinned_host(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int mpcv_solve_host(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx, const double* p,
                    double* x, double* f, double* g, double* lam_g, double* lam_x, int32_t* status,
                    int32_t* iters, int64_t B) {
  if (!h || !lbx || !ubx || !p) return mpcv_set_error(-EINVAL, "mpcv_solve_host: null argument");
  const size_t n = h->n_var, ng = h->n_g, np = h->n_p;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  // input block: x0 | lbx | ubx | p ; output block: x | f | g | lam_g | lam_x | status | iters
  const size_t o_x0 = 0, o_lb = o_x0 + al(x0 ? B * n * 8 : 0), o_ub = o_lb + al(n * 8), o_p = o_ub + al(n * 8);
  const size_t in_bytes = o_p + al(B * np * 8);
  const size_t o_x = in_bytes, o_f = o_x + al(B * n * 8), o_g = o_f + al(B * 8), o_lg = o_g + al(g ? B * ng * 8 : 0),
               o_lx = o_lg + al(lam_g ? B * ng * 8 : 0), o_st = o_lx + al(lam_x ? B * n * 8 : 0), o_it = o_st + al(B * 4);
  const size_t total = o_it + al(B * 4);
  if (int rc = ensure_staging(h, total)) return rc;
  char* hp = (char*)h->hpin;
  char* dp = (char*)h->dstage;
  cudaStream_t st = h->own_stream;
  struct Xfer { const void* user; size_t off, bytes, row; };
  const Xfer ins[] = {{x0, o_x0, x0 ? (size_t)B * n * 8 : 0, n * 8}, {p, o_p, (size_t)B * np * 8, np * 8},
                      {lbx, o_lb, n * 8, 0}, {ubx, o_ub, n * 8, 0}};
  const Xfer outs[] = {{x, o_x, (size_t)B * n * 8, n * 8}, {f, o_f, (size_t)B * 8, 8}, {g, o_g, (size_t)B * ng * 8, ng * 8},
                       {lam_g, o_lg, (size_t)B * ng * 8, ng * 8}, {lam_x, o_lx, (size_t)B * n * 8, n * 8},
                       {status, o_st, (size_t)B * 4, 4}, {iters, o_it, (size_t)B * 4, 4}};
  // pageable user memory goes through the pinned staging block (inputs now, outputs after the synchronise)
  const char* in_src[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 4; ++i) {
    const Xfer& t = ins[i];
    if (!t.user || t.bytes == 0) continue;
    in_src[i] = (const char*)t.user;
    if (!is_pinned_host(t.user)) { std::memcpy(hp + t.off, t.user, t.bytes); in_src[i] = hp + t.off; }
  }
  char* out_dst[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool staged[7] = {false, false, false, false, false, false, false};
  for (int i = 0; i < 7; ++i) {
    const Xfer& t = outs[i];
    if (!t.user || t.bytes == 0) continue;
    out_dst[i] = (char*)const_cast<void*>(t.user);
    if (!is_pinned_host(t.user)) { out_dst[i] = hp + t.off; staged[i] = true; }
  }
  // the bounds are shared by the batch: copied up front.  The per-problem arrays are handed to the phase pipeline,
  // which copies each pipe's share on that pipe's stream (H2D of one share under the solve of another, D2H of a
  // finished share under the stragglers of the rest); any other layout copies them here, around the solve.
  for (int i = 2; i < 4; ++i)
    CUDA_OK(cudaMemcpyAsync(dp + ins[i].off, in_src[i], ins[i].bytes, cudaMemcpyHostToDevice, st));
  mpcv_host_xfer xf = {};
  for (int i = 0; i < 2; ++i) if (in_src[i]) xf.in[i] = {in_src[i], nullptr, dp + ins[i].off, ins[i].row};
  for (int i = 0; i < 7; ++i) if (out_dst[i]) xf.out[i] = {nullptr, out_dst[i], dp + outs[i].off, outs[i].row};
  const bool pipelined = (h->layout == MPCV_LAYOUT_PHASED || h->layout == MPCV_LAYOUT_RESIDENT) && !h->single && B > 0;
  if (!pipelined)
    for (int i = 0; i < 2; ++i)
      if (in_src[i]) CUDA_OK(cudaMemcpyAsync(dp + ins[i].off, in_src[i], ins[i].bytes, cudaMemcpyHostToDevice, st));
  h->host_xfer = pipelined ? &xf : nullptr;
  h->host_xfer_done = false;
  int rc = mpcv_solve(h, x0 ? (const double*)(dp + o_x0) : nullptr, (const double*)(dp + o_lb), (const double*)(dp + o_ub),
                      (const double*)(dp + o_p), (double*)(dp + o_x), (double*)(dp + o_f), g ? (double*)(dp + o_g) : nullptr,
                      lam_g ? (double*)(dp + o_lg) : nullptr, lam_x ? (double*)(dp + o_lx) : nullptr,
                      (int32_t*)(dp + o_st), (int32_t*)(dp + o_it), B, st);
  h->host_xfer = nullptr;
  if (rc) return rc;
  if (pipelined && !h->host_xfer_done) return mpcv_set_error(-EIO, "mpcv_solve_host: the phase pipeline did not take the host transfers");
  if (!pipelined)
    for (int i = 0; i < 7; ++i)
      if (out_dst[i]) CUDA_OK(cudaMemcpyAsync(out_dst[i], dp + outs[i].off, outs[i].bytes, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  for (int i = 0; i < 7; ++i)
    if (staged[i]) std::memcpy(const_cast<void*>(outs[i].user), hp + outs[i].off, outs[i].bytes);
  return 0;
}

int mpcv_fp64_peak(double* tflops, double* ms, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mpcv_set_error(-ENODEV, "no CUDA device");
  cudaStream_t st = (cudaStream_t)stream;
  cudaDeviceProp prop;
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  double* out = nullptr;
  CUDA_OK(cudaMalloc(&out, (size_t)threads * blocks * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_OK(cudaEventRecord(e0, st));
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(out, iters);
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float t;
    CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
    if (rep > 0 && t < best) best = t;
  }
  CUDA_OK(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  const double flops = 2.0 * 8.0 * iters * (double)threads * blocks;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return 0;
}

}  // extern "C"
