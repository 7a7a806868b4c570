// mpcv_phase_inst.cu — the phase-kernel pipeline of ONE model (-DMPCV_INST_MODEL=<id>) and its
// launcher.  Compiled as its own translation unit with the phase functions force-inlined
// (MPCV_INLINE_PHASES): every phase kernel is a straight-line program that keeps its operands in
// registers, whereas the one-kernel solve of mpcv_inst.cu keeps the phases as real calls to bound
// its instruction footprint.
#define MPCV_INLINE_PHASES 1
#include "mpcv_host.h"
#include "mpcv_phase.cuh"

using namespace mpcv;

#ifndef MPCV_INST_MODEL
#error "compile with -DMPCV_INST_MODEL=<model id>"
#endif
#include "mpcv_model_select.h"

struct mpcv_phase_state {
  int* act[2] = {nullptr, nullptr};
  int* retry = nullptr;
  int* slow = nullptr;
  double* slab2 = nullptr;          // second workspace slab (repack target)
  LoopBufs lb = {};                 // closed-loop driver buffers
  long lb_cap = 0;
  size_t slab2_doubles = 0;
  PhaseCtrl* ctrl = nullptr;
  SolveIO* d_io = nullptr;
  PhaseCtrl* h_ctrl = nullptr;      // pinned mirror (host-loop mode)
  long cap = 0;                     // problems the lists / grids are sized for
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool graph_failed = false;
  double* graph_slab = nullptr;     // the graph bakes these in: rebuild when they change
  long graph_stride = 0;
};

static void loop_free(mpcv_phase_state* s) {
  void* ptrs[] = {s->lb.x0, s->lb.p, s->lb.x, s->lb.f, s->lb.state, s->lb.status, s->lb.iters, s->lb.active,
                  s->lb.steps, s->lb.iters_total, s->lb.worst};
  for (void* q : ptrs) if (q) cudaFree(q);
  s->lb = LoopBufs{};
  s->lb_cap = 0;
}

static void phase_free(mpcv_phase_state* s) {
  if (!s) return;
  if (s->exec) cudaGraphExecDestroy(s->exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  if (s->act[0]) cudaFree(s->act[0]);
  if (s->act[1]) cudaFree(s->act[1]);
  if (s->retry) cudaFree(s->retry);
  if (s->slow) cudaFree(s->slow);
  if (s->slab2) cudaFree(s->slab2);
  loop_free(s);
  if (s->ctrl) cudaFree(s->ctrl);
  if (s->d_io) cudaFree(s->d_io);
  if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
  delete s;
}

static int phase_ensure(mpcv_handle* h, long B) {
  if (!h->phase) h->phase = new mpcv_phase_state();
  mpcv_phase_state* s = h->phase;
  if (!s->ctrl) {
    CUDA_OK(cudaMalloc(&s->ctrl, sizeof(PhaseCtrl)));
    // the memset runs on the legacy default stream, which does NOT order against non-blocking streams:
    // wait for it, or it can land after the first ph_begin_kernel and wipe B / n_act
    CUDA_OK(cudaMemset(s->ctrl, 0, sizeof(PhaseCtrl)));
    CUDA_OK(cudaStreamSynchronize(0));
    CUDA_OK(cudaMalloc(&s->d_io, sizeof(SolveIO)));
    CUDA_OK(cudaMallocHost(&s->h_ctrl, sizeof(PhaseCtrl)));
  }
  const long cap = (B + kPhaseThreads - 1) / kPhaseThreads * kPhaseThreads;
  if (cap > s->cap) {
    for (int i = 0; i < 2; ++i) {
      if (s->act[i]) cudaFree(s->act[i]);
      s->act[i] = nullptr;
    }
    if (s->retry) cudaFree(s->retry);
    if (s->slow) cudaFree(s->slow);
    s->retry = s->slow = nullptr;
    s->cap = 0;
    CUDA_OK(cudaMalloc(&s->act[0], cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->act[1], cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->retry, cap * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->slow, cap * sizeof(int)));
    s->cap = cap;
    if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
  }
  // thread-layout slab sized for the capacity (stride = cap)
  const size_t need = (size_t)s->cap * h->L.total;
  if (need > h->slab_doubles) {
    if (h->slab) cudaFree(h->slab);
    h->slab = nullptr;
    h->slab_doubles = 0;
    CUDA_OK(cudaMalloc(&h->slab, need * sizeof(double)));
    h->slab_doubles = need;
  }
  if (need > s->slab2_doubles) {
    if (s->slab2) cudaFree(s->slab2);
    s->slab2 = nullptr;
    s->slab2_doubles = 0;
    CUDA_OK(cudaMalloc(&s->slab2, need * sizeof(double)));
    s->slab2_doubles = need;
    if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
  }
  h->slab_stride = s->cap;
  if (s->exec && (s->graph_slab != h->slab || s->graph_stride != h->slab_stride)) {
    cudaGraphExecDestroy(s->exec); s->exec = nullptr;
    cudaGraphDestroy(s->graph); s->graph = nullptr;
  }
  return 0;
}

template <class Model>
static PhaseArgs phase_args(const mpcv_handle* h) {
  const mpcv_phase_state* s = h->phase;
  PhaseArgs a;
  a.P = h->P; a.L = h->L; a.slab[0] = h->slab; a.slab[1] = s->slab2;
  a.act[0] = s->act[0]; a.act[1] = s->act[1];
  a.retry = s->retry; a.slow = s->slow;
  a.ctrl = s->ctrl; a.io = s->d_io; a.cap = s->cap;
  return a;
}

static size_t phase_smem(const mpcv_handle* h) { return (size_t)h->L.n * sizeof(BndEntry); }

// fixed grids: enough CTAs to fill the GPU once (never more than the work of a full batch needs)
struct PhaseGrids { unsigned prob, stage, warp, group; };
static PhaseGrids phase_grids(const mpcv_handle* h) {
  const long cap = h->phase->cap, sm = h->sm_count;
  const long wpb = kWarpPhaseThreads / 32;
  auto clampu = [](long need, long fill) { return (unsigned)(need < fill ? (need < 1 ? 1 : need) : fill); };
  PhaseGrids g;
  g.prob = clampu(cap / kPhaseThreads, sm * 4);
  g.stage = clampu(cap / kPhaseThreads * h->L.N, sm * 8);
  g.warp = clampu((cap + wpb - 1) / wpb, sm * 8);
  const long gpb = kWarpPhaseThreads / kGroupLanes;
  g.group = clampu((cap + gpb - 1) / gpb, sm * 8);
  return g;
}

template <class K>
static int phase_set_smem(K kern, size_t smem) {
  if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// ---- graph construction: init chain, then a WHILE node whose body is one iteration sweep ----------
static int add_kernel(cudaGraph_t g, cudaGraphNode_t* node, cudaGraphNode_t* dep, void* func, unsigned grid,
                      unsigned block, size_t smem, void** args) {
  cudaKernelNodeParams kp = {};
  kp.func = func;
  kp.gridDim = dim3(grid, 1, 1);
  kp.blockDim = dim3(block, 1, 1);
  kp.sharedMemBytes = (unsigned)smem;
  kp.kernelParams = args;
  kp.extra = nullptr;
  CUDA_OK(cudaGraphAddKernelNode(node, g, dep, dep ? 1 : 0, &kp));
  return 0;
}

template <class Model>
static int phase_build_graph(mpcv_handle* h) {
  mpcv_phase_state* s = h->phase;
  PhaseArgs a = phase_args<Model>(h);
  const PhaseGrids gr = phase_grids(h);
  const size_t smem = phase_smem(h);
  int hess0 = 0, hess1 = 1, use_handle = 1;
  cudaGraph_t g = nullptr;
  CUDA_OK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle handle;
  CUDA_OK(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNode_t n_init, n_der0, n_init2, n_der1, n_while;
  void* a_init[] = {&a};
  void* a_der0[] = {&a, &hess0};
  void* a_der1[] = {&a, &hess1};
  if (int rc = add_kernel(g, &n_init, nullptr, (void*)ph_init_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(g, &n_der0, &n_init, (void*)ph_der0_kernel<Model>, gr.stage, kPhaseThreads, smem, a_der0)) return rc;
  if (int rc = add_kernel(g, &n_init2, &n_der0, (void*)ph_init2_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(g, &n_der1, &n_init2, (void*)ph_der0_kernel<Model>, gr.stage, kPhaseThreads, smem, a_der1)) return rc;
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = handle;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  CUDA_OK(cudaGraphAddNode(&n_while, g, &n_der1, 1, &cp));
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  cudaGraphNode_t b_probe, b_pre, b_repack, b_relist, b_factor, b_retry, b_post, b_trial, b_accept, b_slow, b_der, b_flip;
  void* a_flip[] = {&s->ctrl, &handle, &use_handle};
  if (int rc = add_kernel(body, &b_pre, nullptr, (void*)ph_pre_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_repack, &b_pre, (void*)ph_repack_kernel<Model>, gr.prob, kPhaseThreads, 0, a_init)) return rc;
  if (int rc = add_kernel(body, &b_relist, &b_repack, (void*)ph_repack_list_kernel<Model>, gr.prob, kPhaseThreads, 0, a_init)) return rc;
  if (int rc = add_kernel(body, &b_factor, &b_relist, (void*)ph_factor_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_probe, &b_factor, (void*)ph_probe_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_retry, &b_probe, (void*)ph_retry_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_post, &b_retry, (void*)ph_post_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_trial, &b_post, (void*)ph_trial_kernel<Model>, gr.stage, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_accept, &b_trial, (void*)ph_accept_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_slow, &b_accept, (void*)ph_slow_kernel<Model>, gr.warp, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_der, &b_slow, (void*)ph_der_kernel<Model>, gr.stage, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_flip, &b_der, (void*)ph_flip_kernel, 1, 1, 0, a_flip)) return rc;
  // the stragglers finish in one persistent kernel after the loop
  cudaGraphNode_t n_tail;
  if (int rc = add_kernel(g, &n_tail, &n_while, (void*)ph_tail_kernel<Model>, gr.warp, kWarpPhaseThreads, smem, a_init)) return rc;
  cudaGraphExec_t exec = nullptr;
  CUDA_OK(cudaGraphInstantiate(&exec, g, 0));
  s->graph = g;
  s->exec = exec;
  s->graph_slab = h->slab;
  s->graph_stride = h->slab_stride;
  return 0;
}

// host-driven loop (fallback when the graph cannot be built, and MPCV_PHASE_HOSTLOOP=1): same
// kernels, the host reads the active count back every few sweeps
template <class Model>
static int phase_host_loop(mpcv_handle* h, cudaStream_t st) {
  mpcv_phase_state* s = h->phase;
  const PhaseArgs a = phase_args<Model>(h);
  const PhaseGrids gr = phase_grids(h);
  const size_t smem = phase_smem(h);
  ph_init_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
  ph_der0_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a, 0);
  ph_init2_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
  ph_der0_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a, 1);
  h->launches += 4;
  const cudaGraphConditionalHandle none = 0;
  const int chunk = 4;
  for (long sweep = 0; sweep < (long)h->P.max_iter + 2; sweep += chunk) {
    for (int c = 0; c < chunk; ++c) {
      ph_pre_kernel<Model><<<gr.group, kWarpPhaseThreads, smem, st>>>(a);
      ph_repack_kernel<Model><<<gr.prob, kPhaseThreads, 0, st>>>(a);
      ph_repack_list_kernel<Model><<<gr.prob, kPhaseThreads, 0, st>>>(a);
      ph_factor_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_probe_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_retry_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_post_kernel<Model><<<gr.group, kWarpPhaseThreads, smem, st>>>(a);
      ph_trial_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a);
      ph_accept_kernel<Model><<<gr.group, kWarpPhaseThreads, smem, st>>>(a);
      ph_slow_kernel<Model><<<gr.warp, kWarpPhaseThreads, smem, st>>>(a);
      ph_der_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a);
      ph_flip_kernel<<<1, 1, 0, st>>>(s->ctrl, none, 0);
      h->launches += 12;
    }
    CUDA_OK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(PhaseCtrl), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (s->h_ctrl->n_act[s->h_ctrl->sweep & 1] <= ph_tail_below(s->h_ctrl->B)) break;
  }
  ph_tail_kernel<Model><<<gr.warp, kWarpPhaseThreads, smem, st>>>(a);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class Model>
static int launch_solve_phased(mpcv_handle* h, const SolveIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if (B > 0x7fffffffL) return mpcv_set_error(-EINVAL, "batch too large");
  if (int rc = phase_ensure(h, B)) return rc;
  mpcv_phase_state* s = h->phase;
  const size_t smem = phase_smem(h);
  if (smem > h->max_smem_optin) return mpcv_set_error(-ENOMEM, "bounds table exceeds shared memory; use MPCV_LAYOUT_WARP");
  if (!s->exec && !s->graph_failed) {
    if (phase_set_smem(ph_init_kernel<Model>, smem) || phase_set_smem(ph_der0_kernel<Model>, smem) ||
        phase_set_smem(ph_init2_kernel<Model>, smem) || phase_set_smem(ph_pre_kernel<Model>, smem) ||
        phase_set_smem(ph_factor_kernel<Model>, smem) || phase_set_smem(ph_post_kernel<Model>, smem) ||
        phase_set_smem(ph_trial_kernel<Model>, smem) || phase_set_smem(ph_accept_kernel<Model>, smem) ||
        phase_set_smem(ph_retry_kernel<Model>, smem) || phase_set_smem(ph_probe_kernel<Model>, smem) || phase_set_smem(ph_slow_kernel<Model>, smem) ||
        phase_set_smem(ph_tail_kernel<Model>, smem) ||
        phase_set_smem(ph_der_kernel<Model>, smem))
      return -EIO;
  }
  ph_begin_kernel<<<1, 1, 0, st>>>(s->ctrl, s->d_io, io, (int)B);
  h->launches++;
  const char* env = getenv("MPCV_PHASE_HOSTLOOP");
  const bool want_graph = !(env && env[0] == '1') && !s->graph_failed;
  if (want_graph && !s->exec) {
    if (phase_build_graph<Model>(h) != 0) {
      // keep going with the host-driven loop; remember why
      s->graph_failed = true;
      cudaGetLastError();
      if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
    }
  }
  if (want_graph && s->exec) {
    CUDA_OK(cudaGraphLaunch(s->exec, st));
    h->launches += 5;    // init chain + tail; the sweeps are counted from the device (mpcv_phase_sweeps)
    h->phase_graph_launches++;
    return 0;
  }
  return phase_host_loop<Model>(h, st);
}

// closed loop: per MPC step  prepare -> solve (graph) -> apply, all stream-ordered
template <class Model>
static int launch_loop_phased(mpcv_handle* h, const LoopIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if (int rc = phase_ensure(h, B)) return rc;
  mpcv_phase_state* s = h->phase;
  if (B > s->lb_cap) {
    loop_free(s);
    const size_t n = h->n_var, np = h->n_p, nx = h->nx;
    CUDA_OK(cudaMalloc(&s->lb.x0, B * n * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.p, B * np * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.x, B * n * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.f, B * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.state, B * nx * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.status, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.iters, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.active, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.steps, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.iters_total, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.worst, B * sizeof(int)));
    s->lb_cap = B;
  }
  const LoopBufs lb = s->lb;
  const unsigned grid = (unsigned)((B + 127) / 128);
  lp_begin_kernel<Model><<<grid, 128, 0, st>>>(h->L, io, lb, B);
  h->launches++;
  long long* const lat = h->latency_ns;
  h->latency_ns = nullptr;
  const SolveIO sio{lb.x0, io.lbx, io.ubx, lb.p, lb.x, lb.f, nullptr, nullptr, nullptr, lb.status, lb.iters, nullptr};
  int rc = 0;
  for (int t = 0; t < io.n_steps && rc == 0; ++t) {
    lp_prepare_kernel<Model><<<grid, 128, 0, st>>>(h->L, io, lb, B, t);
    rc = launch_solve_phased<Model>(h, sio, B, st);
    lp_apply_kernel<Model><<<grid, 128, 0, st>>>(h->P, h->L, io, lb, B, t);
    h->launches += 2;
  }
  h->latency_ns = lat;
  if (rc) return rc;
  lp_end_kernel<Model><<<grid, 128, 0, st>>>(io, lb, B);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int phase_sweeps(mpcv_handle* h, cudaStream_t st, int* sweeps, int* cumulative) {
  if (!h->phase || !h->phase->ctrl) { *sweeps = 0; *cumulative = 0; return 0; }
  CUDA_OK(cudaMemcpyAsync(h->phase->h_ctrl, h->phase->ctrl, sizeof(PhaseCtrl), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  *sweeps = h->phase->h_ctrl->sweeps_total;
  *cumulative = h->phase->h_ctrl->sweeps_cum;
  return 0;
}

#define MPCV_CAT2(a, b) a##b
#define MPCV_CAT(a, b) MPCV_CAT2(a, b)
extern const mpcv_phase_vtable MPCV_CAT(mpcv_phase_vtable_, MPCV_INST_MODEL) = {launch_solve_phased<ModelT>, phase_free,
                                                                                phase_sweeps, launch_loop_phased<ModelT>};
