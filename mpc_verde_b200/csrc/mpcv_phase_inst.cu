// mpcv_phase_inst.cu — the phase-kernel pipeline of ONE model (-DMPCV_INST_MODEL=<id>) and its
// launcher.  Compiled as its own translation unit with the phase functions force-inlined
// (MPCV_INLINE_PHASES): every phase kernel is a straight-line program that keeps its operands in
// registers, whereas the one-kernel solve of mpcv_inst.cu keeps the phases as real calls to bound
// its instruction footprint.
#define MPCV_INLINE_PHASES 1
#include "mpcv_host.h"
#include "mpcv_phase.cuh"
#include "mpcv_resident.cuh"

using namespace mpcv;

#ifndef MPCV_INST_MODEL
#error "compile with -DMPCV_INST_MODEL=<model id>"
#endif
#include "mpcv_model_select.h"

// One pipe = one independent instance of the pipeline (lists, control block, graph) over a contiguous share of
// the batch, launched on its own stream.  The sweeps are batch-synchronous, so a single pipe leaves the GPU
// waiting at every launch boundary and, late in the solve, at the latency floor of each phase; with several
// pipes in flight a memory-bound phase of one share overlaps the FP64-bound derivative sweep or a floor-bound
// launch of another.  MPCV_PHASE_PIPES (1..kMaxPipes) overrides the choice.
constexpr int kMaxPipes = 8;
#ifndef MPCV_PIPES_DEFAULT
#define MPCV_PIPES_DEFAULT 4   /* C2, B = 65,536 on B200: 1 pipe 20.4 ms, 2: 19.3, 3: 18.8, 4: 18.7, 8: 25.1 */
#endif
#ifndef MPCV_PIPE_MIN
#define MPCV_PIPE_MIN 16384    /* do not split below this many problems per pipe (8 x 8,192 is slower than one pipe) */
#endif
struct PhasePipe {
  int* act[2] = {nullptr, nullptr};
  int* retry = nullptr;
  int* slow = nullptr;
  double* slab[2] = {nullptr, nullptr};
  PhaseCtrl* ctrl = nullptr;
  SolveIO* d_io = nullptr;
  PhaseCtrl* h_ctrl = nullptr;      // pinned mirror (host-loop mode, sweep accounting)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t stream = nullptr;    // pipes 1.. run on their own stream, forked from / joined to the caller's
  cudaEvent_t done = nullptr;
};
struct mpcv_phase_state {
  PhasePipe pipe[kMaxPipes];
  int npipes = 0;                   // pipes the lists / graphs are laid out for
  int* lists = nullptr;             // one allocation: 4 lists x npipes x cap
  PhaseCtrl* ctrl = nullptr;        // [kMaxPipes]
  SolveIO* d_io = nullptr;          // [kMaxPipes]
  PhaseCtrl* h_ctrl = nullptr;      // [kMaxPipes] pinned
  cudaEvent_t fork = nullptr;
  double* slab2 = nullptr;          // second workspace slab (repack target)
  LoopBufs lb = {};                 // closed-loop driver buffers
  long lb_cap = 0;
  size_t slab2_doubles = 0;
  long cap = 0;                     // problems PER PIPE the lists / grids are sized for
  bool graph_failed = false;
  ResCtrl* res_ctrl = nullptr;      // problem queue of the resident layout
  bool res_attr_set = false;
  double* graph_slab = nullptr;     // the graphs bake these in: rebuild when they change
  long graph_stride = 0;
};

static void loop_free(mpcv_phase_state* s) {
  void* ptrs[] = {s->lb.x0, s->lb.p, s->lb.x, s->lb.f, s->lb.state, s->lb.status, s->lb.iters, s->lb.active,
                  s->lb.steps, s->lb.iters_total, s->lb.worst, s->lb.xctrl, s->lb.index, s->lb.count, s->lb.t0};
  for (void* q : ptrs) if (q) cudaFree(q);
  s->lb = LoopBufs{};
  s->lb_cap = 0;
}

static void phase_drop_graphs(mpcv_phase_state* s) {
  for (PhasePipe& q : s->pipe) {
    if (q.exec) { cudaGraphExecDestroy(q.exec); q.exec = nullptr; }
    if (q.graph) { cudaGraphDestroy(q.graph); q.graph = nullptr; }
  }
}

static void phase_free(mpcv_phase_state* s) {
  if (!s) return;
  phase_drop_graphs(s);
  for (PhasePipe& q : s->pipe) {
    if (q.stream) cudaStreamDestroy(q.stream);
    if (q.done) cudaEventDestroy(q.done);
  }
  if (s->fork) cudaEventDestroy(s->fork);
  if (s->lists) cudaFree(s->lists);
  if (s->res_ctrl) cudaFree(s->res_ctrl);
  if (s->slab2) cudaFree(s->slab2);
  loop_free(s);
  if (s->ctrl) cudaFree(s->ctrl);
  if (s->d_io) cudaFree(s->d_io);
  if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
  delete s;
}

// pipes for a batch of B problems
static int phase_pipes_for(const mpcv_handle* h, long B) {
  int k = h->knobs.pipes > 0 ? h->knobs.pipes : MPCV_PIPES_DEFAULT;
  if (h->knobs.hostloop) k = 1;   // the host loop synchronises: one pipe
  if (k > kMaxPipes) k = kMaxPipes;
  const long min_share = h->knobs.pipe_min > 0 ? h->knobs.pipe_min : MPCV_PIPE_MIN;
  while (k > 1 && B / k < min_share) --k;
  return k;
}

static int phase_ensure(mpcv_handle* h, long B, int K) {
  if (!h->phase) h->phase = new mpcv_phase_state();
  mpcv_phase_state* s = h->phase;
  if (!s->ctrl) {
    CUDA_OK(cudaMalloc(&s->ctrl, kMaxPipes * sizeof(PhaseCtrl)));
    // the memset runs on the legacy default stream, which does NOT order against non-blocking streams:
    // wait for it, or it can land after the first ph_begin_kernel and wipe B / n_act
    CUDA_OK(cudaMemset(s->ctrl, 0, kMaxPipes * sizeof(PhaseCtrl)));
    CUDA_OK(cudaStreamSynchronize(0));
    CUDA_OK(cudaMalloc(&s->d_io, kMaxPipes * sizeof(SolveIO)));
    CUDA_OK(cudaMallocHost(&s->h_ctrl, kMaxPipes * sizeof(PhaseCtrl)));
    memset(s->h_ctrl, 0, kMaxPipes * sizeof(PhaseCtrl));
    CUDA_OK(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    for (int j = 0; j < kMaxPipes; ++j) {
      s->pipe[j].ctrl = s->ctrl + j; s->pipe[j].d_io = s->d_io + j; s->pipe[j].h_ctrl = s->h_ctrl + j;
    }
  }
  for (int j = 1; j < K; ++j) {
    if (!s->pipe[j].stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&s->pipe[j].stream, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&s->pipe[j].done, cudaEventDisableTiming));
    }
  }
  const long share = (B + K - 1) / K;
  const long cap = (share + kPhaseThreads - 1) / kPhaseThreads * kPhaseThreads;
  bool relayout = false;
  if (cap > s->cap || K != s->npipes) {
    // keep the per-pipe capacity monotone so that alternating batch sizes do not reallocate every call
    const long ncap = cap > s->cap ? cap : s->cap;
    if ((long)K * ncap > (long)s->npipes * s->cap || !s->lists) {
      if (s->lists) cudaFree(s->lists);
      s->lists = nullptr;
      s->cap = 0; s->npipes = 0;
      CUDA_OK(cudaMalloc(&s->lists, (size_t)4 * K * ncap * sizeof(int)));
    }
    s->cap = ncap;
    s->npipes = K;
    relayout = true;
  }
  // workspace slabs sized for all pipes (pipe j owns the slots [j * cap, (j + 1) * cap))
  const size_t need = (size_t)s->npipes * s->cap * h->L.total;
  if (need > h->slab_doubles) {
    if (h->slab) cudaFree(h->slab);
    h->slab = nullptr;
    h->slab_doubles = 0;
    CUDA_OK(cudaMalloc(&h->slab, need * sizeof(double)));
    h->slab_doubles = need;
    relayout = true;
  }
  if (need > s->slab2_doubles) {
    if (s->slab2) cudaFree(s->slab2);
    s->slab2 = nullptr;
    s->slab2_doubles = 0;
    CUDA_OK(cudaMalloc(&s->slab2, need * sizeof(double)));
    s->slab2_doubles = need;
    relayout = true;
  }
  h->slab_stride = (long)s->npipes * s->cap;
  if (s->graph_slab != h->slab) relayout = true;
  if (relayout) {
    phase_drop_graphs(s);
    for (int j = 0; j < s->npipes; ++j) {
      PhasePipe& q = s->pipe[j];
      int* base = s->lists + (size_t)4 * j * s->cap;
      q.act[0] = base; q.act[1] = base + s->cap; q.retry = base + 2 * s->cap; q.slow = base + 3 * s->cap;
      q.slab[0] = h->slab + (size_t)j * s->cap * h->L.total;
      q.slab[1] = s->slab2 + (size_t)j * s->cap * h->L.total;
    }
    s->graph_slab = h->slab;
    s->graph_stride = h->slab_stride;
  }
  return 0;
}

template <class Model>
static PhaseArgs phase_args(const mpcv_handle* h, int j) {
  const mpcv_phase_state* s = h->phase;
  const PhasePipe& q = s->pipe[j];
  PhaseArgs a;
  a.P = h->P; a.L = h->L; a.slab[0] = q.slab[0]; a.slab[1] = q.slab[1];
  a.act[0] = q.act[0]; a.act[1] = q.act[1];
  a.retry = q.retry; a.slow = q.slow;
  a.ctrl = q.ctrl; a.io = q.d_io; a.cap = s->cap;
  return a;
}

static size_t phase_smem(const mpcv_handle* h) { return (size_t)h->L.n * sizeof(BndEntry); }
// warp-per-problem kernels (slow, tail): bounds table + one workspace row per warp, when it fits
// (MPCV_WARP_STAGED=0 keeps them on the slab: A/B measurements)
static size_t warp_staged_smem(const mpcv_handle* h) {
  return ph_rows_offset(h->L) + (size_t)(kWarpPhaseThreads / 32) * h->L.total * sizeof(double);
}
// bit 0: ph_slow_kernel, bit 1: ph_tail_kernel
static int warp_staged_mask(const mpcv_handle* h) {
  const int mask = h->knobs.warp_staged >= 0 ? h->knobs.warp_staged : 3;
  return warp_staged_smem(h) <= h->max_smem_optin ? mask : 0;
}
static bool warp_staged(const mpcv_handle* h) { return warp_staged_mask(h) != 0; }
static size_t warp_smem(const mpcv_handle* h) { return warp_staged(h) ? warp_staged_smem(h) : phase_smem(h); }
// fixed grids: enough CTAs to fill the GPU once (never more than the work of a full batch needs)
#ifndef MPCV_REPACK_CTAS
#define MPCV_REPACK_CTAS 11
#endif
struct PhaseGrids { unsigned prob, stage, warp, group, repack; };
static PhaseGrids phase_grids(const mpcv_handle* h) {
  const long cap = h->phase->cap, sm = h->sm_count;
  const long wpb = kWarpPhaseThreads / 32;
  const long pct = h->knobs.grid_pct;
  auto clampu = [pct](long need, long fill) {
    fill = fill * pct / 100;
    if (fill < 1) fill = 1;
    return (unsigned)(need < fill ? (need < 1 ? 1 : need) : fill);
  };
  PhaseGrids g;
  g.prob = clampu(cap / kPhaseThreads, sm * 4);
  // the repack copy is latency-bound (kRepackSplit threads per survivor, 44 registers): as many CTAs as an SM holds
  g.repack = clampu(cap * kRepackSplit / kPhaseThreads, sm * MPCV_REPACK_CTAS);
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

// hand-off threshold of the sweeps: min(B >> shift, cap) active problems.  The resident tail turns a problem over
// faster than a sweep does once the active set is small, so the hand-off comes earlier with it.
#ifndef MPCV_RES_TAIL_SHIFT
#define MPCV_RES_TAIL_SHIFT 2
#endif
#ifndef MPCV_RES_TAIL_BELOW
#define MPCV_RES_TAIL_BELOW 8192
#endif
static bool res_tail_usable(const mpcv_handle* h);
static int res_slots_per_sm(const mpcv_handle* h);
static int phase_tail_cap(const mpcv_handle* h) {
  return h->knobs.tail_cap >= 0 ? h->knobs.tail_cap : (res_tail_usable(h) ? MPCV_RES_TAIL_BELOW : kTailBelow);
}
static int phase_tail_shift(const mpcv_handle* h) {
  return h->knobs.tail_shift >= 0 ? h->knobs.tail_shift : (res_tail_usable(h) ? MPCV_RES_TAIL_SHIFT : MPCV_TAIL_SHIFT);
}

// ---- resident layout: one persistent kernel per solve (mpcv_resident.cuh) -------------------------------------
struct ResConfig { int slots, stride; size_t smem; unsigned grid; };
constexpr size_t kResStaticSmem = 1024;
// slots per CTA the shared memory of an SM allows with kResMinB CTAs resident (0: the workspace does not fit)
static int res_slots_fit(const mpcv_handle* h, int ctas_per_sm) {
  const size_t stride_b = (size_t)(h->L.total | 1) * sizeof(double), tab = ph_rows_offset(h->L);
  size_t budget = h->smem_per_sm / ctas_per_sm;
  budget = budget > 1024 ? budget - 1024 : 0;                       // the driver reserves 1 KB per CTA
  if (budget > h->max_smem_optin) budget = h->max_smem_optin;
  budget = budget > kResStaticSmem ? budget - kResStaticSmem : 0;   // the kernel's static shared memory (slot states)
  if (budget < tab + stride_b) return 0;
  const size_t sl = (budget - tab) / stride_b;
  return (int)(sl > (size_t)kResMaxSlots ? kResMaxSlots : sl);
}
static int res_slots_per_sm(const mpcv_handle* h) {
  const int a = res_slots_fit(h, kResMinB) * kResMinB;
  return a > 0 ? a : res_slots_fit(h, 1);
}
static int res_config(const mpcv_handle* h, long B, ResConfig* c) {
  int ctas = kResMinB;
  int S = res_slots_fit(h, ctas);
  if (S < 1) { ctas = 1; S = res_slots_fit(h, 1); }
  if (S < 1) return mpcv_set_error(-ENOMEM, "problem workspace exceeds shared memory; use MPCV_LAYOUT_PHASED");
  long grid = (long)h->sm_count * ctas;
  if (grid > B) grid = B;
  // small batches: spread the problems over the CTAs instead of filling the first few
  const long per = (B + grid - 1) / grid;
  if (per < S) S = (int)per;
  c->slots = S;
  c->stride = h->L.total | 1;
  c->smem = ph_rows_offset(h->L) + (size_t)S * c->stride * sizeof(double);
  c->grid = (unsigned)grid;
  return 0;
}

// queue heads (one per pipe for the resident tail + one for the stand-alone resident solve) and the kernels' opt-in
// to large dynamic shared memory
template <class Model>
static int res_prepare(mpcv_handle* h) {
  if (!h->phase) h->phase = new mpcv_phase_state();
  mpcv_phase_state* s = h->phase;
  if (!s->res_ctrl) {
    CUDA_OK(cudaMalloc(&s->res_ctrl, (kMaxPipes + 1) * sizeof(ResCtrl)));
    CUDA_OK(cudaMemset(s->res_ctrl, 0, (kMaxPipes + 1) * sizeof(ResCtrl)));
    CUDA_OK(cudaStreamSynchronize(0));
  }
  if (!s->res_attr_set) {
    cudaFuncAttributes fa;
    CUDA_OK(cudaFuncGetAttributes(&fa, res_solve_kernel<Model, false>));
    if (fa.sharedSizeBytes > kResStaticSmem) return mpcv_set_error(-EIO, "resident kernel: static shared memory above its budget");
    CUDA_OK(cudaFuncSetAttribute(res_solve_kernel<Model, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(h->max_smem_optin - kResStaticSmem)));
    CUDA_OK(cudaFuncSetAttribute(res_solve_kernel<Model, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(h->max_smem_optin - kResStaticSmem)));
    s->res_attr_set = true;
  }
  return 0;
}

// the stragglers of the sweeps finish in the resident kernel when an SM holds enough problems (else ph_tail_kernel)
constexpr int kResidentMinSlots = 8;
static bool res_tail_usable(const mpcv_handle* h) {
  // opt-in (MPCV_RES_TAIL=1): measured on C2 it is level with ph_tail_kernel on an average batch (15.5 vs 15.9 ms) and
  // worse on a batch with a 119-iteration straggler (19.2 vs 18.2 ms): a slot that shares its CTA turns over in
  // ~100 us per iteration, a lone warp of ph_tail_kernel in ~30 us
  if (h->knobs.res_tail != 1) return false;
  return res_slots_per_sm(h) >= kResidentMinSlots;
}
// layout of this call: AUTO = the resident kernel for batches below the crossover with the slab pipeline
// (measured on C2: 3.6 vs 7.2 ms at B = 4,096; 31 vs 16.5 ms at B = 65,536), the pipeline above it
#ifndef MPCV_RESIDENT_BELOW
#define MPCV_RESIDENT_BELOW 16384
#endif
static int effective_layout(const mpcv_handle* h, long B) {
  if (!h->layout_auto) return h->layout;
  const long below = h->knobs.resident_below >= 0 ? h->knobs.resident_below : MPCV_RESIDENT_BELOW;
  return (B < below && res_slots_per_sm(h) >= kResidentMinSlots) ? MPCV_LAYOUT_RESIDENT : MPCV_LAYOUT_PHASED;
}

template <class Model>
static ResArgs res_tail_args(const mpcv_handle* h, int j, const ResConfig& c) {
  const mpcv_phase_state* s = h->phase;
  const PhasePipe& q = s->pipe[j];
  ResArgs a = {};
  a.P = h->P; a.L = h->L; a.ctrl = s->res_ctrl + j; a.count = nullptr; a.index = nullptr; a.B = 0;
  a.slots = c.slots; a.stride = c.stride;
  a.pctrl = q.ctrl; a.io_dev = q.d_io; a.slab[0] = q.slab[0]; a.slab[1] = q.slab[1]; a.act[0] = q.act[0]; a.act[1] = q.act[1];
  return a;
}
// the resident tail of pipe j of K: its share of the CTAs the GPU holds, full slots
static int res_tail_config(const mpcv_handle* h, int K, ResConfig* c) {
  if (int rc = res_config(h, (long)h->sm_count * kResMinB * kResMaxSlots, c)) return rc;
  long grid = (long)h->sm_count * kResMinB / K;
  c->grid = (unsigned)(grid < 1 ? 1 : grid);
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
static int phase_build_graph(mpcv_handle* h, int j) {
  PhasePipe* s = &h->phase->pipe[j];
  PhaseArgs a = phase_args<Model>(h, j);
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
  int staged_slow = warp_staged_mask(h) & 1, staged_tail = (warp_staged_mask(h) >> 1) & 1;
  void* a_slow[] = {&a, &staged_slow};
  void* a_tail[] = {&a, &staged_tail};
  void* a_flip[] = {&s->ctrl, &handle, &use_handle};
  if (int rc = add_kernel(body, &b_pre, nullptr, (void*)ph_pre_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_repack, &b_pre, (void*)ph_repack_kernel<Model>, gr.repack, kPhaseThreads, 0, a_init)) return rc;
  if (int rc = add_kernel(body, &b_relist, &b_repack, (void*)ph_repack_list_kernel<Model>, gr.prob, kPhaseThreads, 0, a_init)) return rc;
  if (int rc = add_kernel(body, &b_factor, &b_relist, (void*)ph_factor_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_probe, &b_factor, (void*)ph_probe_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_retry, &b_probe, (void*)ph_retry_kernel<Model>, gr.prob, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_post, &b_retry, (void*)ph_post_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_trial, &b_post, (void*)ph_trial_kernel<Model>, gr.stage, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_accept, &b_trial, (void*)ph_accept_kernel<Model>, gr.group, kWarpPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_slow, &b_accept, (void*)ph_slow_kernel<Model>, gr.warp, kWarpPhaseThreads, warp_smem(h), a_slow)) return rc;
  if (int rc = add_kernel(body, &b_der, &b_slow, (void*)ph_der_kernel<Model>, gr.stage, kPhaseThreads, smem, a_init)) return rc;
  if (int rc = add_kernel(body, &b_flip, &b_der, (void*)ph_flip_kernel, 1, 1, 0, a_flip)) return rc;
  // the stragglers finish in one persistent kernel after the loop: the CTA-resident kernel (their workspaces staged
  // into shared memory, continuous refill from the active list) when an SM holds enough problems, else one warp each
  cudaGraphNode_t n_tail;
  if (res_tail_usable(h)) {
    ResConfig rc_;
    if (int rc = res_tail_config(h, h->phase->npipes, &rc_)) return rc;
    ResArgs ra = res_tail_args<Model>(h, j, rc_);
    void* a_res[] = {&ra};
    if (int rc = add_kernel(g, &n_tail, &n_while, (void*)res_solve_kernel<Model, true>, rc_.grid, kResThreads, rc_.smem, a_res)) return rc;
  } else {
    if (int rc = add_kernel(g, &n_tail, &n_while, (void*)ph_tail_kernel<Model>, gr.warp, kWarpPhaseThreads, warp_smem(h), a_tail)) return rc;
  }
  cudaGraphExec_t exec = nullptr;
  CUDA_OK(cudaGraphInstantiate(&exec, g, 0));
  s->graph = g;
  s->exec = exec;
  return 0;
}

// host-driven loop (fallback when the graph cannot be built, and MPCV_PHASE_HOSTLOOP=1): same
// kernels, the host reads the active count back every few sweeps
template <class Model>
static int phase_host_loop(mpcv_handle* h, cudaStream_t st) {
  PhasePipe* s = &h->phase->pipe[0];
  const PhaseArgs a = phase_args<Model>(h, 0);
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
      ph_repack_kernel<Model><<<gr.repack, kPhaseThreads, 0, st>>>(a);
      ph_repack_list_kernel<Model><<<gr.prob, kPhaseThreads, 0, st>>>(a);
      ph_factor_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_probe_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_retry_kernel<Model><<<gr.prob, kPhaseThreads, smem, st>>>(a);
      ph_post_kernel<Model><<<gr.group, kWarpPhaseThreads, smem, st>>>(a);
      ph_trial_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a);
      ph_accept_kernel<Model><<<gr.group, kWarpPhaseThreads, smem, st>>>(a);
      ph_slow_kernel<Model><<<gr.warp, kWarpPhaseThreads, warp_smem(h), st>>>(a, warp_staged_mask(h) & 1);
      ph_der_kernel<Model><<<gr.stage, kPhaseThreads, smem, st>>>(a);
      ph_flip_kernel<<<1, 1, 0, st>>>(s->ctrl, none, 0);
      h->launches += 12;
    }
    CUDA_OK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(PhaseCtrl), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (s->h_ctrl->n_act[s->h_ctrl->sweep & 1] <= s->h_ctrl->tail_below) break;
  }
  if (res_tail_usable(h)) {
    ResConfig rc_;
    if (int rc = res_tail_config(h, 1, &rc_)) return rc;
    res_solve_kernel<Model, true><<<rc_.grid, kResThreads, rc_.smem, st>>>(res_tail_args<Model>(h, 0, rc_));
  } else {
    ph_tail_kernel<Model><<<gr.warp, kWarpPhaseThreads, warp_smem(h), st>>>(a, (warp_staged_mask(h) >> 1) & 1);
  }
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class Model>
static int launch_solve_resident(mpcv_handle* h, const SolveIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if (B > 0x7fffffffL) return mpcv_set_error(-EINVAL, "batch too large");
  if (!h->phase) h->phase = new mpcv_phase_state();
  mpcv_phase_state* s = h->phase;
  if (int rc = res_prepare<Model>(h)) return rc;
  ResConfig c;
  if (int rc = res_config(h, B, &c)) return rc;
  if (const mpcv_host_xfer* xf = h->host_xfer)
    for (const auto& t : xf->in)
      if (t.host_src) CUDA_OK(cudaMemcpyAsync(t.dev, t.host_src, B * t.row_bytes, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(s->res_ctrl + kMaxPipes, 0, sizeof(ResCtrl), st));
  ResArgs a = {};
  a.P = h->P; a.L = h->L; a.io = io; a.ctrl = s->res_ctrl + kMaxPipes; a.count = io.count; a.index = io.index;
  a.B = B; a.slots = c.slots; a.stride = c.stride;
  res_solve_kernel<Model, false><<<c.grid, kResThreads, c.smem, st>>>(a);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  if (const mpcv_host_xfer* xf = h->host_xfer) {
    for (const auto& t : xf->out)
      if (t.host_dst) CUDA_OK(cudaMemcpyAsync(t.host_dst, t.dev, B * t.row_bytes, cudaMemcpyDeviceToHost, st));
    h->host_xfer_done = true;
  }
  return 0;
}

template <class Model>
static int launch_solve_phased(mpcv_handle* h, const SolveIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if (effective_layout(h, B) == MPCV_LAYOUT_RESIDENT) return launch_solve_resident<Model>(h, io, B, st);
  if (B > 0x7fffffffL) return mpcv_set_error(-EINVAL, "batch too large");
  const int K = phase_pipes_for(h, B);
  if (int rc = phase_ensure(h, B, K)) return rc;
  mpcv_phase_state* s = h->phase;
  const size_t smem = phase_smem(h);
  if (smem > h->max_smem_optin) return mpcv_set_error(-ENOMEM, "bounds table exceeds shared memory; use MPCV_LAYOUT_WARP");
  if (res_tail_usable(h)) { if (int rc = res_prepare<Model>(h)) return rc; }
  if (!s->pipe[0].exec && !s->graph_failed) {
    if (phase_set_smem(ph_init_kernel<Model>, smem) || phase_set_smem(ph_der0_kernel<Model>, smem) ||
        phase_set_smem(ph_init2_kernel<Model>, smem) || phase_set_smem(ph_pre_kernel<Model>, smem) ||
        phase_set_smem(ph_factor_kernel<Model>, smem) || phase_set_smem(ph_post_kernel<Model>, smem) ||
        phase_set_smem(ph_trial_kernel<Model>, smem) || phase_set_smem(ph_accept_kernel<Model>, smem) ||
        phase_set_smem(ph_retry_kernel<Model>, smem) || phase_set_smem(ph_probe_kernel<Model>, smem) || phase_set_smem(ph_slow_kernel<Model>, warp_smem(h)) ||
        phase_set_smem(ph_tail_kernel<Model>, warp_smem(h)) ||
        phase_set_smem(ph_der_kernel<Model>, smem))
      return -EIO;
  }
  const bool want_graph = !h->knobs.hostloop && !s->graph_failed;
  if (want_graph) {
    for (int j = 0; j < K && !s->graph_failed; ++j) {
      if (s->pipe[j].exec) continue;
      if (phase_build_graph<Model>(h, j) != 0) {
        // keep going with the host-driven loop; remember why
        s->graph_failed = true;
        cudaGetLastError();
        phase_drop_graphs(s);
      }
    }
  }
  if (want_graph && !s->graph_failed) {
    // contiguous shares (multiples of 32 problems, so each share starts on a slab block); pipe 0 runs on the
    // caller's stream, the others fork from it and join it again
    const long share = ph_share_of(B, K);
    const int tail_cap = phase_tail_cap(h), tail_shift = phase_tail_shift(h);
    if (K > 1) CUDA_OK(cudaEventRecord(s->fork, st));
    int rc = 0, forked = 0;
    for (int j = 0; j < K && rc == 0; ++j) {
      const long b0 = j * share, nb = (b0 + share <= B) ? share : B - b0;
      if (nb <= 0) break;
      PhasePipe& q = s->pipe[j];
      const cudaStream_t qs = j == 0 ? st : q.stream;
      // (a failing call returns from the lambda only: the pipes forked so far are still joined below)
      rc = [&]() -> int {
        if (j > 0) { CUDA_OK(cudaStreamWaitEvent(qs, s->fork, 0)); forked = j; }
        if (const mpcv_host_xfer* xf = h->host_xfer)
          for (const auto& t : xf->in)
            if (t.host_src) CUDA_OK(cudaMemcpyAsync(t.dev + b0 * t.row_bytes, t.host_src + b0 * t.row_bytes, nb * t.row_bytes, cudaMemcpyHostToDevice, qs));
        ph_begin_kernel<<<1, 1, 0, qs>>>(q.ctrl, q.d_io, io, B, j, K, tail_cap, tail_shift, s->res_ctrl ? &s->res_ctrl[j].next : nullptr);
        CUDA_OK(cudaGraphLaunch(q.exec, qs));
        if (const mpcv_host_xfer* xf = h->host_xfer)
          for (const auto& t : xf->out)
            if (t.host_dst) CUDA_OK(cudaMemcpyAsync(t.host_dst + b0 * t.row_bytes, t.dev + b0 * t.row_bytes, nb * t.row_bytes, cudaMemcpyDeviceToHost, qs));
        h->launches += 6;    // begin + init chain + tail; the sweeps are counted from the device (mpcv_phase_sweeps)
        return 0;
      }();
    }
    // join: the caller's stream waits for every side stream that was forked, also after an error
    for (int j = 1; j <= forked; ++j) {
      PhasePipe& q = s->pipe[j];
      if (cudaEventRecord(q.done, q.stream) == cudaSuccess) cudaStreamWaitEvent(st, q.done, 0);
    }
    if (rc) return rc;
    h->phase_graph_launches++;
    if (h->host_xfer) h->host_xfer_done = true;
    return 0;
  }
  // host-driven loop: one pipe over the whole batch (phase_pipes_for returns 1 when the environment asks for it;
  // after a graph failure re-lay the lists out for one pipe)
  if (K != 1) { if (int rc = phase_ensure(h, B, 1)) return rc; }
  if (const mpcv_host_xfer* xf = h->host_xfer)
    for (const auto& t : xf->in)
      if (t.host_src) CUDA_OK(cudaMemcpyAsync(t.dev, t.host_src, B * t.row_bytes, cudaMemcpyHostToDevice, st));
  ph_begin_kernel<<<1, 1, 0, st>>>(s->pipe[0].ctrl, s->pipe[0].d_io, io, B, 0, 1, phase_tail_cap(h), phase_tail_shift(h),
                                   s->res_ctrl ? &s->res_ctrl[0].next : nullptr);
  h->launches++;
  if (int rc = phase_host_loop<Model>(h, st)) return rc;
  if (const mpcv_host_xfer* xf = h->host_xfer) {
    for (const auto& t : xf->out)
      if (t.host_dst) CUDA_OK(cudaMemcpyAsync(t.host_dst, t.dev, B * t.row_bytes, cudaMemcpyDeviceToHost, st));
    h->host_xfer_done = true;
  }
  return 0;
}

// closed loop: per MPC step  prepare -> solve (graph) -> apply, all stream-ordered
template <class Model>
static int launch_loop_phased(mpcv_handle* h, const LoopIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if (effective_layout(h, B) == MPCV_LAYOUT_RESIDENT) { if (!h->phase) h->phase = new mpcv_phase_state(); }
  else if (int rc = phase_ensure(h, B, phase_pipes_for(h, B))) return rc;
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
    CUDA_OK(cudaMalloc(&s->lb.xctrl, B * nx * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->lb.index, B * sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.count, sizeof(int)));
    CUDA_OK(cudaMalloc(&s->lb.t0, sizeof(long long)));
    s->lb_cap = B;
  }
  const LoopBufs lb = s->lb;
  const unsigned grid = (unsigned)((B + 127) / 128);
  lp_begin_kernel<Model><<<grid, 128, 0, st>>>(h->L, io, lb, B);
  h->launches++;
  long long* const lat = h->latency_ns;
  h->latency_ns = nullptr;
  SolveIO sio{lb.x0, io.lbx, io.ubx, lb.p, lb.x, lb.f, nullptr, nullptr, nullptr, lb.status, lb.iters, nullptr};
  sio.index = lb.index;      // every step solves only the scenarios whose loop is still running
  sio.count = lb.count;
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
  // summed over the pipes: every sweep of every pipe is 12 kernel launches
  mpcv_phase_state* s = h->phase;
  CUDA_OK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, kMaxPipes * sizeof(PhaseCtrl), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  int tot = 0, cum = 0;
  for (int j = 0; j < kMaxPipes; ++j) {
    if (j < s->npipes) tot += s->h_ctrl[j].sweeps_total;
    cum += s->h_ctrl[j].sweeps_cum;
  }
  *sweeps = tot;
  *cumulative = cum;
  return 0;
}

#define MPCV_CAT2(a, b) a##b
#define MPCV_CAT(a, b) MPCV_CAT2(a, b)
extern const mpcv_phase_vtable MPCV_CAT(mpcv_phase_vtable_, MPCV_INST_MODEL) = {launch_solve_phased<ModelT>, phase_free,
                                                                                phase_sweeps, launch_loop_phased<ModelT>,
                                                                                res_slots_per_sm};
