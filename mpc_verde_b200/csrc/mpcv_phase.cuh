// mpcv_phase.cuh — batch-synchronous phase pipeline for the multiple-shooting solve.
//
// Replaces the same reference call as mpcv_ipm.cuh (`sol = solver(x0=,lbx=,ubx=,lbg=,ubg=,p=)`,
// Casadi/multiple_shooting_casadi.py:235-242; `solver.solve()` of the MPCTools scripts) and runs
// exactly the phase functions of Ipm<Model,false,1,WsStrided> — but each phase is its own
// batch-wide launch, so that
//   * every warp of the GPU executes the same small piece of code at the same time (the one-kernel
//     solve thrashes the 32 KB instruction cache: its iteration body is ~200 KB of SASS),
//   * each phase gets the parallel decomposition that fits it: the derivative sweep and the
//     line-search trial evaluation are independent per shooting interval and run one thread per
//     (problem, interval); the Riccati recursion is sequential over the horizon and runs one thread
//     per problem,
//   * finished problems leave the batch: an active list is compacted every iteration
//     (warp-aggregated atomics), so lanes never idle on problems that have converged,
//   * registers are allocated per phase instead of for the union of all phases.
// The per-problem workspace is the structure-of-arrays slab of the thread layout (element i of
// problem b at slab[(b/32)*32*total + i*32 + b%32]); the scalar state of a solve (mu, tau, filter, ...) is parked in
// ws[L.st ..] between launches.
//
// One interior-point iteration = pre -> repack -> factor -> probe -> retry -> post -> trial -> accept -> slow -> der -> flip:
//   pre     lane group / problem  error measures, convergence test (exports finished problems), barrier
//                                 update, Sigma / barrier gradient; ordered compaction of the active list
//   repack  thread per problem    (when an eighth of the slots have emptied) dense copy of the survivors
//   factor  thread per problem    Riccati factorisation with delta_w = 0 + vector recursions
//   probe   4 lanes / problem     the next four values of IPOPT's delta_w sequence tried at once; the first
//                                 success completes the step (only over the problems with wrong inertia)
//   retry   thread per problem    sequential walk through the rest of the sequence (rarely anything left)
//   post    lane group / problem  fraction-to-the-boundary step, merit-function terms
//   trial   thread per (problem, interval)   first line-search trial point
//   accept  lane group / problem  filter test of the full step, dual step, update of (w, z, lambda)
//   slow    warp per problem      backtracking / second-order correction, only over rejected problems
//   der     thread per (problem, interval)   derivative sweep at the new iterate
//   flip    one thread            list swap, repack commit, loop condition of the WHILE node
// then ph_tail_kernel: once at most min(4096, B/16) problems are active the loop ends and the stragglers
// finish one warp per problem, every phase in-kernel.
// A lane group is kGroupLanes = 4 lanes (kWideLanes = 8 below kWideBelow active problems).
// The sequential recursions run one thread per problem (every lane busy, coalesced rows of the blocked
// slab); everything that is parallel over variables or intervals runs with 4-10x more threads so
// that HBM latency is hidden by parallelism instead of being serialised inside one thread.
// The whole solve of a pipe is ONE CUDA graph whose iteration body sits in a conditional WHILE node, so
// mpcv_solve stays asynchronous on the caller's stream (no host round trip per iteration).  A large batch is split
// into several pipes — independent instances of this pipeline over contiguous shares of the batch, each with its own
// lists, control block and graph on its own stream (mpcv_phase_inst.cu) — so that the phases of different shares overlap.
// The warp-per-problem kernels (slow, tail) run on a shared-memory copy of their problem's workspace.
//
// The per-thread bodies below are plain functions shared by the CUDA kernels and by the CPU
// development harness (tests/hostsim), which replays the same schedule with one lane per problem.
#pragma once

#include "mpcv_driver.cuh"

namespace mpcv {

struct PhaseCtrl {
  int n_act[2];   // entries of the two ping-pong active lists
  int sweep;      // iteration sweeps done; list (sweep & 1) is the input of the next `pre` phase
  int B;          // problems of this call
  int sweeps_total;
  int n_retry;    // problems whose first factorisation had the wrong inertia (this sweep)
  int n_slow;     // problems whose full step was rejected by the filter (this sweep)
  int sweeps_cum; // sweeps since the handle was created (launch accounting)
  int cur;        // which of the two slabs holds the workspaces
  int slots;      // workspace slots in use since the last repack (the lists hold slot indices)
  int repacks;    // repacks of this call
  int tail_below; // at most this many active problems: leave the sweeps, finish in ph_tail_kernel
  int row0;       // first queue entry of this pipe's share (entry e of the call <-> I/O row index[e], or e)
};
constexpr int kRepackMin = 512;     // do not bother to repack fewer survivors than this
#ifndef MPCV_REPACK_SPLIT
#define MPCV_REPACK_SPLIT 16   /* threads per survivor of the repack copy, grid = as many CTAs as the SMs hold (same box, C2 per batch: 4 threads on the thread-per-problem grid 15.3 / 15.0 ms, 4 on the full grid 14.8 / 14.7, 8: 14.4, 16: 14.35, 32: 14.65) */
#endif
constexpr int kRepackSplit = MPCV_REPACK_SPLIT;
#ifndef MPCV_TAIL_BELOW
#define MPCV_TAIL_BELOW 4096   /* measured: 512: 24.6 ms, 1024: 24.1, 2048: 23.7, 4096: 23.1, 8192: 23.6, 16384: 24.4, all: 77 */
#endif
constexpr int kTailBelow = MPCV_TAIL_BELOW;    // at most this many active problems: leave the sweeps, finish in ph_tail_kernel
#ifndef MPCV_WIDE_BELOW
#define MPCV_WIDE_BELOW 8192   /* fewer active problems (per pipe) than this: latency-bound, use MPCV_WIDE_LANES lanes per problem; 4 pipes x 16,384: 16384: 16.6 ms, 8192: 16.1, 4096: 16.1 */
#endif
#ifndef MPCV_WIDE_LANES
#define MPCV_WIDE_LANES 8   /* C4 (4,096 scenarios x 100 steps): 4 lanes 354 ms, 8 lanes 299 ms, 32 lanes 426 ms */
#endif
constexpr int kWideBelow = MPCV_WIDE_BELOW;
constexpr int kWideLanes = MPCV_WIDE_LANES;    // lanes per problem of the lane-group kernels below kWideBelow active problems

template <class Model, class WS, int LANES = 1>
struct Phase {
  using IpmT = Ipm<Model, false, LANES, WS>;
  using Ipm1 = Ipm<Model, false, 1, WS>;
  static constexpr int NX = Model::NX, NH = Model::NX + Model::NPG;

  // the solver object of one problem: bounds shared by the batch (precomputed table) or the problem's own rows
  template <class I, class G>
  MPCV_HD static I make_ipm(const Params& P, const Layout& L, WS ws, G g, const SolveIO& io, const BndEntry* tab) {
    if (io.bstride) {
      const long b = (long)ws[L.st + 15];
      return I(P, L, ws, g, io.lbx + b * io.bstride, io.ubx + b * io.bstride, nullptr);
    }
    return I(P, L, ws, g, io.lbx, io.ubx, tab);
  }

  // load x0 / p, push into the interior, z0, lam0; state := running            (thread per problem)
  MPCV_HD static void init_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                const BndEntry* tab, long long t0) {
    const int np = NH + L.N * Model::NPS;
    for (int i = 0; i < L.n; ++i) ws[L.w + i] = io.x0 ? io.x0[b * L.n + i] : 0.0;
    for (int i = 0; i < np; ++i) ws[L.par + i] = io.p[b * np + i];
    ws[L.st + 15] = (double)b;          // the problem's row in the caller's batch (slots move on repack)
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.start();
    ipm.save_state(kRunning);
    ws[L.st + 14] = long_as_double(t0);
  }

  // derivative sweep, one shooting interval                                   (thread per interval)
  MPCV_HD static void der_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, int k, bool want_hess,
                               const BndEntry* tab) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.df = ws[L.st + 0];
    ws[L.qs + k] = ipm.der_stage(k, want_hess);
    if (k == 0) ipm.der_terminal();
  }

  // objective scaling + least-squares multipliers (after the first derivative sweep at df = 1)
  MPCV_HD static void init2_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.load_state();
    ipm.f_curr = ipm.sum_stage_costs();
    ipm.init_scaling_and_multipliers();
    ipm.save_state(kRunning);
  }

  // convergence test, barrier update, Sigma and barrier gradient.  Returns true when the problem stays
  // active; otherwise it is finished and its solution has been exported.        (LANES lanes per problem)
  MPCV_HD static bool pre_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                               const BndEntry* tab, Grp<LANES> g, long long now) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    int st = ipm.load_state();
    if (st != kRunning) return false;                 // failed earlier in this solve: already exported
    ipm.f_curr = ipm.sum_stage_costs();
    st = ipm.check_convergence_update_mu();
    if (st != kRunning) {
      finish(ipm, st, io, b, now);
      return false;
    }
    ipm.prepare_barrier();
    ipm.save_state(kRunning);
    return true;
  }

  // Riccati factorisation with delta_w = 0 and, when the inertia is right, the vector recursions.
  // Returns false when the delta_w schedule has to take over.                    (thread per problem)
  MPCV_HD static bool factor_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    if (!ipm.template riccati_factor_t<true>(0.0, false, 0, L.c)) return false;
    ipm.riccati_forward(L.c);
    return true;
  }

  // IPOPT's delta_w schedule is a deterministic sequence (next_delta_w): attempt `a` of the probe factorises
  // with its a-th element, WITHOUT stores; the first success of the sequence is what the sequential loop
  // would have stopped at.                                                      (kProbe lanes per problem)
  static constexpr int kProbe = 4;
  MPCV_HD static double probe_dw(const Ipm1& ipm, int a) {
    double dw = 0.0;
    for (int i = 0; i <= a; ++i) dw = ipm.next_delta_w(dw);
    return dw;
  }
  MPCV_HD static bool probe_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                 int attempt, double* dw_out) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.delta_w_last = ws[L.st + 6];
    const double dw = probe_dw(ipm, attempt);
    *dw_out = dw;
    if (dw > 1e20) return false;
    return ipm.riccati_probe(dw);
  }

  // the probe's winner: the same factorisation with stores + vector sweeps.  Should rounding make it fail
  // where the store-free probe succeeded, the problem goes to the sequential walk with hint = -dw.
  MPCV_HD static void apply_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                 double dw) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    if (!ipm.template riccati_factor_t<true>(dw, false, 0, L.c)) { ws[L.st + kSlotDwHint] = -dw; return; }
    ws[L.st + 6] = dw;
    ws[L.st + kSlotDwHint] = dw;          // > 0: done, ph_retry_kernel skips it
    ipm.riccati_forward(L.c);
  }
  // the same on ALL lanes of the winner's probe group: the lane-parallel factorisation and forward sweep
  // (riccati_factor_lanes: two short steps per stage with the operands exchanged through the dead step buffers, bit
  // for bit the register form) instead of one lane redoing ~350 dependent instructions per stage while its three
  // neighbours wait                                                               (LANES lanes per problem)
  MPCV_HD static void apply_lanes_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                       Grp<LANES> g, double dw) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    if (!ipm.template riccati_factor_x<true>(dw, false, 0, L.c)) {
      if (g.lane == 0) ws[L.st + kSlotDwHint] = -dw;
      return;
    }
    if (g.lane == 0) { ws[L.st + 6] = dw; ws[L.st + kSlotDwHint] = dw; }
    ipm.riccati_forward(L.c);
  }
  // inertia-correction retries for the problems the probe could not settle: walk on through the schedule
  // sequentially, then the vector sweeps                                          (thread per problem)
  MPCV_HD static void retry_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                 const BndEntry* tab, long long now) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.delta_w_last = ws[L.st + 6];
    const double hint = ws[L.st + kSlotDwHint];
    if (hint > 0.0) return;               // settled by the probe's winning lane
    double dw = hint < 0.0 ? ipm.next_delta_w(-hint) : ipm.next_delta_w(0.0);
    bool ok = false;
    while (dw <= 1e20) {
      ok = ipm.template riccati_factor_t<true>(dw, false, 0, L.c);
      if (ok) break;
      dw = ipm.next_delta_w(dw);
    }
    if (!ok) {
      ipm.load_state();
      finish(ipm, MPCV_ERROR_IN_STEP_COMPUTATION, io, b, now);
      return;
    }
    ws[L.st + 6] = dw;
    ipm.riccati_forward(L.c);
  }

  MPCV_HD static bool running(const Layout& L, WS ws) { return (int)ws[L.st + 9] == kRunning; }

  // fraction-to-the-boundary step and merit-function terms                      (LANES lanes per problem)
  MPCV_HD static void post_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                Grp<LANES> g) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    ipm.load_state();
    ipm.direction_post();
    ipm.save_state(kRunning);
  }

  // first line-search trial point, one shooting interval                         (thread per interval)
  MPCV_HD static void trial_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, int k,
                                 const BndEntry* tab) {
    Ipm1 ipm = make_ipm<Ipm1>(P, L, ws, Grp<1>(0), io, tab);
    ipm.trial_stage(k, ws[L.st + 10], L.d);
  }

  // fast path of the line search: the full step passes the filter.  Returns false when the problem needs
  // the slow path (nothing has been modified then).                              (LANES lanes per problem)
  MPCV_HD static bool accept_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                  Grp<LANES> g) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    ipm.load_state();
    if (!ipm.line_search_first()) return false;
    ipm.save_state(kRunning);
    return true;
  }

  // slow path: backtracking and second-order correction                          (LANES lanes per problem)
  MPCV_HD static void slow_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                const BndEntry* tab, Grp<LANES> g, long long now) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    ipm.load_state();
    const int st = ipm.line_search(true);
    if (st != 0) { finish(ipm, st, io, b, now); return; }
    ipm.save_state(kRunning);
  }

  // Straggler tail: run the remaining iterations of ONE problem to completion, all phases in sequence on
  // the group's lanes (what Ipm::solve() does after its start-up), from the state the sweeps left behind.
  // Precondition: a derivative sweep at the current iterate has been done (as at the top of every sweep).
  MPCV_HD static void tail_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                const BndEntry* tab, Grp<LANES> g, long long now) {
    IpmT ipm = make_ipm<IpmT>(P, L, ws, g, io, tab);
    int st = ipm.load_state();
    if (st != kRunning) return;
    ipm.f_curr = ipm.sum_stage_costs();
    for (;;) {
      st = ipm.check_convergence_update_mu();
      if (st != kRunning) break;
      st = ipm.compute_direction();
      if (st == 0) st = ipm.line_search(false);
      if (st != 0) break;
      ipm.eval_derivatives(true);
    }
    finish(ipm, st, io, b, now);
  }

  template <class I>
  MPCV_HD static void finish(I& ipm, int status, const SolveIO& io, long /*slot*/, long long now) {
    const long b = (long)ipm.ws[ipm.L.st + 15];
    const SolveInfo info = ipm.finish(status);
    export_solution(ipm, info, io, b);
    ipm.save_state(status);
    if (io.ns && ipm.g.lane == 0) io.ns[b] = now - double_as_long(ipm.ws[ipm.L.st + 14]);
  }

  MPCV_HD static double long_as_double(long long v) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(v);
#else
    double d; memcpy(&d, &v, sizeof d); return d;
#endif
  }
  MPCV_HD static long long double_as_long(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long v; memcpy(&v, &d, sizeof v); return v;
#endif
  }
};

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long ph_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// cooperative fill of the relaxed-bounds table in dynamic shared memory
template <class Model>
__device__ __forceinline__ const BndEntry* ph_bounds_table(const Params& P, const Layout& L, const SolveIO& io) {
  extern __shared__ __align__(16) unsigned char ph_smem[];
  BndEntry* tab = reinterpret_cast<BndEntry*>(ph_smem);
  Ipm<Model, false, 1, WsStrided> ipm(P, L, WsStrided{nullptr}, Grp<1>(0), io.lbx, io.ubx, nullptr);
  for (int i = threadIdx.x; i < L.n; i += blockDim.x) tab[i] = ipm.bnd_entry(i);
  __syncthreads();
  return tab;
}

struct PhaseArgs {
  Params P;
  Layout L;
  double* slab[2];     // ping-pong: survivors are repacked densely from one into the other
  int* act[2];
  int* retry;          // problems needing inertia-correction retries this sweep
  int* slow;           // problems needing the slow line-search path this sweep
  PhaseCtrl* ctrl;
  const SolveIO* io;   // device copy of the call's I/O pointers
  long cap;            // problems the lists and the slab are sized for
};

constexpr int kPhaseThreads = 128;       // thread-per-problem / thread-per-interval kernels
#ifndef MPCV_GROUP_THREADS
#define MPCV_GROUP_THREADS 128
#endif
#ifndef MPCV_GROUP_LANES
#define MPCV_GROUP_LANES 4   /* measured: 16 lanes 33.6 ms, 8: 28.2, 4: 27.5 (r1b); 8 -> 4 again 21.7 -> 21.3 ms (r1c) */
#endif
#ifndef MPCV_REPACK_NUM
#define MPCV_REPACK_NUM 7   /* repack when n_active * DEN <= slots * NUM (1/2: 30.3 ms, 3/4: 28.2, 7/8: 27.6) */
#endif
#ifndef MPCV_REPACK_DEN
#define MPCV_REPACK_DEN 8
#endif
constexpr int kWarpPhaseThreads = MPCV_GROUP_THREADS;   // lane-group kernels
// pre / post / accept: kGroupLanes lanes per problem, the problems of a warp being neighbours in the active list.
// Scalar work (pow, filter logic, state hand-over) is then shared by 32 / kGroupLanes problems per warp
// instruction, and a warp-load touches whole sectors: neighbouring problems share each 32-byte sector of the slab.
constexpr int kGroupLanes = MPCV_GROUP_LANES;
#ifndef MPCV_GROUP_MINB
#define MPCV_GROUP_MINB 8   /* 64 registers at 128 threads: measured best on B200 (128 regs: 29.6 ms, 85: 28.2, 64: 27.3 per batch) */
#endif

// Every phase kernel is launched with a FIXED grid (the launches are nodes of a CUDA graph) sized to
// fill the GPU once, and strides over its work list: a sweep that has three problems left costs a few
// microseconds per launch instead of a full grid of CTAs that start only to exit.

template <class Model>
__global__ void __launch_bounds__(kPhaseThreads) ph_init_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[a.ctrl->cur];
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  const long B = a.ctrl->B, row0 = a.ctrl->row0;
  for (long b = (long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long)gridDim.x * blockDim.x) {
    a.act[0][b] = (int)b;
    const long row = io.index ? (long)io.index[row0 + b] : row0 + b;
    Phase<Model, WsStrided>::init_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, row, tab,
                                       io.ns ? ph_globaltimer() : 0);
  }
}

// first derivative sweeps (all problems): thread per (problem, interval)
template <class Model>
__global__ void __launch_bounds__(kPhaseThreads) ph_der0_kernel(const __grid_constant__ PhaseArgs a, int want_hess) {
  double* const slab = a.slab[a.ctrl->cur];
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  const long B = a.ctrl->B, items = B * a.L.N;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const long b = it % B;
    const int k = (int)(it / B);
    Phase<Model, WsStrided>::der_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, k, want_hess != 0, tab);
  }
}

template <class Model>
__global__ void __launch_bounds__(kPhaseThreads) ph_init2_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[a.ctrl->cur];
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  const long B = a.ctrl->B;
  for (long b = (long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long)gridDim.x * blockDim.x)
    Phase<Model, WsStrided>::init2_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, tab);
}

// warp-aggregated append of the flagged lanes' problem indices to a list (order kept within the warp)
__device__ __forceinline__ void ph_append(bool flag, int b, int* list, int* count) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = b;
  }
}

// pre: LANES lanes per problem.  The surviving problems of a CTA pass are appended to the output list as
// ONE contiguous, order-preserving run, so the thread-per-problem Riccati kernels that follow keep
// reading neighbouring problems in neighbouring lanes (whole 32-byte sectors of the blocked slab).
template <class Model, int LANES>
__device__ __forceinline__ void ph_pre_run(const PhaseArgs& a, double* slab, const SolveIO& io, const BndEntry* tab,
                                           int in, int out, int n_in, int* keep_s, int* b_s, int* base_s) {
  constexpr int GPB = kWarpPhaseThreads / LANES;   // problems per CTA pass
  static_assert(GPB <= 32, "one ballot covers the CTA's problems");
  const int grp = threadIdx.x / LANES, lane = threadIdx.x & 31;
  const Grp<LANES> g(lane);
  for (long e0 = (long)blockIdx.x * GPB; e0 < n_in; e0 += (long)gridDim.x * GPB) {
    const long e = e0 + grp;
    bool keep = false;
    int b = 0;
    if (e < n_in) {
      b = a.act[in][e];
      keep = Phase<Model, WsStrided, LANES>::pre_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, b, tab, g,
                                                      io.ns ? ph_globaltimer() : 0);
    }
    if (g.lane == 0) { keep_s[grp] = keep ? 1 : 0; b_s[grp] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
      const bool k = lane < GPB && keep_s[lane] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (lane == 0) *base_s = m ? atomicAdd(&a.ctrl->n_act[out], __popc(m)) : 0;
      __syncwarp();
      if (k) a.act[out][*base_s + __popc(m & ((1u << lane) - 1u))] = b_s[lane];
    }
    __syncthreads();
  }
}

template <class Model>
__global__ void __launch_bounds__(kWarpPhaseThreads, MPCV_GROUP_MINB) ph_pre_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[a.ctrl->cur];
  __shared__ int keep_s[32], b_s[32], base_s;
  const int in = a.ctrl->sweep & 1, out = in ^ 1;
  const int n_in = a.ctrl->n_act[in];
  const bool wide = n_in < kWideBelow;
  if ((long)blockIdx.x * (kWarpPhaseThreads / (wide ? kWideLanes : kGroupLanes)) >= n_in) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  if (wide) ph_pre_run<Model, kWideLanes>(a, slab, io, tab, in, out, n_in, keep_s, b_s, &base_s);
  else ph_pre_run<Model, kGroupLanes>(a, slab, io, tab, in, out, n_in, keep_s, b_s, &base_s);
}

// Repack: once a quarter of the slots have emptied, the survivors' workspaces are copied densely into the
// other slab (slot e <- slot list[e]) and the list becomes the identity.  Between repacks at least 75% of
// every 32-byte sector a warp touches is live data; without it the survivors end up one per sector and
// every phase moves 4x the bytes it uses.
__device__ __forceinline__ bool ph_repack_wanted(const PhaseCtrl* c, int n) {
  return n >= kRepackMin && (long)n * MPCV_REPACK_DEN <= (long)c->slots * MPCV_REPACK_NUM;
}

template <class Model>
__global__ void __launch_bounds__(kPhaseThreads) ph_repack_kernel(const __grid_constant__ PhaseArgs a) {
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const int n = a.ctrl->n_act[out];
  if (!ph_repack_wanted(a.ctrl, n)) return;
  const double* src = a.slab[a.ctrl->cur];
  double* dst = a.slab[a.ctrl->cur ^ 1];
  const int total = a.L.total;
  // kRepackSplit threads per survivor, each copying an interleaved share of the elements (the copy is latency-
  // bound: more threads in flight, same bytes)
  const long items = (long)n * kRepackSplit;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const long e = it % n;                      // a warp copies the same share of 32 neighbouring survivors
    const int part = (int)(it / n);
    const int sl = a.act[out][e];
    const double* ps = src + (long)(sl >> 5) * ((long)total * 32) + (sl & 31);
    double* pd = dst + (e >> 5) * ((long)total * 32) + (e & 31);
    // live after `pre`: everything except the step (d, lam+), the trial residuals, the Riccati factors and
    // cost-to-go and the per-interval costs, which the rest of the sweep rewrites before reading
    const Layout& L = a.L;
    auto live = [&](int i) {
      return i < total && !((i >= L.d && i < L.d + L.n) || (i >= L.qs && i < L.qs + L.N) ||
                            (i >= L.lamp && i < L.lamp + L.m) || (i >= L.ct && i < L.ct + L.m) ||
                            (L.park != L.hw && i >= L.park && i < L.park + L.n + L.m) ||
                            i >= L.ric);          // ric and pp are the last two regions
    };
    const int nchunks = (total + 7) / 8;
    for (int c = part; c < nchunks; c += kRepackSplit) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = live(c * 8 + j) ? ps[(long)(c * 8 + j) * 32] : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (live(c * 8 + j)) pd[(long)(c * 8 + j) * 32] = v[j];
    }
  }
}

// after the copy (all threads of ph_repack_kernel are done): the list becomes the identity
template <class Model>
__global__ void ph_repack_list_kernel(const __grid_constant__ PhaseArgs a) {
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const int n = a.ctrl->n_act[out];
  if (!ph_repack_wanted(a.ctrl, n)) return;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) a.act[out][e] = (int)e;
}

// slab index for the kernels that run after the repack of the current sweep (the decision is a pure
// function of the control block, which does not change between `pre` and `flip`; `flip` commits it)
__device__ __forceinline__ int ph_cur_after_repack(const PhaseCtrl* c) {
  const int n = c->n_act[(c->sweep & 1) ^ 1];
  return ph_repack_wanted(c, n) ? c->cur ^ 1 : c->cur;
}

// register cap of the thread-per-problem Riccati kernels (CTAs per SM they are compiled for): Model::FAC_MINB
#ifdef MPCV_FAC_MINB
#define MPCV_FAC_MINB_OF(Model) MPCV_FAC_MINB
#else
#define MPCV_FAC_MINB_OF(Model) Model::FAC_MINB
#endif
template <class Model>
__global__ void __launch_bounds__(kPhaseThreads, MPCV_FAC_MINB_OF(Model)) ph_factor_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const int n = a.ctrl->n_act[out];
  if ((long)blockIdx.x * blockDim.x >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long e0 = (long)blockIdx.x * blockDim.x; e0 < n; e0 += (long)gridDim.x * blockDim.x) {
    const long e = e0 + threadIdx.x;
    bool retry = false;
    int b = 0;
    if (e < n) {
      b = a.act[out][e];
      retry = !Phase<Model, WsStrided>::factor_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, tab);
    }
    ph_append(retry, b, a.retry, &a.ctrl->n_retry);
  }
}

#ifndef MPCV_PROBE_LANES
#define MPCV_PROBE_LANES 0   /* 1: the probe's winner redoes its factorisation on all four lanes of the group (apply_lanes_body) instead of one; same box, C2 per batch, five processes each: 14.1 - 14.6 ms (0) against 15.0 - 16.3 (1) — the exchanges go through the slab here, not through shared memory */
#endif
// probe: kProbe lanes per problem of the retry list, each trying one element of the delta_w sequence
template <class Model>
__global__ void __launch_bounds__(kPhaseThreads, 4) ph_probe_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  constexpr int NP = Phase<Model, WsStrided>::kProbe;
  const long n = a.ctrl->n_retry, items = n * NP;
  if ((long)blockIdx.x * blockDim.x >= items) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long i0 = (long)blockIdx.x * blockDim.x; i0 < items; i0 += (long)gridDim.x * blockDim.x) {
    const long it = i0 + threadIdx.x;
    const int attempt = threadIdx.x & (NP - 1);
    bool ok = false;
    double dw = 0.0;
    WsStrided ws{nullptr};
    if (it < items) {
      ws = WsStrided::of(slab, a.L.total, a.retry[it / NP]);
      ok = Phase<Model, WsStrided>::probe_body(a.P, a.L, ws, io, tab, attempt, &dw);
    }
    // first success within the group of NP lanes
    const unsigned lane = threadIdx.x & 31, gbase = lane & ~(NP - 1);
    const unsigned m = (__ballot_sync(0xffffffffu, ok) >> gbase) & ((1u << NP) - 1u);
    const int first = m ? __ffs(m) - 1 : NP - 1;
    const double dsel = __shfl_sync(0xffffffffu, dw, gbase + first);
    // the winning lane repeats its factorisation with stores (operands are in L1/L2 now) and runs the vector
    // sweeps; only when none of the probed values worked the problem is left to ph_retry_kernel.  (Handing the
    // stored factorisation to a thread-per-problem launch over the retry list instead: 17.4 -> 17.9 ms.)
#if MPCV_PROBE_LANES
    if (it < items && m) Phase<Model, WsStrided, NP>::apply_lanes_body(a.P, a.L, ws, io, tab, Grp<NP>((int)lane), dsel);
#else
    if (it < items && m && attempt == first) Phase<Model, WsStrided>::apply_body(a.P, a.L, ws, io, tab, dsel);
#endif
    if (it < items && !m && attempt == 0) ws[a.L.st + kSlotDwHint] = -dsel;
  }
}

template <class Model>
__global__ void __launch_bounds__(kPhaseThreads, 4) ph_retry_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int n = a.ctrl->n_retry;
  if ((long)blockIdx.x * blockDim.x >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int b = a.retry[e];
    Phase<Model, WsStrided>::retry_body(a.P, a.L, WsStrided::of(slab, a.L.total, b), io, b, tab,
                                        io.ns ? ph_globaltimer() : 0);
  }
}

template <class Model, int LANES>
__device__ __forceinline__ void ph_post_run(const PhaseArgs& a, double* slab, const SolveIO& io, const BndEntry* tab,
                                            int out, int n) {
  const int gpb = blockDim.x / LANES, grp = threadIdx.x / LANES;
  const Grp<LANES> g(threadIdx.x & 31);
  for (long e = (long)blockIdx.x * gpb + grp; e < n; e += (long)gridDim.x * gpb) {
    const int b = a.act[out][e];
    const WsStrided ws = WsStrided::of(slab, a.L.total, b);
    if (!Phase<Model, WsStrided>::running(a.L, ws)) continue;     // retries exhausted: already exported
    Phase<Model, WsStrided, LANES>::post_body(a.P, a.L, ws, io, tab, g);
  }
}

template <class Model>
__global__ void __launch_bounds__(kWarpPhaseThreads, MPCV_GROUP_MINB) ph_post_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const int n = a.ctrl->n_act[out];
  const bool wide = n < kWideBelow;
  if ((long)blockIdx.x * (kWarpPhaseThreads / (wide ? kWideLanes : kGroupLanes)) >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  if (wide) ph_post_run<Model, kWideLanes>(a, slab, io, tab, out, n);
  else ph_post_run<Model, kGroupLanes>(a, slab, io, tab, out, n);
}

template <class Model>
__global__ void __launch_bounds__(kPhaseThreads) ph_trial_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const long n = a.ctrl->n_act[out], items = n * a.L.N;
  if ((long)blockIdx.x * blockDim.x >= items) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int b = a.act[out][it % n];
    const int k = (int)(it / n);
    const WsStrided ws = WsStrided::of(slab, a.L.total, b);
    if (!Phase<Model, WsStrided>::running(a.L, ws)) continue;
    Phase<Model, WsStrided>::trial_body(a.P, a.L, ws, io, k, tab);
  }
}

// The decision (switching condition, Armijo / filter tests, filter augmentation) is branchy scalar code on which
// the problems of a warp part ways; the step that follows is two loops over the variables.  Left to the compiler
// the groups do not come back together before the loops (measured at sweep 9 of C2: the update ran twice per warp
// with 12 of 32 lanes each), so the warp is synchronised explicitly between the two halves; the pass loop is
// warp-uniform for that.
template <class Model, int LANES>
__device__ __forceinline__ void ph_accept_run(const PhaseArgs& a, double* slab, const SolveIO& io, const BndEntry* tab,
                                              int out, int n) {
  using IpmT = Ipm<Model, false, LANES, WsStrided>;
  const int gpb = blockDim.x / LANES, grp = threadIdx.x / LANES;
  const Grp<LANES> g(threadIdx.x & 31);
  for (long e0 = (long)blockIdx.x * gpb; e0 < n; e0 += (long)gridDim.x * gpb) {
    const long e = e0 + grp;
    bool act = e < n;
    const int b = act ? a.act[out][e] : 0;
    const WsStrided ws = WsStrided::of(slab, a.L.total, b);
    act = act && Phase<Model, WsStrided>::running(a.L, ws);
    IpmT ipm = Phase<Model, WsStrided, LANES>::template make_ipm<IpmT>(a.P, a.L, ws, g, io, tab);
    typename IpmT::LsFirst r;
    r.ok = false;
    if (act) {
      ipm.load_state();
      r = ipm.line_search_first_decide();
      if (r.ok) ipm.ls_filter_augment(ipm.ls_alpha_max, r.phi_t, r.pw);
    }
    __syncwarp();
    if (act && r.ok) {
      ipm.ls_take_step(ipm.ls_alpha_max);
      ipm.save_state(kRunning);
    } else if (act && g.lane == 0) {
      a.slow[atomicAdd(&a.ctrl->n_slow, 1)] = b;
    }
  }
}

template <class Model>
__global__ void __launch_bounds__(kWarpPhaseThreads, MPCV_GROUP_MINB) ph_accept_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const int n = a.ctrl->n_act[out];
  const bool wide = n < kWideBelow;
  if ((long)blockIdx.x * (kWarpPhaseThreads / (wide ? kWideLanes : kGroupLanes)) >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  if (wide) ph_accept_run<Model, kWideLanes>(a, slab, io, tab, out, n);
  else ph_accept_run<Model, kGroupLanes>(a, slab, io, tab, out, n);
}

// ---- warp-per-problem kernels: the problem's workspace staged in shared memory ------------------------------
// ph_slow_kernel and ph_tail_kernel run a long dependent chain per problem (trial evaluations, Riccati re-solves,
// whole iterations) on a few warps: every access to the slab is an exposed HBM round trip.  They are register-
// bound at two CTAs per SM, so shared memory is free: the warp copies its problem's workspace into a shared row
// (cp.async, every element in flight at once), runs the unchanged phase body on the row, and writes back what
// the sweeps read afterwards (slow) or nothing at all (tail: the problem is finished and exported from the row).
// (generic addressing on purpose: with __builtin_assume(__isShared(p)) nvcc 12.9 miscompiles the 32-lane phase
// bodies — the same source on generic pointers is bit-identical to the run on the slab)
// Guard: tests/test_gpu_parity.py::test_warp_kernels_on_shared_rows_equal_the_slab_run (staged run == run on the slab,
// bit for bit).  -DMPCV_WSSHARED_ASSUME=1 (scripts/build_variant.sh as 0 -DMPCV_WSSHARED_ASSUME=1) puts the assumption
// back: round 1's code failed 750 of 1,024 tail problems with it (nvcc 12.9.86, sm_100a); on the end-of-round code
// (lane-parallel Riccati in the 32-lane bodies) that build passes the guard test — the assumption stays off all the
// same, an LDS instead of a generic load is not worth a latent miscompile.
#ifndef MPCV_WSSHARED_ASSUME
#define MPCV_WSSHARED_ASSUME 0
#endif
struct WsShared {
  double* row;
  __device__ __forceinline__ double& operator[](int i) const {
#if MPCV_WSSHARED_ASSUME
    __builtin_assume(__isShared(row + i));
#endif
    return row[i];
  }
  __device__ __forceinline__ WsShared view(int off) const { return WsShared{row + off}; }
};
__device__ __forceinline__ void ph_cp_async8(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void ph_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// shared rows of a CTA start after the bounds table
__host__ __device__ inline size_t ph_rows_offset(const Layout& L) { return ((size_t)L.n * sizeof(BndEntry) + 15) & ~(size_t)15; }
__device__ __forceinline__ double* ph_row_of_warp(const Layout& L, int warp) {
  extern __shared__ __align__(16) unsigned char ph_smem[];
  return reinterpret_cast<double*>(ph_smem + ph_rows_offset(L)) + (long)warp * L.total;
}
__device__ __forceinline__ void ph_stage_in(const Layout& L, const WsStrided& ws, double* row, int lane) {
  for (int i = lane; i < L.total; i += 32) ph_cp_async8(row + i, &ws[i]);
  ph_cp_async_wait();
  __syncwarp();
}

// `staged` = 1: dynamic shared memory holds one workspace row per warp
template <class Model>
__global__ void __launch_bounds__(kWarpPhaseThreads) ph_slow_kernel(const __grid_constant__ PhaseArgs a, int staged) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int n = a.ctrl->n_slow;
  const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((long)blockIdx.x * wpb >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  const Layout& L = a.L;
  for (long e = (long)blockIdx.x * wpb + warp; e < n; e += (long)gridDim.x * wpb) {
    const int b = a.slow[e];
    const WsStrided ws = WsStrided::of(slab, L.total, b);
    const long long now = io.ns ? ph_globaltimer() : 0;
    if (staged) {
      double* const row = ph_row_of_warp(L, warp);
      ph_stage_in(L, ws, row, lane);
      const WsShared sw{row};
      Phase<Model, WsShared, 32>::slow_body(a.P, L, sw, io, b, tab, Grp<32>(lane), now);
      __syncwarp();
      // what the sweeps read next: the iterate, its multipliers and the solve state (d, lam+, the trial residuals
      // and the parked step in hw are dead: the derivative sweep and the next factorisation rewrite them)
      auto put = [&](int off, int cnt) { for (int i = lane; i < cnt; i += 32) ws[off + i] = sw[off + i]; };
      put(L.w, L.n); put(L.zl, L.n); put(L.zu, L.n); put(L.lam, L.m); put(L.st, kStateSlots);
      __syncwarp();
    } else {
      Phase<Model, WsStrided, 32>::slow_body(a.P, L, ws, io, b, tab, Grp<32>(lane), now);
    }
  }
}

// Register cap of the derivative sweep (CTAs per SM it is compiled for, Model::DER_MINB).  The unicycle sweep wants
// 162 registers (3 CTAs per SM: 2.6 warps per scheduler, FP64 pipe 59 % busy, profiles/r2h_full_metrics.csv); capped at
// 128 it spills 132 bytes and runs 4 CTAs per SM — same box, C2 per batch: 15.0 - 15.3 ms uncapped, 14.4 at 128
// registers, 15.0 - 15.6 at 96 (416 bytes of spills).  The Frenet jets already spill at 255: no cap.
#ifdef MPCV_DER_MINB
#define MPCV_DER_MINB_OF(Model) MPCV_DER_MINB
#else
#define MPCV_DER_MINB_OF(Model) Model::DER_MINB
#endif
template <class Model>
__global__ void __launch_bounds__(kPhaseThreads, MPCV_DER_MINB_OF(Model)) ph_der_kernel(const __grid_constant__ PhaseArgs a) {
  double* const slab = a.slab[ph_cur_after_repack(a.ctrl)];
  const int out = (a.ctrl->sweep & 1) ^ 1;
  const long n = a.ctrl->n_act[out], items = n * a.L.N;
  if ((long)blockIdx.x * blockDim.x >= items) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
    const int b = a.act[out][it % n];
    const int k = (int)(it / n);
    const WsStrided ws = WsStrided::of(slab, a.L.total, b);
    if (!Phase<Model, WsStrided>::running(a.L, ws)) continue;   // failed this sweep: nothing to refresh
    Phase<Model, WsStrided>::der_body(a.P, a.L, ws, io, k, true, tab);
  }
}

// After the WHILE loop: the stragglers (at most kTailBelow problems, typically the 1-2 % that need 2-4x the
// mean iteration count) finish here, one warp per problem, every phase in-kernel.  A sweep over so few
// problems is pure launch-plus-latency floor (~200 us for 10 dependent launches); in here an iteration costs
// what its own dependent chain costs and problems do not wait for each other.
#ifndef MPCV_TAIL_MINB
#define MPCV_TAIL_MINB 1
#endif
template <class Model>
__global__ void __launch_bounds__(kWarpPhaseThreads, MPCV_TAIL_MINB) ph_tail_kernel(const __grid_constant__ PhaseArgs a, int staged) {
  double* const slab = a.slab[a.ctrl->cur];
  const int in = a.ctrl->sweep & 1;
  const int n = a.ctrl->n_act[in];
  const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((long)blockIdx.x * wpb >= n) return;
  const SolveIO io = *a.io;
  const BndEntry* tab = ph_bounds_table<Model>(a.P, a.L, io);
  for (long e = (long)blockIdx.x * wpb + warp; e < n; e += (long)gridDim.x * wpb) {
    const int b = a.act[in][e];
    const WsStrided ws = WsStrided::of(slab, a.L.total, b);
    const long long now = io.ns ? ph_globaltimer() : 0;
    if (staged) {
      // the problem runs to completion and is exported from the row: nothing goes back to the slab
      double* const row = ph_row_of_warp(a.L, warp);
      ph_stage_in(a.L, ws, row, lane);
      Phase<Model, WsShared, 32>::tail_body(a.P, a.L, WsShared{row}, io, b, tab, Grp<32>(lane), now);
      __syncwarp();
    } else {
      Phase<Model, WsStrided, 32>::tail_body(a.P, a.L, ws, io, b, tab, Grp<32>(lane), now);
    }
  }
}

// hand-off threshold: 1/16 of the batch, at most kTailBelow (a small batch would otherwise run entirely in the
// tail kernel, which is ~3x less efficient per iteration than the sweeps)
#ifndef MPCV_TAIL_SHIFT
#define MPCV_TAIL_SHIFT 4
#endif
// end of an iteration sweep: swap the lists and tell the WHILE node whether anyone is left
static __global__ void ph_flip_kernel(PhaseCtrl* ctrl, cudaGraphConditionalHandle handle, int use_handle) {
  const int in = ctrl->sweep & 1, out = in ^ 1;
  if (ph_repack_wanted(ctrl, ctrl->n_act[out])) {      // commit this sweep's repack
    ctrl->cur ^= 1;
    ctrl->slots = ctrl->n_act[out];
    ctrl->repacks += 1;
  }
  ctrl->n_act[in] = 0;
  ctrl->n_retry = 0;
  ctrl->n_slow = 0;
  ctrl->sweep += 1;
  ctrl->sweeps_total += 1;
  ctrl->sweeps_cum += 1;
  if (use_handle) cudaGraphSetConditional(handle, ctrl->n_act[out] > ctrl->tail_below ? 1u : 0u);
}

// start of a solve on pipe j of K: the pipe's share of the queue (contiguous, a multiple of 32 entries so that every
// share starts on a slab block).  The number of entries is the host's B or, for closed loops over the live
// scenarios, a device counter.
__host__ __device__ inline long ph_share_of(long cnt, int K) { return ((cnt + K - 1) / K + 31) / 32 * 32; }
static __global__ void ph_begin_kernel(PhaseCtrl* ctrl, SolveIO* dst, const SolveIO io, long B_host, int j, int K,
                                       int tail_cap, int tail_shift, int* res_queue_head) {
  *dst = io;
  if (res_queue_head) *res_queue_head = 0;     // queue of the resident tail kernel (mpcv_resident.cuh)
  const long cnt = io.count ? (long)*io.count : B_host;
  const long share = ph_share_of(cnt, K), e0 = (long)j * share;
  long nb = cnt - e0;
  nb = nb < 0 ? 0 : (nb > share ? share : nb);
  const int B = (int)nb;
  const int t = B >> tail_shift;
  ctrl->row0 = (int)e0;
  ctrl->B = B;
  ctrl->tail_below = t < 32 ? 32 : (t > tail_cap ? tail_cap : t);
  ctrl->n_act[0] = B; ctrl->n_act[1] = 0; ctrl->sweep = 0; ctrl->sweeps_total = 0;
  ctrl->n_retry = 0; ctrl->n_slow = 0;
  ctrl->cur = 0; ctrl->slots = B; ctrl->repacks = 0;
}
// ---------------------------------------------------------------------------------------
// closed loop over the phase pipeline (the scripts' MPC loop, Casadi/multiple_shooting_casadi.py:224-298,
// Trajectory_tracking.py:101-118): per MPC step  lp_prepare -> [solve graph] -> lp_apply, stream-ordered.
// Same semantics as closed_loop_problem() (mpcv_driver.cuh), with the per-step solve batch-wide.
// ---------------------------------------------------------------------------------------
struct LoopBufs {
  double* x0;       // [B, n]   guess of the next solve
  double* p;        // [B, n_p] parameter vector of the next solve
  double* x;        // [B, n]   solution of the last solve
  double* f;        // [B]
  double* state;    // [B, nx]  plant state
  int* status;      // [B]
  int* iters;       // [B]
  int* active;      // [B]      1 while the problem's loop is running
  int* steps;       // [B]
  int* iters_total; // [B]
  int* worst;       // [B]
  double* xctrl;    // [B, nx]  the controller's x0 (the plant state unless MPCV_LOOP_X0_FROM_PREDICTION)
  int* index;       // [B]      live scenarios of the current step (queue of the solve)
  int* count;       // [1]      their number
  long long* t0;    // [1]      device-timer stamp of the current step's start
};

template <class Model>
__global__ void lp_begin_kernel(const Layout L, const LoopIO io, const LoopBufs lb, long B) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* os = io.out_states + b * (long)(io.n_steps + 1) * NX;
  for (int i = 0; i < NX; ++i) { const double v = io.x_init[b * NX + i]; lb.state[b * NX + i] = v; lb.xctrl[b * NX + i] = v; os[i] = v; }
  if (b == 0) *lb.count = 0;
  // first guess: X_k = state, U = 0 (repmat(state_init) of MS:213); the scripts' own w0 = 0 in reference mode
  double* g = lb.x0 + b * L.n;
  for (int i = 0; i < L.n; ++i) g[i] = 0.0;
  if (io.warm_mode != MPCV_WARM_REFERENCE)
    for (int k = 0; k <= L.N; ++k)
      for (int i = 0; i < NX; ++i) g[k * NZ + i] = io.x_init[b * NX + i];
  lb.active[b] = 1; lb.steps[b] = 0; lb.iters_total[b] = 0; lb.worst[b] = 0;
}

template <class Model>
__global__ void lp_prepare_kernel(const Layout L, const LoopIO io, const LoopBufs lb, long B, int t) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, NPG = Model::NPG, NPS = Model::NPS;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && io.out_step_ns) *lb.t0 = ph_globaltimer();
  bool live = false;
  if (b < B) {
    const int np = NX + NPG + L.N * NPS;
    double* pr = lb.p + b * np;
    const double* st = lb.state + b * NX;
    // model parameters: constant, or those of step t (LTV: Trjectory_tracking_le_LTV.py:126-143)
    const double* pg = io.pglob_traj ? io.pglob_traj + (b * (long)io.n_steps + t) * NPG : io.pglob + b * NPG;
    for (int i = 0; i < NPG; ++i) pr[NX + i] = pg[i];
    live = lb.active[b] != 0;
    if (live && io.stop_radius > 0.0 && NPG >= NX) {
      double d2 = 0.0;
      for (int i = 0; i < NX; ++i) { const double e = st[i] - pr[NX + i]; d2 += e * e; }
      if (!(sqrt(d2) > io.stop_radius)) { lb.active[b] = 0; live = false; }   // while norm_2(state-target) > 1e-1 (MS:226)
    }
    if (live) {
      const double* xc = lb.xctrl + b * NX;
      for (int i = 0; i < NX; ++i) pr[i] = xc[i];
      if (NPS > 0) {
        // horizon window p[t..t+N) of this scenario's reference trajectory, or the step's own window table
        const double* src = (io.flags & MPCV_LOOP_PTRAJ_WINDOWS)
                                ? io.ptraj + (b * (long)io.n_steps + t) * (long)L.N * NPS
                                : io.ptraj + (b * (long)(io.n_steps + L.N) + t) * NPS;
        for (int i = 0; i < L.N * NPS; ++i) pr[NX + NPG + i] = src[i];
      }
      if (io.warm_mode == MPCV_WARM_COLD) {
        double* g = lb.x0 + b * L.n;
        for (int i = 0; i < L.n; ++i) g[i] = 0.0;
        for (int k = 0; k <= L.N; ++k)
          for (int i = 0; i < NX; ++i) g[k * NZ + i] = xc[i];
      }
    }
  }
  // the queue of this step's solve: scenarios whose loop is still running (finished ones are not solved again)
  ph_append(live, (int)b, lb.index, lb.count);
}

template <class Model>
__global__ void lp_apply_kernel(const Params P, const Layout L, const LoopIO io, const LoopBufs lb, long B, int t) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, NPG = Model::NPG, NPS = Model::NPS;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) {
    if (io.out_step_ns) io.out_step_ns[t] = ph_globaltimer() - *lb.t0;   // `times` of single_shooting_v1.py:209-212
    *lb.count = 0;                                                        // the next step builds its own queue
  }
  if (b >= B || !lb.active[b]) return;
  const int N = L.N, np = NX + NPG + N * NPS;
  const double* pr = lb.p + b * np;
  const double* x = lb.x + b * L.n;       // projected solution (honor_original_bounds)
  double* g = lb.x0 + b * L.n;
  double* st = lb.state + b * NX;
  double state[NX], u0[NU], xn[NX], q;
  for (int i = 0; i < NX; ++i) state[i] = st[i];
  for (int i = 0; i < NU; ++i) u0[i] = x[NX + i];
  // plant step with the same discretisation: state = F(p, u0) (MS:273)
  Model::val(P, state, u0, pr + NX, pr + NX + NPG, xn, &q);
  // `uprev` is never updated by the reference (Inverted_pendulum/...:64): replay on request
  if (Model::HAS_UPREV && io.warm_mode == MPCV_WARM_REFERENCE) xn[NX - 1] = io.x_init[b * NX + NX - 1];
  double* os = io.out_states + b * (long)(io.n_steps + 1) * NX;
  double* oc = io.out_controls + b * (long)io.n_steps * NU;
  for (int i = 0; i < NU; ++i) oc[t * NU + i] = u0[i];
  for (int i = 0; i < NX; ++i) { os[(t + 1) * NX + i] = xn[i]; st[i] = xn[i]; }
  // the next solve starts from the plant state, or from the solver's own prediction x_1
  // (solver.fixvar("x",0,solver.var["x",1]), Trajectory_tracking.py:111-112)
  double* xc = lb.xctrl + b * NX;
  for (int i = 0; i < NX; ++i) xc[i] = (io.flags & MPCV_LOOP_X0_FROM_PREDICTION) ? x[NZ + i] : xn[i];
  if (Model::HAS_UPREV && io.warm_mode == MPCV_WARM_REFERENCE) xc[NX - 1] = io.x_init[b * NX + NX - 1];
  if (io.out_horizons) {
    // predicted horizon of this solve (cat_states of single_shooting_v1.py:185-188)
    double* oh = io.out_horizons + (b * (long)io.n_steps + t) * (long)(N + 1) * NX;
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) oh[k * NX + i] = x[k * NZ + i];
  }
  lb.steps[b] += 1;
  lb.iters_total[b] += lb.iters[b];
  if (lb.status[b] != 0 && lb.worst[b] == 0) lb.worst[b] = lb.status[b];
  // next guess
  if (io.warm_mode == MPCV_WARM_SHIFT) {
    for (int k = 0; k < N; ++k) {
      for (int i = 0; i < NX; ++i) g[k * NZ + i] = x[(k + 1) * NZ + i];
      const int ks = (k + 1 < N) ? k + 1 : N - 1;            // the last control is kept
      for (int i = 0; i < NU; ++i) g[k * NZ + NX + i] = x[ks * NZ + NX + i];
    }
    for (int i = 0; i < NX; ++i) g[N * NZ + i] = x[N * NZ + i];
  } else if (io.warm_mode == MPCV_WARM_REFERENCE) {
    int q2 = 0;   // MS:279-287  w0 = [vec(shifted X); vec(shifted U)]
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) { const int ks = (k + 1 <= N) ? k + 1 : N; g[q2++] = x[ks * NZ + i]; }
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i) { const int ks = (k + 1 < N) ? k + 1 : N - 1; g[q2++] = x[ks * NZ + NX + i]; }
  } else {
    for (int i = 0; i < L.n; ++i) g[i] = x[i];
  }
}

template <class Model>
__global__ void lp_end_kernel(const LoopIO io, const LoopBufs lb, long B) {
  constexpr int NX = Model::NX, NU = Model::NU;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int steps = lb.steps[b];
  if (io.out_steps) io.out_steps[b] = steps;
  if (io.out_iters) io.out_iters[b] = lb.iters_total[b];
  if (io.out_status) io.out_status[b] = lb.worst[b];
  double* os = io.out_states + b * (long)(io.n_steps + 1) * NX;
  double* oc = io.out_controls + b * (long)io.n_steps * NU;
  for (int t = steps; t < io.n_steps; ++t) {
    for (int i = 0; i < NX; ++i) os[(t + 1) * NX + i] = lb.state[b * NX + i];
    for (int i = 0; i < NU; ++i) oc[t * NU + i] = 0.0;
  }
}
#endif  // __CUDACC__

}  // namespace mpcv
