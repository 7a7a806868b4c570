// mpcv_resident.cuh — CTA-resident, phase-synchronous solve (layout MPCV_LAYOUT_RESIDENT).
//
// Replaces the same reference call as mpcv_ipm.cuh / mpcv_phase.cuh (`sol = solver(x0=,lbx=,ubx=,lbg=,ubg=,p=)`,
// Casadi/multiple_shooting_casadi.py:235-242; `solver.solve()` of the MPCTools scripts) with the SAME phase functions
// of Ipm<> — but the per-problem workspace never leaves the SM:
//   * a persistent CTA owns `slots` problems; each slot is one row of shared memory (Layout.total doubles, odd
//     stride), so only x0 / p in and x / f / status out touch HBM (the slab pipeline of mpcv_phase.cuh streams the
//     8.6 KB workspace of every problem through HBM ~3x per iteration sweep: 634x the algorithmic bytes);
//   * the CTA walks the phases of one interior-point iteration in lock step (`__syncthreads()` between phases), each
//     phase with the thread mapping that fits it — lane groups for the vector phases (pre / post / accept), a few
//     lanes per slot for the lane-parallel Riccati factorisation, one thread per (slot, interval) for the derivative
//     and trial sweeps — so all warps of a CTA sit in one small phase body at a time (the one-kernel solve thrashes
//     the 32 KB instruction cache), without 290 launches per solve;
//   * continuous batching: a slot whose problem converged is refilled at once from a global queue (one atomicAdd per
//     problem), so no lane idles on finished problems, there is no repack and no straggler tail: a problem that needs
//     119 iterations just keeps its slot while its neighbours turn over;
//   * several CTAs per SM run out of step with each other: the latency-bound recursion of one overlaps the FP64-bound
//     derivative sweep of another.
// Slot life cycle: EMPTY -(pre: refill, start)-> NEW -(trial phase: derivative sweep at df = 1; accept phase:
// objective scaling + least-squares multipliers)-> RUN -(pre: converged -> export)-> EMPTY ... ; DONE once the queue is dry.
// The rare slow path (backtracking / second-order correction, 1-2 % of the iterations) runs one warp per slot
// beside the derivative sweep of the others.
#pragma once

#include "mpcv_phase.cuh"

namespace mpcv {

#if defined(__CUDACC__)

#ifndef MPCV_RES_THREADS
#define MPCV_RES_THREADS 192   /* 2 CTAs x 192 threads leave 168 registers per thread: the derivative sweep needs 162 */
#endif
#ifndef MPCV_RES_MINB
#define MPCV_RES_MINB 2      /* CTAs per SM: out of step with each other, they fill each other's latency-bound phases */
#endif
#ifndef MPCV_RES_GL
#define MPCV_RES_GL 8        /* lanes per slot in the vector phases */
#endif
#ifndef MPCV_RES_RL
#define MPCV_RES_RL 4        /* lanes per slot in the Riccati factorisation (NX columns + the vector part) */
#endif
constexpr int kResThreads = MPCV_RES_THREADS;
constexpr int kResMinB = MPCV_RES_MINB;
constexpr int kResGL = MPCV_RES_GL;
constexpr int kResRL = MPCV_RES_RL;
constexpr int kResMaxSlots = 64;

struct ResCtrl {
  int next;        // head of the problem queue
  int pad[3];
};

struct ResArgs {
  Params P;
  Layout L;
  SolveIO io;
  ResCtrl* ctrl;
  const int* count;   // device: problems of this call (null = B)
  const int* index;   // device: I/O row of queue entry e (null = e): closed loops solve only the live scenarios
  long B;
  int slots;          // problems resident per CTA
  int stride;         // doubles per slot row
  // TAIL mode (the stragglers of the phase pipeline): the queue is the pipe's active list, a refill stages the
  // problem's workspace in from the slab with its solve state, the I/O pointers are the pipe's device copy
  const PhaseCtrl* pctrl;
  const SolveIO* io_dev;
  double* slab[2];
  const int* act[2];
};

enum { RS_EMPTY = 0, RS_NEW = 1, RS_RUN = 2, RS_DONE = 3 };

template <class Model>
struct Resident {
  using WS = WsShared;
  using PhG = Phase<Model, WS, kResGL>;
  using Ph1 = Phase<Model, WS, 1>;
  using IpmG = Ipm<Model, false, kResGL, WS>;
  using IpmR = Ipm<Model, false, kResRL, WS>;
  static constexpr int NX = Model::NX, NH = Model::NX + Model::NPG;

  // refill: load x0 / p of problem `b`, push into the interior, z0, lam0          (kResGL lanes per slot)
  __device__ static void load_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                   const BndEntry* tab, Grp<kResGL> g, long long t0) {
    const int np = NH + L.N * Model::NPS;
    for (int i = g.lane; i < L.n; i += kResGL) ws[L.w + i] = io.x0 ? io.x0[b * L.n + i] : 0.0;
    for (int i = g.lane; i < np; i += kResGL) ws[L.par + i] = io.p[b * np + i];
    if (g.lane == 0) {
      ws[L.st + 14] = PhG::long_as_double(t0);
      ws[L.st + 15] = (double)b;
    }
    g.sync();
    IpmG ipm = PhG::template make_ipm<IpmG>(P, L, ws, g, io, tab);
    ipm.start();
    ipm.save_state(kRunning);
    g.sync();
  }

  // Riccati factorisation with IPOPT's delta_w schedule, then the forward sweep         (kResRL lanes per slot)
  __device__ static bool factor_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, long b,
                                     const BndEntry* tab, Grp<kResRL> g, long long now) {
    IpmR ipm = PhG::template make_ipm<IpmR>(P, L, ws, g, io, tab);
    ipm.delta_w_last = ws[L.st + 6];
    double dw = 0.0;
    bool ok = ipm.template riccati_factor_x<true>(0.0, false, 0, L.c);
    while (!ok) {
      dw = ipm.next_delta_w(dw);
      if (dw > 1e20) break;
      ok = ipm.template riccati_factor_x<true>(dw, false, 0, L.c);
    }
    if (!ok) {
      ipm.load_state();
      PhG::finish(ipm, MPCV_ERROR_IN_STEP_COMPUTATION, io, b, now);
      g.sync();
      return false;
    }
    if (dw > 0.0 && g.lane == 0) ws[L.st + 6] = dw;
    ipm.riccati_forward(L.c);
    return true;
  }

  // objective scaling + least-squares multipliers of a fresh problem                   (kResGL lanes per slot)
  __device__ static void init2_body(const Params& P, const Layout& L, WS ws, const SolveIO& io, const BndEntry* tab,
                                    Grp<kResGL> g) {
    IpmG ipm = PhG::template make_ipm<IpmG>(P, L, ws, g, io, tab);
    ipm.load_state();
    ipm.f_curr = ipm.sum_stage_costs();
    ipm.init_scaling_and_multipliers();
    ipm.save_state(kRunning);
    g.sync();
  }
};

// the rare slow path of one slot on one warp: backtracking / second-order correction, then the derivative sweep at
// the accepted point.  Out of line: it is large and almost never runs.
template <class Model>
__device__ __noinline__ void res_slow_path(const Params& P, const Layout& L, double* row, const SolveIO& io,
                                           const BndEntry* tab, int lane, int* state_out) {
  const WsShared ws{row};
  const long b = (long)ws[L.st + 15];
  Phase<Model, WsShared, 32>::slow_body(P, L, ws, io, b, tab, Grp<32>(lane), io.ns ? ph_globaltimer() : 0);
  __syncwarp();
  if (Phase<Model, WsShared, 1>::running(L, ws)) {
    for (int k = lane; k < L.N; k += 32) Phase<Model, WsShared, 1>::der_body(P, L, ws, io, k, true, tab);
  } else if (lane == 0) {
    *state_out = RS_EMPTY;
  }
  __syncwarp();
}

template <class Model, bool TAIL>
__global__ void __launch_bounds__(kResThreads, kResMinB) res_solve_kernel(const __grid_constant__ ResArgs a) {
  using R = Resident<Model>;
  using WS = WsShared;
  extern __shared__ __align__(16) unsigned char ph_smem[];
  __shared__ int sstate[kResMaxSlots], sslow[kResMaxSlots], slowlist[kResMaxSlots];
  __shared__ int n_slow, slow_next;
  const Layout& L = a.L;
  const Params& P = a.P;
  const SolveIO io = TAIL ? *a.io_dev : a.io;
  const int tid = threadIdx.x, lane = tid & 31;
  // TAIL: the stragglers the sweeps left in the pipe's active list, their workspaces in the current slab
  const int tin = TAIL ? (a.pctrl->sweep & 1) : 0;
  const long B = TAIL ? (long)a.pctrl->n_act[tin] : (a.count ? (long)*a.count : a.B);
  if (TAIL && (long)blockIdx.x >= B) return;
  const double* const tslab = TAIL ? a.slab[a.pctrl->cur] : nullptr;
  const int* const tlist = TAIL ? a.act[tin] : nullptr;
  // few problems: spread them over the CTAs instead of filling the first ones
  int S = a.slots;
  if (TAIL) { const long per = (B + gridDim.x - 1) / gridDim.x; if (per < S) S = (int)(per < 1 ? 1 : per); }
  const BndEntry* tab = ph_bounds_table<Model>(P, L, io);
  double* const rows = reinterpret_cast<double*>(ph_smem + ph_rows_offset(L));
  auto row = [&](int s) { return WS{rows + (long)s * a.stride}; };
  for (int s = tid; s < kResMaxSlots; s += blockDim.x) { sstate[s] = s < S ? RS_EMPTY : RS_DONE; sslow[s] = 0; }
  if (tid == 0) { n_slow = 0; slow_next = 0; }
  __syncthreads();
  constexpr int GL = kResGL, RL = kResRL;
  const int ggrp = tid / GL, ngg = kResThreads / GL;
  const int rgrp = tid / RL, nrg = kResThreads / RL;
  const Grp<GL> gg(lane);
  const Grp<RL> gr(lane);

  for (;;) {
    // ---- pre: convergence test (export + refill), barrier update, Sigma / barrier gradient ----
    int alive = 0;
    if (tid == 0) { n_slow = 0; slow_next = 0; }
    for (int s0 = 0; s0 < S; s0 += ngg) {
      const int s = s0 + ggrp;
      if (s < S) {
        int state = sstate[s];
        const WS ws = row(s);
        if (state == RS_RUN) {
          const long b = (long)ws[L.st + 15];
          if (!R::PhG::pre_body(P, L, ws, io, b, tab, gg, io.ns ? ph_globaltimer() : 0)) state = RS_EMPTY;
        }
        if (state == RS_EMPTY) {
          int e = 0;
          if (gg.lane == 0) e = atomicAdd(&a.ctrl->next, 1);
          e = gg.bcast(e);
          if (e < B) {
            gg.sync();
            if (TAIL) {
              // the problem comes with its iterate, derivatives and solve state: copy its workspace in and go on
              const WsStrided src = WsStrided::of(const_cast<double*>(tslab), L.total, tlist[e]);
              for (int i = gg.lane; i < L.total; i += GL) ph_cp_async8(&ws[i], &src[i]);
              ph_cp_async_wait();
              gg.sync();
              state = Phase<Model, WS, 1>::running(L, ws) ? RS_RUN : RS_EMPTY;
              if (state == RS_RUN) {
                const long b = (long)ws[L.st + 15];
                if (!R::PhG::pre_body(P, L, ws, io, b, tab, gg, io.ns ? ph_globaltimer() : 0)) state = RS_EMPTY;
              }
            } else {
              const long b = a.index ? (long)a.index[e] : (long)e;
              R::load_body(P, L, ws, io, b, tab, gg, io.ns ? ph_globaltimer() : 0);
              state = RS_NEW;
            }
          } else {
            state = RS_DONE;
          }
        }
        if (gg.lane == 0) { sstate[s] = state; sslow[s] = 0; }
        alive |= state != RS_DONE;
      }
    }
    if (!__syncthreads_or(alive)) break;

    // ---- factor: Riccati factorisation (delta_w schedule), vector recursions ----
    for (int s0 = 0; s0 < S; s0 += nrg) {
      const int s = s0 + rgrp;
      if (s < S && sstate[s] == RS_RUN) {
        const WS ws = row(s);
        const long b = (long)ws[L.st + 15];
        if (!R::factor_body(P, L, ws, io, b, tab, gr, io.ns ? ph_globaltimer() : 0) && gr.lane == 0) sstate[s] = RS_EMPTY;
      }
    }
    __syncthreads();

    // ---- post: fraction-to-the-boundary step, merit-function terms ----
    for (int s0 = 0; s0 < S; s0 += ngg) {
      const int s = s0 + ggrp;
      if (s < S && sstate[s] == RS_RUN) R::PhG::post_body(P, L, row(s), io, tab, gg);
    }
    __syncthreads();

    // ---- trial: first line-search trial point; fresh problems: derivative sweep at df = 1 ----
    {
      const int items = S * L.N;
      for (int it = tid; it < items; it += kResThreads) {
        const int s = it / L.N, k = it - s * L.N;
        const int state = sstate[s];
        if (state == RS_RUN) R::Ph1::trial_body(P, L, row(s), io, k, tab);
        else if (state == RS_NEW) R::Ph1::der_body(P, L, row(s), io, k, false, tab);
      }
    }
    __syncthreads();

    // ---- accept: filter test of the full step, dual step, update; fresh problems: scaling + multipliers ----
    for (int s0 = 0; s0 < S; s0 += ngg) {
      const int s = s0 + ggrp;
      const int state = s < S ? sstate[s] : RS_DONE;
      const WS ws = row(s < S ? s : 0);
      typename R::IpmG ipm = R::PhG::template make_ipm<typename R::IpmG>(P, L, ws, gg, io, tab);
      typename R::IpmG::LsFirst r;
      r.ok = false;
      if (state == RS_RUN) {
        ipm.load_state();
        r = ipm.line_search_first_decide();
        if (r.ok) ipm.ls_filter_augment(ipm.ls_alpha_max, r.phi_t, r.pw);
      }
      __syncwarp();
      if (state == RS_RUN) {
        if (r.ok) {
          ipm.ls_take_step(ipm.ls_alpha_max);
          ipm.save_state(kRunning);
        } else if (gg.lane == 0) {
          sslow[s] = 1;
          slowlist[atomicAdd(&n_slow, 1)] = s;
        }
      } else if (state == RS_NEW) {
        R::init2_body(P, L, ws, io, tab, gg);
        if (gg.lane == 0) sstate[s] = RS_RUN;
      }
    }
    __syncthreads();

    // ---- der: derivative sweep at the new iterate; rejected full steps: slow path, one warp per slot ----
    {
      const int items = S * L.N;
      for (int it = tid; it < items; it += kResThreads) {
        const int s = it / L.N, k = it - s * L.N;
        if (sstate[s] == RS_RUN && !sslow[s]) R::Ph1::der_body(P, L, row(s), io, k, true, tab);
      }
      __syncwarp();
      if (n_slow > 0) {
        for (;;) {
          int i = 0;
          if (lane == 0) i = atomicAdd(&slow_next, 1);
          i = __shfl_sync(0xffffffffu, i, 0);
          if (i >= n_slow) break;
          const int s = slowlist[i];
          res_slow_path<Model>(P, L, rows + (long)s * a.stride, io, tab, lane, &sstate[s]);
        }
      }
    }
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace mpcv
