// mpcv_driver.cuh — per-problem entry points shared by the thread- and warp-layout kernels:
//   solve_problem       : one `solver(x0, lbx, ubx, lbg, ubg, p)` call
//   closed_loop_problem : the scripts' MPC loop (solve -> apply u0 -> plant step -> shift)
//       Casadi/multiple_shooting_casadi.py:224-298, single_shooting_v1.py:164-214 (+17-27),
//       single_shooting_v2.py:201-266; MPCTools loops (Trajectory_tracking.py:101-118).
#pragma once

#include "mpcv_ipm.cuh"

namespace mpcv {

// Default parameter stager: plain (coalesced across lanes) loads.  The warp-layout kernel
// substitutes a TMA bulk-copy stager (cp.async.bulk + mbarrier) for shared-memory workspaces.
struct PlainStager {
  // copy `count` doubles src -> ws[off ..]; returns the element shift applied to `off`
  template <class WS, int LANES>
  MPCV_D int load(const WS& ws, int off, const double* src, int count, const Grp<LANES>& g) const {
    for (int i = g.lane; i < count; i += LANES) ws[off + i] = src[i];
    return 0;
  }
};

// batch-leading row-major I/O pointers of one call (any output may be null)
struct SolveIO {
  const double* x0;    // [B, n]  (null = zeros)
  const double* lbx;   // [n]
  const double* ubx;   // [n]
  const double* p;     // [B, n_p]
  double* x;           // [B, n]
  double* f;           // [B]
  double* g;           // [B, n_g]
  double* lam_g;       // [B, n_g]
  double* lam_x;       // [B, n]
  int* status;         // [B]
  int* iters;          // [B]
  long long* ns;       // [B] per-problem latency (globaltimer), optional
  long bstride = 0;            // 0: lbx / ubx are [n], shared by the batch; n: per-problem rows [B, n]
  const int* index = nullptr;  // device, optional: I/O row of queue entry e (closed loops solve only live scenarios)
  const int* count = nullptr;  // device, optional: number of queue entries (overrides B)
};

template <class Model, bool SINGLE, int LANES, class WS>
MPCV_D void export_solution(const Ipm<Model, SINGLE, LANES, WS>& ipm, const SolveInfo& info,
                            const SolveIO& io, long b) {
  constexpr int NX = Model::NX;
  const Layout& L = ipm.L;
  const WS& ws = ipm.ws;
  const int ng = NX * (L.N + 1);
  const int lane = ipm.g.lane;
  if (io.x) for (int i = lane; i < L.n; i += LANES) io.x[b * L.n + i] = ipm.projected(i);
  if (io.lam_x) for (int i = lane; i < L.n; i += LANES) io.lam_x[b * L.n + i] = (ws[L.zu + i] - ws[L.zl + i]) / info.df;
  if (io.g) for (int i = lane; i < ng; i += LANES) io.g[b * ng + i] = SINGLE ? ws[L.xs + i] : ws[L.c + i];
  if (io.lam_g) for (int i = lane; i < ng; i += LANES) io.lam_g[b * ng + i] = ws[L.lam + i] / info.df;
  if (lane == 0) {
    if (io.f) io.f[b] = info.f;
    if (io.status) io.status[b] = info.status;
    if (io.iters) io.iters[b] = info.iters;
  }
}

template <class Model, bool SINGLE, int LANES, class WS, class Stager = PlainStager>
MPCV_D void solve_problem(const Params& P, const Layout& L, WS ws, Grp<LANES> g, const SolveIO& io, long b,
                          const Stager& stager = Stager()) {
  constexpr int NH = Model::NX + Model::NPG;
  const int np = NH + L.N * Model::NPS;
  for (int i = g.lane; i < L.n; i += LANES) ws[L.w + i] = io.x0 ? io.x0[b * L.n + i] : 0.0;
  for (int i = g.lane; i < NH; i += LANES) ws[L.par + i] = io.p[b * np + i];
  Ipm<Model, SINGLE, LANES, WS> ipm(P, L, ws, g, io.lbx + b * io.bstride, io.ubx + b * io.bstride);
  if (Model::NPS > 0) ipm.ps_base += stager.load(ws, L.par + NH, io.p + b * np + NH, L.N * Model::NPS, g);
  g.sync();
  const SolveInfo info = ipm.solve();
  export_solution(ipm, info, io, b);
}

struct LoopIO {
  const double* x_init;   // [B, nx]
  const double* pglob;    // [B, npg]
  const double* ptraj;    // [B, n_steps+N, nps]
  const double* lbx;
  const double* ubx;
  double* out_states;     // [B, n_steps+1, nx]
  double* out_controls;   // [B, n_steps, nu]
  int* out_steps;         // [B]
  int* out_iters;         // [B]
  int* out_status;        // [B]
  int n_steps, warm_mode;
  double stop_radius;
  const double* pglob_traj = nullptr;   // [B, n_steps, npg] per-step model parameters (LTV), overrides pglob
  double* out_horizons = nullptr;       // [B, n_steps, N+1, nx] predicted states of every solve
  long long* out_step_ns = nullptr;     // [n_steps] device-timer duration of every MPC step (phased / resident layouts)
  int flags = 0;                        // MPCV_LOOP_*
};

template <class Model, bool SINGLE, int LANES, class WS, class Stager = PlainStager>
MPCV_D void closed_loop_problem(const Params& P, const Layout& L, WS ws, Grp<LANES> g, const LoopIO& io, long b,
                                const Stager& stager = Stager()) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  const int N = L.N, lane = g.lane;
  double state[NX], xctrl[NX];    // plant state; the controller's x0 (the same unless MPCV_LOOP_X0_FROM_PREDICTION)
#pragma unroll
  for (int i = 0; i < NX; ++i) xctrl[i] = state[i] = io.x_init[b * NX + i];
  if (io.pglob) for (int i = lane; i < Model::NPG; i += LANES) ws[L.par + NX + i] = io.pglob[b * Model::NPG + i];
  double* os = io.out_states + b * (long)(io.n_steps + 1) * NX;
  double* oc = io.out_controls + b * (long)io.n_steps * NU;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) os[i] = state[i];
  }
  // first guess: X_k = state, U = 0 (repmat(state_init) of MS:213); the scripts' own w0 = 0 in
  // reference mode (MS:134,149,169)
  for (int i = lane; i < L.n; i += LANES) ws[L.w + i] = 0.0;
  g.sync();
  if (!SINGLE && io.warm_mode != MPCV_WARM_REFERENCE) {
    for (int k = lane; k <= N; k += LANES) {
#pragma unroll
      for (int i = 0; i < NX; ++i) ws[L.w + k * NZ + i] = state[i];
    }
  }
  g.sync();
  int steps = 0, iters_total = 0, worst = 0;
  Ipm<Model, SINGLE, LANES, WS> ipm(P, L, ws, g, io.lbx, io.ubx);
  for (int t = 0; t < io.n_steps; ++t) {
    if (io.pglob_traj) {
      // LTV: the model of step t (Trjectory_tracking_le_LTV.py:126-143 re-discretises Ac(c[t]) every step)
      g.sync();
      for (int i = lane; i < Model::NPG; i += LANES)
        ws[L.par + NX + i] = io.pglob_traj[(b * (long)io.n_steps + t) * Model::NPG + i];
      g.sync();
    }
    if (io.stop_radius > 0.0 && Model::NPG >= NX) {
      double d2 = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) { const double e = state[i] - ws[L.par + NX + i]; d2 += e * e; }
      if (!(sqrt(d2) > io.stop_radius)) break;   // while norm_2(state-target) > 1e-1 (MS:226)
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) ws[L.par + i] = xctrl[i];
    }
    if (Model::NPS > 0) {
      // horizon window p[t..t+N) of this scenario's reference trajectory (overlapping slices of
      // one array: nothing is materialised per step), or the step's own window table
      const double* src = (io.flags & MPCV_LOOP_PTRAJ_WINDOWS)
                              ? io.ptraj + (b * (long)io.n_steps + t) * (long)N * Model::NPS
                              : io.ptraj + (b * (long)(io.n_steps + N) + t) * Model::NPS;
      ipm.ps_base = L.par + NX + Model::NPG + stager.load(ws, L.par + NX + Model::NPG, src, N * Model::NPS, g);
    }
    if (io.warm_mode == MPCV_WARM_COLD) {
      g.sync();
      for (int i = lane; i < L.n; i += LANES) ws[L.w + i] = 0.0;
      g.sync();
      if (!SINGLE) {
        for (int k = lane; k <= N; k += LANES) {
#pragma unroll
          for (int i = 0; i < NX; ++i) ws[L.w + k * NZ + i] = xctrl[i];
        }
      }
    }
    g.sync();
    const SolveInfo info = ipm.solve();
    iters_total += info.iters;
    if (info.status != 0 && worst == 0) worst = info.status;
    // projected solution back into w (honor_original_bounds) so that u0 and the guess use sol['x']
    for (int i = lane; i < L.n; i += LANES) ws[L.w + i] = ipm.projected(i);
    g.sync();
    double u0[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) u0[i] = ws[L.w + ipm.iu(0, i)];
    if (io.out_horizons) {
      // predicted horizon of this solve (cat_states of single_shooting_v1.py:185-188)
      double* oh = io.out_horizons + (b * (long)io.n_steps + t) * (long)(N + 1) * NX;
      for (int q = lane; q < (N + 1) * NX; q += LANES)
        oh[q] = SINGLE ? ws[L.xs + q] : ws[L.w + (q / NX) * NZ + q % NX];
    }
    double xpred[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xpred[i] = SINGLE ? ws[L.xs + NX + i] : ws[L.w + NZ + i];
    // plant step with the same discretisation: state = F(p, u0) (MS:273); Euler in SSv1:17-19
    {
      double xn[NX], q;
      Model::val(P, state, u0, ipm.pg(), ipm.ps(0), xn, &q);
#pragma unroll
      for (int i = 0; i < NX; ++i) state[i] = xn[i];
      // `uprev` is never updated by the reference (Inverted_pendulum/...:64): replay on request
      if (Model::HAS_UPREV && io.warm_mode == MPCV_WARM_REFERENCE) state[NX - 1] = io.x_init[b * NX + NX - 1];
      // next x0: the plant state, or the solver's own prediction x_1 (fixvar("x",0,var["x",1]),
      // Trajectory_tracking.py:111-112)
#pragma unroll
      for (int i = 0; i < NX; ++i) xctrl[i] = (io.flags & MPCV_LOOP_X0_FROM_PREDICTION) ? xpred[i] : state[i];
      if (Model::HAS_UPREV && io.warm_mode == MPCV_WARM_REFERENCE) xctrl[NX - 1] = io.x_init[b * NX + NX - 1];
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) oc[t * NU + i] = u0[i];
#pragma unroll
      for (int i = 0; i < NX; ++i) os[(t + 1) * NX + i] = state[i];
    }
    ++steps;
    // next guess
    if (io.warm_mode == MPCV_WARM_SHIFT) {
      if (lane == 0) {
        if (SINGLE) {
          for (int k = 0; k + 1 < N; ++k)
            for (int i = 0; i < NU; ++i) ws[L.w + k * NU + i] = ws[L.w + (k + 1) * NU + i];
        } else {
          for (int k = 0; k < N; ++k) {
            for (int i = 0; i < NX; ++i) ws[L.w + k * NZ + i] = ws[L.w + (k + 1) * NZ + i];
            if (k + 1 < N)
              for (int i = 0; i < NU; ++i) ws[L.w + k * NZ + NX + i] = ws[L.w + (k + 1) * NZ + NX + i];
          }
        }
      }
    } else if (io.warm_mode == MPCV_WARM_REFERENCE) {
      // the scripts' own guess vectors (layout quirks included), staged through the step buffer
      if (lane == 0) {
        if (SINGLE) {
          const bool colmajor = Model::MODEL_ID == MPCV_MODEL_UNICYCLE_EULER_NODE;   // SSv1:173
          for (int k = 0; k < N; ++k)
            for (int i = 0; i < NU; ++i) {
              const int ks = (k + 1 < N) ? k + 1 : N - 1;
              ws[L.d + (colmajor ? i * N + k : k * NU + i)] = ws[L.w + ks * NU + i];
            }
        } else {
          int q = 0;   // MS:279-287  w0 = [vec(shifted X); vec(shifted U)]
          for (int k = 0; k <= N; ++k)
            for (int i = 0; i < NX; ++i) { const int ks = (k + 1 <= N) ? k + 1 : N; ws[L.d + q++] = ws[L.w + ks * NZ + i]; }
          for (int k = 0; k < N; ++k)
            for (int i = 0; i < NU; ++i) { const int ks = (k + 1 < N) ? k + 1 : N - 1; ws[L.d + q++] = ws[L.w + ks * NZ + NX + i]; }
        }
        for (int i = 0; i < L.n; ++i) ws[L.w + i] = ws[L.d + i];
      }
    }
    g.sync();
  }
  if (lane == 0) {
    if (io.out_steps) io.out_steps[b] = steps;
    if (io.out_iters) io.out_iters[b] = iters_total;
    if (io.out_status) io.out_status[b] = worst;
    for (int t = steps; t < io.n_steps; ++t) {
#pragma unroll
      for (int i = 0; i < NX; ++i) os[(t + 1) * NX + i] = state[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) oc[t * NU + i] = 0.0;
    }
  }
}

}  // namespace mpcv
