// mpcv_ipm.cuh — per-problem primal-dual interior-point solve with IPOPT semantics.
//
// Replaces the reference's `sol = solver(x0=,lbx=,ubx=,lbg=,ubg=,p=)` call
// (Casadi/multiple_shooting_casadi.py:235-242, single_shooting_v1.py:174-181,
// single_shooting_v2.py:212-219; `solver.solve()` in the MPCTools scripts).  IPOPT itself is a
// third-party binary; the semantics followed here are those of Waechter & Biegler (2006)
// with IPOPT 3.12 defaults: bound relaxation + push, z0 = 1, least-squares multiplier
// initialisation, monotone barrier update, fraction-to-boundary, filter line search with
// second-order correction, inertia correction by delta_w, the scaled E_0 <= tol test and
// the final projection into the original bounds.
//
// KKT solve per problem:
//   multiple shooting — block-tridiagonal Riccati recursion (backward sweep = adjoint sweep
//     over the horizon, forward sweep = linearised rollout); "inertia correct" <=> every
//     R + B'PB Cholesky succeeds.
//   single shooting   — costate (adjoint) sweep for the exact condensed Hessian, dense
//     Cholesky of H + Sigma + delta_w I.
//
// Execution shape: a group of LANES threads owns one problem.  LANES = 1 is the
// thread-per-problem layout (workspace: warp-blocked structure-of-arrays in HBM, see WsStrided,
// so a warp's accesses coalesce).  LANES = 32 is the warp-per-problem
// layout (workspace in shared memory, stage-parallel derivative / trial evaluation, shuffle
// reductions for the merit function, step norms and convergence tests; the sequential
// Riccati sweeps run on lane 0).
#pragma once

#include "mpcv_models.cuh"

namespace mpcv {

constexpr double kInfBound = 1e19;   // IPOPT nlp_lower_bound_inf / nlp_upper_bound_inf

// ---------------------------------------------------------------------------------------
// group-of-lanes primitives
// ---------------------------------------------------------------------------------------
template <int LANES>
struct Grp {
  int lane;
  unsigned mask;
#if defined(__CUDACC__)
  MPCV_D explicit Grp(int tid_in_warp) {
    lane = tid_in_warp % LANES;
    mask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (tid_in_warp - lane));
  }
  MPCV_D void sync() const { if (LANES > 1) __syncwarp(mask); }
  MPCV_D double sum(double v) const {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LANES);
    return v;
  }
  MPCV_D double max(double v) const {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, LANES));
    return v;
  }
  MPCV_D double min(double v) const {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o, LANES));
    return v;
  }
  MPCV_D int bcast(int v) const { return LANES > 1 ? __shfl_sync(mask, v, 0, LANES) : v; }
  // true when `ok` holds on every lane of the group (one vote instead of a shuffle tree)
  MPCV_D bool all(bool ok) const { return LANES > 1 ? (__ballot_sync(mask, ok) & mask) == mask : ok; }
#else
  explicit Grp(int) : lane(0), mask(1u) { static_assert(LANES == 1, "host harness is single-lane"); }
  void sync() const {}
  double sum(double v) const { return v; }
  double max(double v) const { return v; }
  double min(double v) const { return v; }
  int bcast(int v) const { return v; }
  bool all(bool ok) const { return ok; }
#endif
};

// Warp-blocked structure of arrays in HBM (thread layout and phase kernels): the 32 problems of
// block (b >> 5) keep element i in 32 consecutive doubles, so a warp reads or writes one 256-byte run
// per element AND element offsets are compile-time immediates (i * 256 bytes) off one per-thread
// base pointer — no index arithmetic per access.
struct WsStrided {
  double* base;
  MPCV_HD double& operator[](int i) const {
    double* p = base + (long)i * 32;
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isGlobal(p));   // plain ld.global / st.global instead of generic accesses
#endif
    return *p;
  }
  MPCV_HD WsStrided view(int off) const { return WsStrided{base + (long)off * 32}; }
  MPCV_HD static WsStrided of(double* slab, int total, long b) {
    double* p = slab + (b >> 5) * ((long)total * 32) + (b & 31);
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isGlobal(p));   // plain ld.global / st.global instead of generic accesses
#endif
    return WsStrided{p};
  }
};
// contiguous workspace view (shared memory / host harness)
struct WsDense {
  double* base;
  MPCV_HD double& operator[](int i) const { return base[i]; }
  MPCV_HD WsDense view(int off) const { return WsDense{base + off}; }
};
template <class WS>
struct WsView {
  WS ws;
  int off;
  MPCV_HD double operator[](int i) const { return ws[off + i]; }
};

// ---------------------------------------------------------------------------------------
// workspace layout (in doubles) for one problem
// ---------------------------------------------------------------------------------------
struct Layout {
  int n, m, N;
  int w, lam, zl, zu, d, lamp, grad, c, ct, ab, hw, ric, pp, par, xs, hred, gam, tmp;
  int sig, rb;   // Sigma_i and barrier-gradient r_i of the current iterate (filled once per iteration)
  int qs;        // per-stage interval costs (stage-parallel kernels hand them to the per-problem reduction)
  int st;        // persistent scalar state of the solve (phase-kernel pipeline), kStateSlots doubles
  int park;      // where the second-order correction parks the plain step (n + m doubles): the Hessian blocks, or a region
                 // of its own when the model keeps ONE copy of them for all stages (Model::LTI)
  int total;
};

// Relaxed bounds of one variable, precomputed once per kernel (the bounds are shared by the batch).
struct BndEntry { double lo, hi; int flags, pad; };   // flags: 1 = lower, 2 = upper, 4 = fixed
constexpr int kStateSlots = 56;    // 0..15 scalars, 16..31 filter phi, 32..39 scalars, 40..55 filter theta
constexpr int kSlotDwHint = 34;   // phase pipeline: delta_w found by the parallel probe (> 0) or -(last delta_w probed)
constexpr int kRunning = 1000;    // internal "not finished" status

template <class Model, bool SINGLE>
MPCV_HD Layout make_layout(int N) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  Layout L;
  L.N = N;
  L.n = SINGLE ? NU * N : NZ * N + NX;
  L.m = SINGLE ? 0 : NX * (N + 1);
  // Region order: regions are grouped by the phases that read them, so that the part of a problem's workspace a
  // lane-group kernel walks is one contiguous range of slab rows:
  //   pre    [ab .. st]   (ab grad c lam zl zu qs w st; writes sig rb st)
  //   accept [lam .. par) (lam zl zu qs w st d lamp ct + the x0 rows of par; writes w zl zu lam st)
  // hw, ric, pp (Riccati kernels only) come last; ph_repack_kernel relies on ric, pp being the last two.
  int o = 0;
  L.lamp = L.c = L.ct = L.ric = L.pp = L.xs = L.hred = L.gam = L.tmp = 0;
  // Linear time-invariant models with a constant cost Hessian (Model::LTI): A_k, B_k, W_k are the same for every
  // stage of a solve — ONE copy (a second for the stages folded by move blocking) instead of N; every reader finds it
  // in L1.  C3 (N = 40): 2,040 of 5,700 doubles per problem, and the most-read ones, leave the slab.
  const int ncopy = (Model::LTI && !SINGLE) ? 2 : N;
  L.ab = o; o += ncopy * (NX * NX + NX * NU);
  L.grad = o; o += L.n;
  L.sig = o; o += L.n;
  L.rb = o; o += L.n;
  if (!SINGLE) { L.c = o; o += L.m; }
  L.lam = o; o += NX * (N + 1);
  L.zl = o; o += L.n;
  L.zu = o; o += L.n;
  L.qs = o; o += N;
  L.w = o; o += L.n;
  L.st = o; o += kStateSlots;
  L.d = o; o += L.n;
  if (!SINGLE) {
    L.lamp = o; o += L.m;
    L.ct = o; o += L.m;
  }
  L.par = o; o += NX + Model::NPG + N * Model::NPS + 2;   // +2: alignment slack for bulk-staged stage params
  L.hw = o; o += ncopy * (NZ * (NZ + 1) / 2);
  L.park = L.hw;
  if (Model::LTI && !SINGLE) { L.park = o; o += L.n + L.m; }
  if (SINGLE) {
    L.xs = o; o += NX * (N + 1);
    L.hred = o; o += (NU * N) * (NU * N + 1) / 2;
    L.gam = o; o += NX * NU * N;
    L.tmp = o; o += NZ * NU * N;
  } else {
    L.ric = o; o += N * (NU * NX + NU + NU * (NU + 1) / 2);
    L.pp = o; o += (N + 1) * (NX * (NX + 1) / 2 + NX);
  }
  L.total = o;
  return L;
}

struct SolveInfo {
  int status, iters;
  double f;       // unscaled objective
  double df;      // objective scaling factor
};

MPCV_HD bool compare_le(double lhs, double rhs, double basval) {
  return lhs - rhs <= 10.0 * 2.220446049250313e-16 * fabs(basval);
}

// ---------------------------------------------------------------------------------------
template <class Model, bool SINGLE, int LANES, class WS>
struct Ipm {
  static constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  static constexpr int NW = NZ * (NZ + 1) / 2, NPX = NX * (NX + 1) / 2, NF = NU * (NU + 1) / 2;
  static constexpr int NAB = NX * NX + NX * NU, NRIC = NU * NX + NU + NF, NPP = NPX + NX;
  static constexpr int FILTER_MAX = 16;

  const Params& P;
  const Layout& L;
  WS ws;
  Grp<LANES> g;
  const double* lbx;   // [n] original bounds, shared by the batch
  const double* ubx;
  const BndEntry* btab;   // optional precomputed relaxed bounds (shared memory); null = compute on the fly
  const int N;

  int ps_base;         // offset of the stage parameters (shifted by 0/1 to match the source's 16-byte phase)
  // persistent scalar state of one solve (saved to / restored from ws[L.st..] between phase kernels)
  double df, mu, tau, f_curr;
  double theta_max, theta_min, delta_w_last;
  // the filter entries themselves stay in the workspace (state slots 16.., 24..): arrays indexed by a run-time
  // nfil would live in local memory and be copied through it at every hand-over between phases
  int nfil, iter;
  bool resto_sigma = false;   // identity-Hessian factorisations add the diagonal in L.sig (restoration: affine scaling)
  int acc_count;        // consecutive iterates within the acceptable tolerances
  double f_last;        // objective at the previous convergence test
  MPCV_D double& fil_phi(int q) const { return ws[L.st + 16 + q]; }
  MPCV_D double& fil_th(int q) const { return ws[L.st + 40 + q]; }
  // Filter augmentation as IPOPT's Filter::AddEntry does it: entries the new one dominates (both coordinates >= the
  // new ones) leave the filter first — whatever they reject, the new entry rejects too.  IPOPT's filter is
  // unbounded; this one holds FILTER_MAX entries (the largest filter over the cold-started C2 batch is 5, over the
  // all-zeros-guess batch of SURVEY Appendix E 17: tests/test_oracle_golden.py), drops the OLDEST entry when full and
  // counts that in the handle's diagnostics (Params::diag[0], mpcv_diag).  Every lane computes the same survivor
  // count from the same reads; lane 0 rewrites the entries.
  MPCV_D void filter_add(double phi_e, double th_e) {
    int kept = 0;
    unsigned keep = 0u;
    for (int q = 0; q < nfil; ++q) {
      if (!(fil_phi(q) >= phi_e && fil_th(q) >= th_e)) { keep |= 1u << q; ++kept; }
    }
    g.sync();                      // all lanes have read the entries
    if (g.lane == 0) {
      int k = 0;
      for (int q = 0; q < nfil; ++q) {
        if (keep & (1u << q)) { if (k != q) { fil_phi(k) = fil_phi(q); fil_th(k) = fil_th(q); } ++k; }
      }
      if (kept == FILTER_MAX) {    // full: the oldest entry goes
        for (int q = 1; q < FILTER_MAX; ++q) { fil_phi(q - 1) = fil_phi(q); fil_th(q - 1) = fil_th(q); }
#if defined(__CUDA_ARCH__)
        if (P.diag) atomicAdd(P.diag, 1ull);
#else
        if (P.diag) ++*P.diag;
#endif
      }
      const int at = kept == FILTER_MAX ? FILTER_MAX - 1 : kept;
      fil_phi(at) = phi_e; fil_th(at) = th_e;
    }
    nfil = kept == FILTER_MAX ? FILTER_MAX : kept + 1;
    g.sync();
  }
  // search direction -> line search hand-over
  double ls_alpha_max, ls_theta, ls_gBD, ls_phi;
  // barrier log-sum of the current iterate = log-sum of the trial point accepted last (same slacks): reused by
  // direction_post instead of ~2 logs per bounded variable
  double lg_curr, dmp_curr;   // (and the one-sided slack sum of the damping term)
  bool lg_valid;
  mutable double lg_trial, dmp_trial;   // log-sum / slack sum of the most recent trial evaluation

  MPCV_D Ipm(const Params& p, const Layout& l, WS w, Grp<LANES> grp, const double* lb, const double* ub,
             const BndEntry* tab = nullptr)
      : P(p), L(l), ws(w), g(grp), lbx(lb), ubx(ub), btab(tab), N(l.N), ps_base(l.par + NX + Model::NPG),
        df(1.0), mu(0.1), tau(0.99), f_curr(0), theta_max(-1.0), theta_min(-1.0), delta_w_last(0.0), nfil(0),
        iter(0), acc_count(0), f_last(-1e50), ls_alpha_max(1.0), ls_theta(0.0), ls_gBD(0.0), ls_phi(0.0), lg_curr(0.0), dmp_curr(0.0), lg_valid(false), lg_trial(0.0), dmp_trial(0.0) {}

  // ---- state hand-over between phase kernels ---------------------------------------------------
  MPCV_D void save_state(int status) const {
    const int o = L.st;
    if (g.lane != 0) return;
    ws[o + 0] = df; ws[o + 1] = mu; ws[o + 2] = tau; ws[o + 3] = f_curr;
    ws[o + 4] = theta_max; ws[o + 5] = theta_min; ws[o + 6] = delta_w_last;
    ws[o + 7] = (double)nfil; ws[o + 8] = (double)iter; ws[o + 9] = (double)status;
    ws[o + 10] = ls_alpha_max; ws[o + 11] = ls_theta; ws[o + 12] = ls_gBD; ws[o + 13] = ls_phi;
    ws[o + 35] = lg_curr; ws[o + 36] = lg_valid ? 1.0 : 0.0;
    ws[o + 37] = (double)acc_count; ws[o + 38] = f_last; ws[o + 39] = dmp_curr;
  }
  MPCV_D int load_state() {
    const int o = L.st;
    df = ws[o + 0]; mu = ws[o + 1]; tau = ws[o + 2]; f_curr = ws[o + 3];
    theta_max = ws[o + 4]; theta_min = ws[o + 5]; delta_w_last = ws[o + 6];
    nfil = (int)ws[o + 7]; iter = (int)ws[o + 8];
    ls_alpha_max = ws[o + 10]; ls_theta = ws[o + 11]; ls_gBD = ws[o + 12]; ls_phi = ws[o + 13];
    lg_curr = ws[o + 35]; lg_valid = ws[o + 36] != 0.0;
    acc_count = (int)ws[o + 37]; f_last = ws[o + 38]; dmp_curr = ws[o + 39];
    return (int)ws[o + 9];
  }

  // ---- variable indexing ------------------------------------------------------------------
  MPCV_D int ix(int k, int i) const { return k * NZ + i; }            // multiple shooting only
  MPCV_D int iu(int k, int i) const { return SINGLE ? k * NU + i : k * NZ + NX + i; }
  MPCV_D bool blocked(int k) const { return Model::HAS_UPREV && P.ntu > 0 && k >= P.ntu; }
  // storage index of stage k's (A, B) and Hessian blocks: the stage itself, or one of the two shared copies
  static constexpr bool kSharedStages = Model::LTI && !SINGLE;
  MPCV_D int sk(int k) const { return kSharedStages ? (blocked(k) ? 1 : 0) : k; }
  // stage and component of variable i (multiple shooting)
  MPCV_D bool var_is_blocked_u(int i) const {
    if (!(Model::HAS_UPREV && P.ntu > 0)) return false;
    const int k = SINGLE ? i / NU : i / NZ;
    const int c = SINGLE ? i % NU : i % NZ - NX;
    return c == 0 && k < N && k >= P.ntu;
  }

  struct Bnd { double lo, hi; bool hasl, hasu, fixed; };
  MPCV_D Bnd bnd_compute(int i) const {
    Bnd b;
    const double l = lbx[i], u = ubx[i];
    b.hasl = l > -kInfBound;
    b.hasu = u < kInfBound;
    b.fixed = (b.hasl && b.hasu && l == u) || var_is_blocked_u(i);
    if (b.fixed) b.hasl = b.hasu = false;
    b.lo = b.hasl ? l - P.bound_relax * fmax(1.0, fabs(l)) : -INFINITY;
    b.hi = b.hasu ? u + P.bound_relax * fmax(1.0, fabs(u)) : INFINITY;
    return b;
  }
  // Rows of variables without a bound are never touched: their multipliers stay 0 from start() on and their Sigma is 0,
  // so the loads are predicated on the (shared, shared-memory resident) bound flags.  For the linear models, where
  // only the control is boxed (C3: 40 of 245 variables), that is a tenth of the slab traffic of a sweep.
  // Measured (same box): C3 139.3 -> 130.3 ms, C5 50.4 -> 44.6 ms per step; the unicycle batch C2, where 40 of 53
  // variables are boxed, 11.64 -> 11.76 (the flag look-ups cost more than 11 skipped rows save) — so the predication
  // is a property of the model (Model::LTI: the linear family).  Either way the values are the same bits.
  static constexpr bool kSkipUnbounded = Model::LTI;
  MPCV_D bool bounded(const Bnd& b) const { return !kSkipUnbounded || b.hasl || b.hasu; }
  MPCV_D double zl_at(int i, const Bnd& b) const { return (!kSkipUnbounded || b.hasl) ? ws[L.zl + i] : 0.0; }
  MPCV_D double zu_at(int i, const Bnd& b) const { return (!kSkipUnbounded || b.hasu) ? ws[L.zu + i] : 0.0; }
  MPCV_D double sig_at(int i) const {
    if (!kSkipUnbounded) return ws[L.sig + i];
    const Bnd b = bnd(i);
    return (b.hasl || b.hasu) ? ws[L.sig + i] : 0.0;
  }
  // Lane-strided loop over [0, n) in batches of UNR iterations per lane: the loads of the whole batch
  // (ld) are issued before the first value is used (use), so a lane has UNR x more memory requests in
  // flight.  Iterations run in ascending i per lane, exactly like the plain loop.
#ifndef MPCV_UNR
#define MPCV_UNR 1   /* measured on B200: batching costs more in registers than it hides in latency */
#endif
  static constexpr int UNR = LANES >= 32 ? (MPCV_UNR > 2 ? 2 : MPCV_UNR) : MPCV_UNR;
  template <class LD, class USE>
  MPCV_D void lane_loop(int n, LD ld, USE use) const {
    for (int i0 = g.lane; i0 < n; i0 += LANES * UNR) {
      decltype(ld(0)) v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) { const int i = i0 + u * LANES; if (i < n) v[u] = ld(i); }
#pragma unroll
      for (int u = 0; u < UNR; ++u) { const int i = i0 + u * LANES; if (i < n) use(i, v[u]); }
    }
  }
  struct V1 { double a; };
  struct V2 { double a, b; };
  struct V3 { double a, b, c; };
  struct V4 { double a, b, c, d; };
  struct V6 { double a, b, c, d, e, f; };

  MPCV_D Bnd bnd(int i) const {
    if (btab) {
#if defined(__CUDA_ARCH__)
      __builtin_assume(__isShared(btab));   // LDS instead of generic loads
#endif
      const BndEntry e = btab[i];
      Bnd b;
      b.lo = e.lo; b.hi = e.hi;
      b.hasl = (e.flags & 1) != 0; b.hasu = (e.flags & 2) != 0; b.fixed = (e.flags & 4) != 0;
      return b;
    }
    return bnd_compute(i);
  }
  MPCV_D BndEntry bnd_entry(int i) const {
    const Bnd b = bnd_compute(i);
    BndEntry e;
    e.lo = b.lo; e.hi = b.hi; e.flags = (b.hasl ? 1 : 0) | (b.hasu ? 2 : 0) | (b.fixed ? 4 : 0); e.pad = 0;
    return e;
  }

  MPCV_D WsView<WS> pg() const { return WsView<WS>{ws, L.par + NX}; }
  MPCV_D WsView<WS> ps(int k) const { return WsView<WS>{ws, ps_base + k * Model::NPS}; }

  // ---- stage access -----------------------------------------------------------------------
  // load (x_k,u_k) of the vector at offset `off` shifted by alpha * (vector at offset `doff`)
  MPCV_D void load_xu(int k, int off, double alpha, int doff, double* x, double* u) const {
    if (!SINGLE) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = ws[off + ix(k, i)];
        if (alpha != 0.0) v += alpha * ws[doff + ix(k, i)];
        x[i] = v;
      }
    }
    if (k < N) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double v = ws[off + iu(k, i)];
        if (alpha != 0.0) v += alpha * ws[doff + iu(k, i)];
        u[i] = v;
      }
      if (blocked(k)) u[0] = x[NX - 1];
    }
  }

  // ---- derivatives at the current iterate ---------------------------------------------------
  // One shooting interval: fills ab, hw (when want_hess), grad (df * grad f) and the defect
  // c_{k+1} of stage k; returns the interval cost q_k.  (Multiple shooting; stage-parallel.)
  MPCV_D double der_stage(int k, bool want_hess) const {
    double x[NX], u[NU], lamn[NX], xn[NX], A[NX * NX], B[NX * NU], q, gq[NZ], W[NW];
    load_xu(k, L.w, 0.0, 0, x, u);
#pragma unroll
    for (int i = 0; i < NX; ++i) lamn[i] = ws[L.lam + (k + 1) * NX + i];
    Model::der(P, x, u, pg(), ps(k), lamn, df, want_hess, xn, A, B, &q, gq, W);
    if (blocked(k)) fold_blocked(A, B, gq, want_hess ? W : nullptr);
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) ws[L.ab + sk(k) * NAB + i] = A[i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) ws[L.ab + sk(k) * NAB + NX * NX + i] = B[i];
    if (want_hess) {
#pragma unroll
      for (int i = 0; i < NW; ++i) ws[L.hw + sk(k) * NW + i] = W[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) ws[L.grad + ix(k, i)] = df * gq[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) ws[L.grad + iu(k, i)] = df * gq[NX + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) ws[L.c + (k + 1) * NX + i] = xn[i] - ws[L.w + ix(k + 1, i)];
    return q;
  }
  // rows that belong to no interval: c_0 = xbar - X0 (MS:125-130) and the terminal gradient
  MPCV_D void der_terminal() const {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      ws[L.c + i] = ws[L.par + i] - ws[L.w + ix(0, i)];
      ws[L.grad + ix(N, i)] = 0.0;                         // no terminal cost in the scripts
    }
  }
  // scaled objective from the per-stage costs left in ws[L.qs..] by the stage-parallel kernels
  // (same summation order as the in-line accumulation of eval_derivatives / eval_trial)
  MPCV_D double sum_stage_costs() const {
    double fpart = 0.0;
    for (int k = 0; k < N; ++k) fpart += ws[L.qs + k];
    return df * fpart;
  }

  // fills ab, hw (when want_hess), grad (df * grad f), c (MS) and f_curr (scaled objective)
  MPCV_DN void eval_derivatives(bool want_hess) {
    if (SINGLE) {
      eval_derivatives_single(want_hess);
      return;
    }
    double fpart = 0.0;
    for (int k = g.lane; k < N; k += LANES) fpart += der_stage(k, want_hess);
    if (g.lane == 0) der_terminal();
    f_curr = df * g.sum(fpart);
    g.sync();
  }

  // exact chain rule for a blocked stage: z = (x, u0 = x_{NX-1})
  MPCV_D void fold_blocked(double* A, double* B, double* gq, double* W) const {
    constexpr int j = NX - 1;
#pragma unroll
    for (int i = 0; i < NX; ++i) { A[i * NX + j] += B[i * NU]; B[i * NU] = 0.0; }
    gq[j] += gq[NX]; gq[NX] = 0.0;
    if (W) {
      const double wuu = W[tri(NX, NX)], wuj = W[tri(NX, j)];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        if (i != j) W[tri(j, i)] += W[tri(NX, i)];
        W[tri(NX, i)] = 0.0;
      }
      W[tri(j, j)] += 2.0 * wuj + wuu;
      W[tri(NX, NX)] = 0.0;
    }
  }

  // single shooting: rollout, costate sweep, exact Hessian blocks with the costates as multipliers
  MPCV_DN void eval_derivatives_single(bool want_hess) {
    if (g.lane == 0) {
      double x[NX], u[NU], xn[NX], A[NX * NX], B[NX * NU], q, gq[NZ], W[NW], lamn[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = ws[L.par + i]; ws[L.xs + i] = x[i]; }
      double fsum = 0.0;
      // forward: states, A, B, cost gradient
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = ws[L.w + iu(k, i)];
        if (blocked(k)) u[0] = x[NX - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) lamn[i] = 0.0;
        Model::der(P, x, u, pg(), ps(k), lamn, df, false, xn, A, B, &q, gq, W);
        if (blocked(k)) fold_blocked(A, B, gq, nullptr);
        fsum += q;
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) ws[L.ab + sk(k) * NAB + i] = A[i];
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) ws[L.ab + sk(k) * NAB + NX * NX + i] = B[i];
        // stash dq/dz in the Hessian slot until the costates are known
#pragma unroll
        for (int i = 0; i < NZ; ++i) ws[L.hw + sk(k) * NW + i] = gq[i];
#pragma unroll
        for (int i = 0; i < NX; ++i) { x[i] = xn[i]; ws[L.xs + (k + 1) * NX + i] = xn[i]; }
      }
      f_curr = df * fsum;
      // backward costate (adjoint) sweep: mu_N = 0, mu_k = df q_x + A_k' mu_{k+1};
      // reduced gradient dJ/du_k = df q_u + B_k' mu_{k+1}
#pragma unroll
      for (int i = 0; i < NX; ++i) { ws[L.lam + N * NX + i] = 0.0; lamn[i] = 0.0; }
      for (int k = N - 1; k >= 0; --k) {
        double mk[NX];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = df * ws[L.hw + sk(k) * NW + NX + i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += ws[L.ab + sk(k) * NAB + NX * NX + j * NU + i] * lamn[j];
          ws[L.grad + iu(k, i)] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = df * ws[L.hw + sk(k) * NW + i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += ws[L.ab + sk(k) * NAB + j * NX + i] * lamn[j];
          mk[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { lamn[i] = mk[i]; ws[L.lam + k * NX + i] = mk[i]; }
      }
      if (want_hess) {
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int i = 0; i < NX; ++i) { x[i] = ws[L.xs + k * NX + i]; lamn[i] = ws[L.lam + (k + 1) * NX + i]; }
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = ws[L.w + iu(k, i)];
          if (blocked(k)) u[0] = x[NX - 1];
          Model::der(P, x, u, pg(), ps(k), lamn, df, true, xn, A, B, &q, gq, W);
          if (blocked(k)) fold_blocked(A, B, gq, W);
#pragma unroll
          for (int i = 0; i < NW; ++i) ws[L.hw + sk(k) * NW + i] = W[i];
        }
      }
    }
    f_curr = g.sum(g.lane == 0 ? f_curr : 0.0);
    g.sync();
  }

  // Barrier log-sum  sum_i log(s_i)  accumulated as log of partial products: eight slacks per log() call.
  // A slack lies between ~1e-11 (active bound at the final mu) and the width of the box, so a product of
  // eight stays far inside the double range; the sum differs from the term-by-term one by a few ulp, far
  // below the 10 eps |phi| slack of the filter's comparisons.  FP64 log is ~50 instructions: this is a
  // quarter of the instructions of the line-search kernels.
  struct LogSum {
    double lg = 0.0, prod = 1.0;
    double damp = 0.0;      // sum of the slacks of ONE-sided bounds (IPOPT's linear damping term kappa_d mu (x - x_L))
    int cnt = 0;
    MPCV_D void add(double s) {
      prod *= s;
      if (++cnt == 8) { lg += log(prod); prod = 1.0; cnt = 0; }
    }
    // slacks of variable i with bounds b at value v; returns false when a slack is not positive
    template <class B>
    MPCV_D bool add_var(const B& b, double v) {
      bool ok = true;
      if (b.hasl) { const double s = v - b.lo; if (!(s > 0.0)) ok = false; add(s); if (!b.hasu) damp += s; }
      if (b.hasu) { const double s = b.hi - v; if (!(s > 0.0)) ok = false; add(s); if (!b.hasl) damp += s; }
      return ok;
    }
    MPCV_D double value() const { return cnt ? lg + log(prod) : lg; }
  };
  // kappa_d of IPOPT (one-sided bounds only: none of the reference scripts has one; two-sided boxes are undamped)
  static constexpr double kKappaD = 1e-4;
  MPCV_D double phi_of(double f, double lg, double damp) const { return f - mu * lg + kKappaD * mu * damp; }

  // ---- objective / constraint violation / barrier at  w + alpha * (vector at doff) -------------
  // returns scaled f; theta = ||c||_1; barrier log terms; optionally stores the residuals in ct
  MPCV_DN void eval_trial(double alpha, int doff, bool store_ct, double* f_out, double* theta_out,
                         double* phi_out) const {
    double fpart = 0.0, thpart = 0.0, logpart = 0.0;
    bool bad = false;
    if (SINGLE) {
      if (g.lane == 0) {
        double x[NX], u[NU], xn[NX], q;
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = ws[L.par + i];
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = ws[L.w + iu(k, i)] + alpha * ws[doff + iu(k, i)];
          if (blocked(k)) u[0] = x[NX - 1];
          Model::val(P, x, u, pg(), ps(k), xn, &q);
          fpart += q;
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = xn[i];
        }
      }
    } else {
      for (int k = g.lane; k < N; k += LANES) {
        double x[NX], u[NU], xn[NX], q;
        load_xu(k, L.w, alpha, doff, x, u);
        Model::val(P, x, u, pg(), ps(k), xn, &q);
        fpart += q;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          const double r = xn[i] - (ws[L.w + ix(k + 1, i)] + alpha * ws[doff + ix(k + 1, i)]);
          thpart += fabs(r);
          if (store_ct) ws[L.ct + (k + 1) * NX + i] = r;
        }
      }
      if (g.lane == 0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          const double r = ws[L.par + i] - (ws[L.w + ix(0, i)] + alpha * ws[doff + ix(0, i)]);
          thpart += fabs(r);
          if (store_ct) ws[L.ct + i] = r;
        }
      }
    }
    LogSum ls;
    for (int i = g.lane; i < L.n; i += LANES) {
      const Bnd b = bnd(i);
      if (b.hasl || b.hasu) {
        if (!ls.add_var(b, ws[L.w + i] + alpha * ws[doff + i])) bad = true;
      }
    }
    logpart = ls.value();
    const double f = df * g.sum(fpart);
    const double th = g.sum(thpart);
    double lg = g.sum(logpart);
    lg_trial = lg;
    dmp_trial = g.sum(ls.damp);
    const bool anybad = g.max(bad ? 1.0 : 0.0) > 0.0;
    *f_out = f;
    *theta_out = th;
    *phi_out = anybad ? INFINITY : phi_of(f, lg, dmp_trial);
    if (store_ct) g.sync();
  }

  // Stage-parallel form of eval_trial for the phase-kernel pipeline (multiple shooting):
  // trial_stage(k) leaves the defect of interval k in ct and its cost in qs; trial_reduce() adds the
  // x0 rows, the barrier terms and sums in the order eval_trial accumulates.
  MPCV_D void trial_stage(int k, double alpha, int doff) const {
    double x[NX], u[NU], xn[NX], q;
    load_xu(k, L.w, alpha, doff, x, u);
    Model::val(P, x, u, pg(), ps(k), xn, &q);
    ws[L.qs + k] = q;
#pragma unroll
    for (int i = 0; i < NX; ++i)
      ws[L.ct + (k + 1) * NX + i] = xn[i] - (ws[L.w + ix(k + 1, i)] + alpha * ws[doff + ix(k + 1, i)]);
  }
  MPCV_D void trial_reduce(double alpha, int doff, double* f_out, double* theta_out, double* phi_out) const {
    double thpart = 0.0, logpart = 0.0;
    bool bad = false;
    lane_loop(L.m - NX, [&](int j) { return V1{ws[L.ct + NX + j]}; }, [&](int, const V1& v) { thpart += fabs(v.a); });
    for (int i = g.lane; i < NX; i += LANES) {
      const double r = ws[L.par + i] - (ws[L.w + ix(0, i)] + alpha * ws[doff + ix(0, i)]);
      thpart += fabs(r);
      ws[L.ct + i] = r;
    }
    LogSum ls;
    lane_loop(L.n, [&](int i) { return V2{ws[L.w + i], ws[doff + i]}; }, [&](int i, const V2& q) {
      const Bnd b = bnd(i);
      if (b.hasl || b.hasu) {
        if (!ls.add_var(b, q.a + alpha * q.b)) bad = true;
      }
    });
    logpart = ls.value();
    const double f = sum_stage_costs();
    const double lg = g.sum(logpart);
    lg_trial = lg;
    dmp_trial = g.sum(ls.damp);
    const bool anybad = g.max(bad ? 1.0 : 0.0) > 0.0;
    *f_out = f;
    *theta_out = g.sum(thpart);
    *phi_out = anybad ? INFINITY : phi_of(f, lg, dmp_trial);
    g.sync();
  }

  // ---- error measures (scaled as in IPOPT's E_mu) ------------------------------------------------
  struct Err { double dual, prim, cmin, cmax, sd, sc; bool any_bound; };
  MPCV_DN Err errors() const {
    double dual = 0.0, prim = 0.0, cmin = INFINITY, cmax = -INFINITY, zsum = 0.0, lsum = 0.0, nz = 0.0;
    // dual infeasibility  grad + J' lam - zl + zu, stage-parallel
    if (SINGLE) {
      for (int i = g.lane; i < L.n; i += LANES) {
        const Bnd b = bnd(i);
        if (b.fixed) continue;
        dual = fmax(dual, fabs(ws[L.grad + i] - zl_at(i, b) + zu_at(i, b)));
      }
    } else {
      for (int k = g.lane; k <= N; k += LANES) {
        double lk[NX], ln[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) { lk[i] = ws[L.lam + k * NX + i]; ln[i] = (k < N) ? ws[L.lam + (k + 1) * NX + i] : 0.0; }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          const int v = ix(k, i);
          const Bnd bv = bnd(v);
          double d = ws[L.grad + v] - lk[i] - zl_at(v, bv) + zu_at(v, bv);
          if (k < N) {
#pragma unroll
            for (int j = 0; j < NX; ++j) d += ws[L.ab + sk(k) * NAB + j * NX + i] * ln[j];
          }
          dual = fmax(dual, fabs(d));
        }
        if (k < N) {
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            const int v = iu(k, i);
            const Bnd bv = bnd(v);
            if (bv.fixed) continue;
            double d = ws[L.grad + v] - zl_at(v, bv) + zu_at(v, bv);
#pragma unroll
            for (int j = 0; j < NX; ++j) d += ws[L.ab + sk(k) * NAB + NX * NX + j * NU + i] * ln[j];
            dual = fmax(dual, fabs(d));
          }
        }
      }
      lane_loop(L.m, [&](int i) { return V2{ws[L.c + i], ws[L.lam + i]}; }, [&](int, const V2& v) {
        prim = fmax(prim, fabs(v.a));
        lsum += fabs(v.b);
      });
    }
    lane_loop(L.n, [&](int i) {
      const Bnd b = bnd(i);
      return V3{bounded(b) ? ws[L.w + i] : 0.0, zl_at(i, b), zu_at(i, b)};
    }, [&](int i, const V3& v) {
      const Bnd b = bnd(i);
      if (b.hasl) {
        const double z = v.b, c = (v.a - b.lo) * z;
        cmin = fmin(cmin, c); cmax = fmax(cmax, c); zsum += fabs(z); nz += 1.0;
      }
      if (b.hasu) {
        const double z = v.c, c = (b.hi - v.a) * z;
        cmin = fmin(cmin, c); cmax = fmax(cmax, c); zsum += fabs(z); nz += 1.0;
      }
    });
    Err e;
    dual = g.max(dual); prim = g.max(prim); e.cmin = g.min(cmin); e.cmax = g.max(cmax);
    zsum = g.sum(zsum); lsum = g.sum(lsum); nz = g.sum(nz);
    const double s_max = 100.0;
    e.any_bound = nz > 0.0;
    e.sd = (L.m + nz > 0.0) ? fmax(s_max, (lsum + zsum) / (L.m + nz)) / s_max : 1.0;
    e.sc = e.any_bound ? fmax(s_max, zsum / nz) / s_max : 1.0;
    e.dual = dual / e.sd;
    e.prim = prim;
    return e;
  }
  MPCV_D static double compl_err(const Err& e, double mu_t) {
    if (!e.any_bound) return 0.0;
    return fmax(fabs(e.cmax - mu_t), fabs(e.cmin - mu_t)) / e.sc;
  }
  MPCV_D static double Emu(const Err& e, double mu_t) { return fmax(e.dual, fmax(e.prim, compl_err(e, mu_t))); }

  // ---- Sigma and barrier gradient of variable i -----------------------------------------------------
  // computed once per iteration (after the barrier update) and reused by the factorisation
  // retries, the vector recursions and the merit-function slope
  MPCV_D void sigma_r_compute(int i, double* sg, double* r) const {
    const Bnd b = bnd(i);
    double s = 0.0, ri = ws[L.grad + i];
    // one reciprocal per bound serves Sigma = z / s and the barrier gradient mu / s
    if (b.hasl) { const double inv = 1.0 / (ws[L.w + i] - b.lo); s += ws[L.zl + i] * inv; ri -= mu * inv; }
    if (b.hasu) { const double inv = 1.0 / (b.hi - ws[L.w + i]); s += ws[L.zu + i] * inv; ri += mu * inv; }
    if (b.hasl != b.hasu) ri += b.hasl ? kKappaD * mu : -kKappaD * mu;
    *sg = s; *r = ri;
  }
  MPCV_D void prepare_barrier() const {
    lane_loop(L.n, [&](int i) {
      const Bnd b = bnd(i);
      return V4{ws[L.grad + i], bounded(b) ? ws[L.w + i] : 0.0, zl_at(i, b), zu_at(i, b)};
    }, [&](int i, const V4& v) {
      const Bnd b = bnd(i);
      double s = 0.0, ri = v.a;
      // one reciprocal per bound serves Sigma = z / s and the barrier gradient mu / s
      if (b.hasl) { const double inv = 1.0 / (v.b - b.lo); s += v.c * inv; ri -= mu * inv; }
      if (b.hasu) { const double inv = 1.0 / (b.hi - v.b); s += v.d * inv; ri += mu * inv; }
      if (b.hasl != b.hasu) ri += b.hasl ? kKappaD * mu : -kKappaD * mu;      // linear damping of one-sided bounds
      if (bounded(b)) ws[L.sig + i] = s;
      ws[L.rb + i] = ri;
    });
    g.sync();
  }
  MPCV_D void sigma_r(int i, double* sg, double* r) const { *sg = sig_at(i); *r = ws[L.rb + i]; }

  // ---- Riccati factorisation (multiple shooting) ------------------------------------------------------
  // identity=true replaces the Lagrangian Hessian by I and drops Sigma (least-squares multipliers).
  // Stores per stage: K (NU x NX), chol(F) packed, P_k packed; returns false on wrong inertia.
  // L1 prefetch of a workspace element (no register cost; no-op on the host harness)
  MPCV_D void pf(int i) const {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(&ws[i]));
#else
    (void)i;
#endif
  }
  // operands of one backward Riccati step, loaded as a block; the rows of stage k-1 are prefetched into L1
  // while stage k computes (the recursion is a chain of dependent steps: without the prefetch every step
  // starts with an exposed HBM round trip)
  struct FacIn { double A[NX * NX], B[NX * NU], W[NW], sg[NZ]; };
  MPCV_D void prefetch_fac(int k, bool identity) const {
    if (LANES != 1) return;          // shared-memory workspaces have nothing to prefetch
#pragma unroll
    for (int i = 0; i < NAB; ++i) pf(L.ab + sk(k) * NAB + i);
    if (!identity) {
#pragma unroll
      for (int i = 0; i < NW; ++i) pf(L.hw + sk(k) * NW + i);
#pragma unroll
      for (int i = 0; i < NZ; ++i) pf(L.sig + k * NZ + i);
    }
  }
  MPCV_D void load_fac(int k, bool identity, FacIn& s) const {
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) s.A[i] = ws[L.ab + sk(k) * NAB + i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) s.B[i] = ws[L.ab + sk(k) * NAB + NX * NX + i];
    if (identity) {
#pragma unroll
      for (int i = 0; i < NW; ++i) s.W[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NZ; ++i) { s.W[tri(i, i)] = 1.0; s.sg[i] = resto_sigma ? sig_at(k * NZ + i) : 0.0; }
    } else {
#pragma unroll
      for (int i = 0; i < NW; ++i) s.W[i] = ws[L.hw + sk(k) * NW + i];
#pragma unroll
      for (int i = 0; i < NZ; ++i) s.sg[i] = sig_at(k * NZ + i);
    }
  }

  MPCV_DN bool riccati_factor(double dw, bool identity) { return riccati_factor_x<false>(dw, identity, 0, -1); }
  // lane groups share the stage algebra (riccati_factor_lanes); a lone lane keeps everything in registers
#if defined(MPCV_HOST_LANE_RICCATI)
  static constexpr bool kLaneRiccati = true;      // tests/hostsim: replay the lane-parallel form with one lane
#else
  static constexpr bool kLaneRiccati = LANES > 1;
#endif
  template <bool FUSE>
  MPCV_D bool riccati_factor_x(double dw, bool identity, int rmode, int coff) {
    // (the scratch of the lane form must fit into the step buffers: always true but for toy horizons)
    if (kLaneRiccati && (NX + 2 * NU) * (NX + 1) <= L.n + L.m) return riccati_factor_lanes<FUSE>(dw, identity, rmode, coff);
    return riccati_factor_t<FUSE>(dw, identity, rmode, coff);
  }

  // ---- lane-parallel Riccati factorisation ---------------------------------------------------------------
  // The backward recursion is sequential over the horizon, but inside a stage the NX columns of (P A, G, K),
  // the vector part (P c + p, g, k_ff) and afterwards the entries of P_k and p_k are independent outputs: a
  // stage is TWO short steps on the group's lanes with the operands exchanged through the workspace, instead
  // of ~350 dependent instructions on lane 0.  Every output is accumulated by one lane with the expressions
  // and the order of riccati_factor_t, so the two forms agree bit for bit (tests/hostsim replays this one with
  // a single lane).  Scratch: the step buffers d, lam+ (contiguous, dead until the forward sweep writes them).
  template <bool FUSE>
  MPCV_D bool riccati_factor_lanes(double dw, bool identity, int rmode, int coff) {
    // scratch in the step buffers: y = P A | P c + p  (NX x (NX+1)), z = G | g  (NU x (NX+1)), t = K | kff
    constexpr int NC = NX + 1;
    const WS sY = ws.view(L.d), sZ = ws.view(L.d + NX * NC), sT = ws.view(L.d + NX * NC + NU * NC);
    // terminal stage: P_N = Sigma_N + dw (identity: I), p_N = r_N
    {
      const WS pn = ws.view(L.pp + N * NPP);
      for (int i = g.lane; i < NPX + (FUSE ? NX : 0); i += LANES) {
        if (i < NPX) {
          int r = 0;
#pragma unroll
          for (int q = 1; q < NX; ++q) r += (i >= tri(q, 0)) ? 1 : 0;
          const int c = i - tri(r, 0);
          double v = 0.0;
          if (r == c) {
            double sg = 0.0, rr;
            if (!identity || resto_sigma) sigma_r(ix(N, r), &sg, &rr);
            v = (identity ? 1.0 : 0.0) + sg + dw;
          }
          pn[i] = v;
        } else {
          pn[i] = rvar(rmode, ix(N, i - NPX));
        }
      }
    }
    g.sync();
    int ok = 1;
    for (int k = N - 1; k >= 0; --k) {
      const WS ab = ws.view(L.ab + sk(k) * NAB), hw = ws.view(L.hw + sk(k) * NW);
      const WS ric = ws.view(L.ric + k * NRIC), ppn = ws.view(L.pp + (k + 1) * NPP), ppk = ws.view(L.pp + k * NPP);
      // ---- step A: item c < NX = column c of (P A, G, K); item NX = the vector part (P c + p, g, kff).  One
      // instruction stream for both kinds (operands selected, not branched on): the lanes of a group stay together.
      int okl = 1;
      for (int it = g.lane; it < NX + (FUSE ? 1 : 0); it += LANES) {
        const bool col = it < NX;
        double Pm[NX * NX], B[NX * NU], PB[NX * NU], F[NU * NU], y[NX], z[NU], t[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) { const double v = ppn[tri(i, j)]; Pm[i * NX + j] = v; Pm[j * NX + i] = v; }
        }
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) B[i] = ab[NX * NX + i];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            double v = 0.0;
#pragma unroll
            for (int l = 0; l < NX; ++l) v += Pm[i * NX + l] * B[l * NU + j];
            PB[i * NU + j] = v;
          }
        }
        // F = Ruu + Sigma_u + dw + B'PB (lower)
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            double v = identity ? (i == j ? 1.0 : 0.0) : hw[tri(NX + i, NX + j)];
            if ((!identity || resto_sigma) && i == j) v += sig_at(k * NZ + NX + i) + dw;
#pragma unroll
            for (int l = 0; l < NX; ++l) v += B[l * NU + i] * PB[l * NU + j];
            F[i * NU + j] = v;
          }
        }
        bool fixed[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) fixed[i] = bnd(iu(k, i)).fixed;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          if (fixed[i]) {
#pragma unroll
            for (int j = 0; j < NU; ++j) { if (j <= i) F[i * NU + j] = 0.0; else F[j * NU + i] = 0.0; }
            F[i * NU + i] = 1.0;
          }
        }
        // Cholesky of F in place (lower, reciprocal pivots on the diagonal)
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double dj = F[j * NU + j];
#pragma unroll
          for (int l = 0; l < j; ++l) dj -= F[j * NU + l] * F[j * NU + l];
          if (!(dj > 0.0) || !(dj < INFINITY)) { okl = 0; dj = 1.0; }
          dj = rsqrt_(dj);
          F[j * NU + j] = dj;
#pragma unroll
          for (int i = j + 1; i < NU; ++i) {
            double v = F[i * NU + j];
#pragma unroll
            for (int l = 0; l < j; ++l) v -= F[i * NU + l] * F[j * NU + l];
            F[i * NU + j] = v * dj;
          }
        }
        // y = P a + y0:  column c: a = A[:, c], y0 = 0  (P A);  vector item: a = c_{k+1}, y0 = p_{k+1}  (P c + p)
        const WS av = col ? ws.view(L.ab + sk(k) * NAB + it) : ws.view(coff >= 0 ? coff + (k + 1) * NX : L.ab + sk(k) * NAB);
        const int astr = col ? NX : 1;
        const bool azero = !col && coff < 0;
        double a[NX];
#pragma unroll
        for (int l = 0; l < NX; ++l) { const double v = av[l * astr]; a[l] = azero ? 0.0 : v; }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = col ? 0.0 : ppn[NPX + i];
#pragma unroll
          for (int l = 0; l < NX; ++l) v += Pm[i * NX + l] * a[l];
          y[i] = v;
        }
        // z = z0 + B' y:  column: z0 = S_ux[:, c]  (G);  vector: z0 = r_u  (g);  fixed controls drop out
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v;
          if (col) v = identity ? 0.0 : hw[tri(NX + i, 0) + it];
          else v = rvar(rmode, iu(k, i));
#pragma unroll
          for (int l = 0; l < NX; ++l) v += B[l * NU + i] * y[l];
          z[i] = fixed[i] ? 0.0 : v;
        }
        // t = -F^{-1} z  (K[:, c] | kff)
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = -z[i];
#pragma unroll
          for (int l = 0; l < i; ++l) v -= F[i * NU + l] * t[l];
          t[i] = v * F[i * NU + i];
        }
#pragma unroll
        for (int i = NU - 1; i >= 0; --i) {
          double v = t[i];
#pragma unroll
          for (int l = i + 1; l < NU; ++l) v -= F[l * NU + i] * t[l];
          t[i] = v * F[i * NU + i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) sY[i * NC + it] = y[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          sZ[i * NC + it] = z[i];
          sT[i * NC + it] = t[i];
          ric[col ? i * NX + it : NU * NX + i] = t[i];
        }
        if (it == 0) {
#pragma unroll
          for (int i = 0; i < NU; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) ric[NU * NX + NU + tri(i, j)] = F[i * NU + j];
          }
        }
      }
      ok = g.all(okl != 0) ? 1 : 0;
      g.sync();
      if (!ok) break;
      // ---- step B: the entries of P_k = Qxx + A'PA + G'K and of p_k = r_x + A' Pd + K' g ----
      for (int it = g.lane; it < NPX + (FUSE ? NX : 0); it += LANES) {
        const bool mat = it < NPX;
        int i = 0;
#pragma unroll
        for (int q = 1; q < NX; ++q) i += (it >= tri(q, 0)) ? 1 : 0;
        const int j = mat ? it - tri(i, 0) : NX;      // column of (y, z, t): a matrix column or the vector part
        if (!mat) i = it - NPX;
        double v;
        if (mat) {
          v = identity ? (i == j ? 1.0 : 0.0) : hw[it];
          if ((!identity || resto_sigma) && i == j) v += sig_at(k * NZ + i) + dw;
        } else {
          v = rvar(rmode, ix(k, i));
        }
#pragma unroll
        for (int l = 0; l < NX; ++l) v += ab[l * NX + i] * sY[l * NC + j];
        // G'K for the matrix entries, K'g for the vector ones: (row l of z | t) x (row l of t | z)
#pragma unroll
        for (int l = 0; l < NU; ++l) v += mat ? sZ[l * NC + i] * sT[l * NC + j] : sT[l * NC + i] * sZ[l * NC + NX];
        ppk[mat ? it : NPX + i] = v;
      }
      g.sync();
    }
    return ok != 0;
  }

  // inertia probe: the factorisation without any store (only its verdict matters)
  MPCV_D bool riccati_probe(double dw) { return riccati_factor_t<false, false>(dw, false, 0, -1); }
  // FUSE: the backward VECTOR recursion of riccati_solve(rmode, coff) runs inside the factorisation loop,
  // on the A, B, K, F, P_{k+1} already in registers (the phase pipeline's Riccati kernels are HBM-bound:
  // a separate backward pass re-reads 30 of them per stage).  Same expressions in the same order as
  // riccati_solve; riccati_forward() completes the step.
  template <bool FUSE, bool STORE = true>
  MPCV_D bool riccati_factor_t(double dw, bool identity, int rmode, int coff) {
    int ok = 1;
    if (g.lane == 0) {
      double pv[NX];        // p_{k+1} (FUSE)
      if (FUSE) {
#pragma unroll
        for (int i = 0; i < NX; ++i) { pv[i] = rvar(rmode, ix(N, i)); ws[L.pp + N * NPP + NPX + i] = pv[i]; }
      }
      double Pm[NX * NX];   // P_{k+1}, full symmetric
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) Pm[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double sg = 0.0, r;
        if (!identity || resto_sigma) sigma_r(ix(N, i), &sg, &r);
        Pm[i * NX + i] = (identity ? 1.0 : 0.0) + sg + dw;
      }
      if (STORE) store_P(N, Pm);
      FacIn cur;
      for (int k = N - 1; k >= 0 && ok; --k) {
        load_fac(k, identity, cur);
        if (k > 0) prefetch_fac(k - 1, identity);
        double A[NX * NX], B[NX * NU], W[NW];
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) A[i] = cur.A[i];
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) B[i] = cur.B[i];
#pragma unroll
        for (int i = 0; i < NW; ++i) W[i] = cur.W[i];
        if (!identity || resto_sigma) {
#pragma unroll
          for (int i = 0; i < NZ; ++i) W[tri(i, i)] += cur.sg[i] + dw;
        }
        // PA = P A, PB = P B
        double PA[NX * NX], PB[NX * NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j < NX; ++j) {
            double v = 0.0;
#pragma unroll
            for (int l = 0; l < NX; ++l) v += Pm[i * NX + l] * A[l * NX + j];
            PA[i * NX + j] = v;
          }
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            double v = 0.0;
#pragma unroll
            for (int l = 0; l < NX; ++l) v += Pm[i * NX + l] * B[l * NU + j];
            PB[i * NU + j] = v;
          }
        }
        // F = Ruu + B'PB (lower), G = Sux + B'PA
        double F[NU * NU], G[NU * NX];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            double v = W[tri(NX + i, NX + j)];
#pragma unroll
            for (int l = 0; l < NX; ++l) v += B[l * NU + i] * PB[l * NU + j];
            F[i * NU + j] = v;
          }
#pragma unroll
          for (int j = 0; j < NX; ++j) {
            double v = W[tri(NX + i, j)];
#pragma unroll
            for (int l = 0; l < NX; ++l) v += B[l * NU + i] * PA[l * NX + j];
            G[i * NX + j] = v;
          }
        }
        // fixed controls (blocked / lb == ub): unit pivot, no coupling
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          if (bnd(iu(k, i)).fixed) {
#pragma unroll
            for (int j = 0; j < NU; ++j) { if (j <= i) F[i * NU + j] = 0.0; else F[j * NU + i] = 0.0; }
            F[i * NU + i] = 1.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) G[i * NX + j] = 0.0;
          }
        }
        // Cholesky of F in place (lower); the diagonal holds the RECIPROCAL of the Cholesky pivot, so the
        // triangular solves here and in riccati_solve multiply instead of divide
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          double dj = F[j * NU + j];
#pragma unroll
          for (int l = 0; l < j; ++l) dj -= F[j * NU + l] * F[j * NU + l];
          if (!(dj > 0.0) || !(dj < INFINITY)) { ok = 0; dj = 1.0; }
          dj = rsqrt_(dj);
          F[j * NU + j] = dj;
#pragma unroll
          for (int i = j + 1; i < NU; ++i) {
            double v = F[i * NU + j];
#pragma unroll
            for (int l = 0; l < j; ++l) v -= F[i * NU + l] * F[j * NU + l];
            F[i * NU + j] = v * dj;
          }
        }
        // K = -F^{-1} G  (column by column)
        double K[NU * NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
          double t[NU];
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            double v = -G[i * NX + c];
#pragma unroll
            for (int l = 0; l < i; ++l) v -= F[i * NU + l] * t[l];
            t[i] = v * F[i * NU + i];
          }
#pragma unroll
          for (int i = NU - 1; i >= 0; --i) {
            double v = t[i];
#pragma unroll
            for (int l = i + 1; l < NU; ++l) v -= F[l * NU + i] * t[l];
            t[i] = v * F[i * NU + i];
          }
#pragma unroll
          for (int i = 0; i < NU; ++i) K[i * NX + c] = t[i];
        }
        if (FUSE) {
          // Pd = P_{k+1} c_{k+1} + p_{k+1};  g = r_u + B' Pd;  kff = -F^{-1} g;  p_k = r_x + A' Pd + K' g
          double Pd[NX], gk[NU], t[NU];
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            double v = pv[i];
            if (coff >= 0) {
#pragma unroll
              for (int j = 0; j < NX; ++j) v += Pm[i * NX + j] * ws[coff + (k + 1) * NX + j];
            }
            Pd[i] = v;
          }
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            double v = rvar(rmode, iu(k, i));
#pragma unroll
            for (int j = 0; j < NX; ++j) v += B[j * NU + i] * Pd[j];
            if (bnd(iu(k, i)).fixed) v = 0.0;
            gk[i] = v;
          }
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            double v = -gk[i];
#pragma unroll
            for (int l = 0; l < i; ++l) v -= F[i * NU + l] * t[l];
            t[i] = v * F[i * NU + i];
          }
#pragma unroll
          for (int i = NU - 1; i >= 0; --i) {
            double v = t[i];
#pragma unroll
            for (int l = i + 1; l < NU; ++l) v -= F[l * NU + i] * t[l];
            t[i] = v * F[i * NU + i];
          }
#pragma unroll
          for (int i = 0; i < NU; ++i) ws[L.ric + k * NRIC + NU * NX + i] = t[i];
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            double v = rvar(rmode, ix(k, i));
#pragma unroll
            for (int j = 0; j < NX; ++j) v += A[j * NX + i] * Pd[j];
#pragma unroll
            for (int j = 0; j < NU; ++j) v += K[j * NX + i] * gk[j];
            pv[i] = v;
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) ws[L.pp + k * NPP + NPX + i] = pv[i];
        }
        // P_k = Qxx + A'PA + G'K  (symmetric)
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            double v = W[tri(i, j)];
#pragma unroll
            for (int l = 0; l < NX; ++l) v += A[l * NX + i] * PA[l * NX + j];
#pragma unroll
            for (int l = 0; l < NU; ++l) v += G[l * NX + i] * K[l * NX + j];
            Pm[i * NX + j] = v;
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = i + 1; j < NX; ++j) Pm[i * NX + j] = Pm[j * NX + i];
        }
        if (STORE) {
          store_P(k, Pm);
          const int ro = L.ric + k * NRIC;
#pragma unroll
          for (int i = 0; i < NU * NX; ++i) ws[ro + i] = K[i];
#pragma unroll
          for (int i = 0; i < NU; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) ws[ro + NU * NX + NU + tri(i, j)] = F[i * NU + j];
          }
        }
      }
    }
    ok = g.bcast(ok);
    g.sync();
    return ok != 0;
  }
  MPCV_D void store_P(int k, const double* Pm) const {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) ws[L.pp + k * NPP + tri(i, j)] = Pm[i * NX + j];
    }
  }
  MPCV_D void load_P(int k, double* Pm) const {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) { const double v = ws[L.pp + k * NPP + tri(i, j)]; Pm[i * NX + j] = v; Pm[j * NX + i] = v; }
    }
  }

  // ---- Riccati solve: vector recursions with the stored factors ------------------------------------
  // rhs: r = barrier gradient (rmode 0) or grad - zl + zu (rmode 1, least squares);
  // residuals from offset coff (L.c, L.ct) or zero (coff < 0).  Writes d and lam+ (L.lamp).
  // Both sweeps prefetch the next stage's rows into L1 (see prefetch_fac).
  struct BwdIn { double P[NPX], c[NX], A[NX * NX], B[NX * NU], K[NU * NX], F[NF], ru[NU], rx[NX]; };
  struct FwdIn { double P[NPX], p[NX], K[NU * NX], kff[NU], A[NX * NX], B[NX * NU], c[NX]; };
  MPCV_D double rvar(int rmode, int v) const {
    if (rmode == 2) return 0.0;                                        // feasibility step: minimise |d|^2 only
    if (rmode == 1) { const Bnd b = bnd(v); return ws[L.grad + v] - zl_at(v, b) + zu_at(v, b); }
    return ws[L.rb + v];
  }
  MPCV_D void load_bwd(int k, int rmode, int coff, BwdIn& s) const {
    const int ao = L.ab + sk(k) * NAB, ro = L.ric + k * NRIC;
#pragma unroll
    for (int i = 0; i < NPX; ++i) s.P[i] = ws[L.pp + (k + 1) * NPP + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) s.c[i] = (coff >= 0) ? ws[coff + (k + 1) * NX + i] : 0.0;
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) s.A[i] = ws[ao + i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) s.B[i] = ws[ao + NX * NX + i];
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) s.K[i] = ws[ro + i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s.F[i] = ws[ro + NU * NX + NU + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) s.ru[i] = rvar(rmode, iu(k, i));
#pragma unroll
    for (int i = 0; i < NX; ++i) s.rx[i] = rvar(rmode, ix(k, i));
  }
  MPCV_D void prefetch_bwd(int k, int rmode, int coff) const {
    if (LANES != 1) return;
#pragma unroll
    for (int i = 0; i < NPX; ++i) pf(L.pp + (k + 1) * NPP + i);
    if (coff >= 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) pf(coff + (k + 1) * NX + i);
    }
#pragma unroll
    for (int i = 0; i < NAB; ++i) pf(L.ab + sk(k) * NAB + i);
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) pf(L.ric + k * NRIC + i);
#pragma unroll
    for (int i = 0; i < NF; ++i) pf(L.ric + k * NRIC + NU * NX + NU + i);
    if (rmode == 0) {
#pragma unroll
      for (int i = 0; i < NZ; ++i) pf(L.rb + k * NZ + i);
    }
  }
  MPCV_D void prefetch_fwd(int k, int coff) const {
    if (LANES != 1) return;
#pragma unroll
    for (int i = 0; i < NPP; ++i) pf(L.pp + k * NPP + i);
    if (k == N) return;
#pragma unroll
    for (int i = 0; i < NU * NX + NU; ++i) pf(L.ric + k * NRIC + i);
#pragma unroll
    for (int i = 0; i < NAB; ++i) pf(L.ab + sk(k) * NAB + i);
    if (coff >= 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) pf(coff + (k + 1) * NX + i);
    }
  }
  MPCV_D void load_fwd(int k, int coff, FwdIn& s) const {
#pragma unroll
    for (int i = 0; i < NPX; ++i) s.P[i] = ws[L.pp + k * NPP + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) s.p[i] = ws[L.pp + k * NPP + NPX + i];
    if (k == N) return;
    const int ao = L.ab + sk(k) * NAB, ro = L.ric + k * NRIC;
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) s.K[i] = ws[ro + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) s.kff[i] = ws[ro + NU * NX + i];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) s.A[i] = ws[ao + i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) s.B[i] = ws[ao + NX * NX + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) s.c[i] = (coff >= 0) ? ws[coff + (k + 1) * NX + i] : 0.0;
  }

  MPCV_DN void riccati_solve(int rmode, int coff) const {
    if (g.lane == 0) {
      double pv[NX];   // p_{k+1}
#pragma unroll
      for (int i = 0; i < NX; ++i) { pv[i] = rvar(rmode, ix(N, i)); ws[L.pp + N * NPP + NPX + i] = pv[i]; }
      BwdIn cur;
      for (int k = N - 1; k >= 0; --k) {
        load_bwd(k, rmode, coff, cur);
        if (k > 0) prefetch_bwd(k - 1, rmode, coff);
        double Pm[NX * NX], Pd[NX], gk[NU], t[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) { const double v = cur.P[tri(i, j)]; Pm[i * NX + j] = v; Pm[j * NX + i] = v; }
        }
        // Pd = P_{k+1} c_{k+1} + p_{k+1}
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = pv[i];
          if (coff >= 0) {
#pragma unroll
            for (int j = 0; j < NX; ++j) v += Pm[i * NX + j] * cur.c[j];
          }
          Pd[i] = v;
        }
        const int ro = L.ric + k * NRIC;
        // g = r_u + B' Pd ; kff = -F^{-1} g
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = cur.ru[i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += cur.B[j * NU + i] * Pd[j];
          if (bnd(iu(k, i)).fixed) v = 0.0;
          gk[i] = v;
        }
        double F[NU * NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) F[i * NU + j] = cur.F[tri(i, j)];
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = -gk[i];
#pragma unroll
          for (int l = 0; l < i; ++l) v -= F[i * NU + l] * t[l];
          t[i] = v * F[i * NU + i];
        }
#pragma unroll
        for (int i = NU - 1; i >= 0; --i) {
          double v = t[i];
#pragma unroll
          for (int l = i + 1; l < NU; ++l) v -= F[l * NU + i] * t[l];
          t[i] = v * F[i * NU + i];
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) ws[ro + NU * NX + i] = t[i];
        // p_k = r_x + A' Pd + K' g
        double pk[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = cur.rx[i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += cur.A[j * NX + i] * Pd[j];
#pragma unroll
          for (int j = 0; j < NU; ++j) v += cur.K[j * NX + i] * gk[j];
          pk[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { pv[i] = pk[i]; ws[L.pp + k * NPP + NPX + i] = pk[i]; }
      }
    }
    riccati_forward(coff);
  }
  // forward sweep: dx_0 = c_0, du = K dx + kff, dx+ = A dx + B du + c+, lam+_k = P_k dx_k + p_k
  MPCV_D void riccati_forward(int coff) const {
    if (LANES > 1) { riccati_forward_lanes(coff); return; }
    if (g.lane == 0) {
      double dx[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = (coff >= 0) ? ws[coff + i] : 0.0;
      FwdIn fc;
      for (int k = 0; k <= N; ++k) {
        load_fwd(k, coff, fc);
        if (k < N) prefetch_fwd(k + 1, coff);
        double Pm[NX * NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) { const double v = fc.P[tri(i, j)]; Pm[i * NX + j] = v; Pm[j * NX + i] = v; }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = fc.p[i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += Pm[i * NX + j] * dx[j];
          ws[L.lamp + k * NX + i] = v;
          ws[L.d + ix(k, i)] = dx[i];
        }
        if (k == N) break;
        double du[NU], dn[NX];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = fc.kff[i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += fc.K[i * NX + j] * dx[j];
          du[i] = v;
          ws[L.d + iu(k, i)] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = fc.c[i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += fc.A[i * NX + j] * dx[j];
#pragma unroll
          for (int j = 0; j < NU; ++j) v += fc.B[i * NU + j] * du[j];
          dn[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) dx[i] = dn[i];
      }
    }
    g.sync();
  }
  // Lane groups: only the state recursion dx_{k+1} = A dx + B (K dx + kff) + c is a chain; it runs on lane 0 with
  // nothing else on it.  The multipliers lam+_k = P_k dx_k + p_k are independent outputs once the dx are known and
  // are computed by all lanes afterwards.  Same expressions as the single-lane form: bit-identical.
  MPCV_D void riccati_forward_lanes(int coff) const {
    if (g.lane == 0) {
      double dx[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) { dx[i] = (coff >= 0) ? ws[coff + i] : 0.0; ws[L.d + ix(0, i)] = dx[i]; }
      for (int k = 0; k < N; ++k) {
        const WS ab = ws.view(L.ab + sk(k) * NAB), ric = ws.view(L.ric + k * NRIC);
        double du[NU], dn[NX];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = ric[NU * NX + i];
#pragma unroll
          for (int j = 0; j < NX; ++j) v += ric[i * NX + j] * dx[j];
          du[i] = v;
          ws[L.d + iu(k, i)] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = (coff >= 0) ? ws[coff + (k + 1) * NX + i] : 0.0;
#pragma unroll
          for (int j = 0; j < NX; ++j) v += ab[i * NX + j] * dx[j];
#pragma unroll
          for (int j = 0; j < NU; ++j) v += ab[NX * NX + i * NU + j] * du[j];
          dn[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { dx[i] = dn[i]; ws[L.d + ix(k + 1, i)] = dn[i]; }
      }
    }
    g.sync();
    for (int it = g.lane; it < (N + 1) * NX; it += LANES) {
      const int k = it / NX, i = it - k * NX;
      const WS pk = ws.view(L.pp + k * NPP);
      double v = pk[NPX + i];
#pragma unroll
      for (int j = 0; j < NX; ++j) v += pk[j <= i ? tri(i, j) : tri(j, i)] * ws[L.d + ix(k, j)];
      ws[L.lamp + it] = v;
    }
    g.sync();
  }

  // ---- single shooting: condensed Hessian, dense Cholesky ---------------------------------------------
  // H = sum_k S_k' W_k S_k,  S_k = d(x_k,u_k)/dU,  W_k with the costates as multipliers.
  MPCV_DN bool condensed_factor(double dw) {
    int ok = 1;
    const int nU = NU * N;
    if (g.lane == 0) {
      const int H = L.hred, Gm = L.gam, Tm = L.tmp;
      for (int i = 0; i < nU * (nU + 1) / 2; ++i) ws[H + i] = 0.0;
      for (int i = 0; i < NX * nU; ++i) ws[Gm + i] = 0.0;
      for (int k = 0; k < N; ++k) {
        const int ncol = NU * (k + 1);
        double W[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) W[i] = ws[L.hw + sk(k) * NW + i];
        auto Wf = [&](int a, int b) { return a >= b ? W[tri(a, b)] : W[tri(b, a)]; };
        // T = W S  (NZ x ncol); S rows 0..NX-1 = Gamma_k, rows NX.. = unit columns of stage k
        for (int c = 0; c < ncol; ++c) {
#pragma unroll
          for (int a = 0; a < NZ; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NX; ++b) v += Wf(a, b) * ws[Gm + b * nU + c];
            if (c >= NU * k) v += Wf(a, NX + c - NU * k);
            ws[Tm + a * nU + c] = v;
          }
        }
        for (int r = 0; r < ncol; ++r) {
          for (int c = 0; c <= r; ++c) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < NX; ++a) v += ws[Gm + a * nU + r] * ws[Tm + a * nU + c];
            if (r >= NU * k) v += ws[Tm + (NX + r - NU * k) * nU + c];
            ws[H + tri(r, c)] += v;
          }
        }
        // Gamma_{k+1} = A Gamma_k + B E_k
        const int ao = L.ab + sk(k) * NAB;
        for (int c = 0; c < ncol; ++c) {
          double col[NX], out[NX];
#pragma unroll
          for (int b = 0; b < NX; ++b) col[b] = ws[Gm + b * nU + c];
#pragma unroll
          for (int a = 0; a < NX; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NX; ++b) v += ws[ao + a * NX + b] * col[b];
            if (c >= NU * k) v += ws[ao + NX * NX + a * NU + (c - NU * k)];
            out[a] = v;
          }
#pragma unroll
          for (int a = 0; a < NX; ++a) ws[Gm + a * nU + c] = out[a];
        }
      }
      for (int i = 0; i < nU; ++i) {
        const Bnd b = bnd(i);
        if (b.fixed) {
          for (int j = 0; j < nU; ++j) ws[H + (j <= i ? tri(i, j) : tri(j, i))] = 0.0;
          ws[H + tri(i, i)] = 1.0;
        } else {
          double sg, r;
          sigma_r(i, &sg, &r);
          ws[H + tri(i, i)] += sg + dw;
        }
      }
      // packed Cholesky
      for (int j = 0; j < nU && ok; ++j) {
        double dj = ws[H + tri(j, j)];
        for (int l = 0; l < j; ++l) { const double t = ws[H + tri(j, l)]; dj -= t * t; }
        if (!(dj > 0.0) || !(dj < INFINITY)) { ok = 0; break; }
        dj = sqrt(dj);
        ws[H + tri(j, j)] = dj;
        for (int i = j + 1; i < nU; ++i) {
          double v = ws[H + tri(i, j)];
          for (int l = 0; l < j; ++l) v -= ws[H + tri(i, l)] * ws[H + tri(j, l)];
          ws[H + tri(i, j)] = v / dj;
        }
      }
    }
    ok = g.bcast(ok);
    g.sync();
    return ok != 0;
  }
  MPCV_DN void condensed_solve() const {
    const int nU = NU * N;
    if (g.lane == 0) {
      const int H = L.hred;
      for (int i = 0; i < nU; ++i) {
        double sg, r;
        sigma_r(i, &sg, &r);
        double v = bnd(i).fixed ? 0.0 : -r;
        for (int l = 0; l < i; ++l) v -= ws[H + tri(i, l)] * ws[L.d + l];
        ws[L.d + i] = v / ws[H + tri(i, i)];
      }
      for (int i = nU - 1; i >= 0; --i) {
        double v = ws[L.d + i];
        for (int l = i + 1; l < nU; ++l) v -= ws[H + tri(l, i)] * ws[L.d + l];
        ws[L.d + i] = v / ws[H + tri(i, i)];
      }
    }
    g.sync();
  }

  // ---- step-length helpers ----------------------------------------------------------------------------
  // Fraction-to-the-boundary rules.  alpha = min(1, min_i num_i / den_i) over the bounds the step moves towards
  // (num_i, den_i < 0).  The minimum is tracked as a (num, den) pair compared by cross-multiplication,
  // num_i den_best < num_best den_i, and divided once per lane at the end: no FP64 division (and no divergent
  // slow path of it) inside the loop; the value is the quotient of the minimising bound, as before.
  struct Ratio {
    double num = -1.0, den = -1.0;     // 1
    MPCV_D void take(double n, double d) { if (d < 0.0 && n * den < num * d) { num = n; den = d; } }
    MPCV_D double value() const { return num / den; }
  };
  MPCV_D double ftb_primal() const {
    Ratio r;
    lane_loop(L.n, [&](int i) { return V2{ws[L.w + i], ws[L.d + i]}; }, [&](int i, const V2& v) {
      const Bnd b = bnd(i);
      if (b.hasl) r.take(-tau * (v.a - b.lo), v.b);
      if (b.hasu) r.take(-tau * (b.hi - v.a), -v.b);
    });
    return g.min(r.value());
  }
  // fraction-to-the-boundary rule for the bound multipliers
  MPCV_D double ftb_dual() const {
    Ratio r;
    lane_loop(L.n, [&](int i) {
      const Bnd b = bnd(i);
      if (!bounded(b)) return V4{0.0, 0.0, 0.0, 0.0};
      return V4{ws[L.w + i], ws[L.d + i], zl_at(i, b), zu_at(i, b)};
    }, [&](int i, const V4& v) {
      const Bnd b = bnd(i);
      if (b.hasl) r.take(-tau * v.c, dz_of(v.a - b.lo, v.c, -v.b));
      if (b.hasu) r.take(-tau * v.d, dz_of(b.hi - v.a, v.d, v.b));
    });
    return g.min(r.value());
  }
  // dz = mu / s - z + (z / s) ds  with one reciprocal per bound (ds = -d for a lower, +d for an upper bound).
  // Evaluated twice per accepted step (step length, then update) rather than parked in the workspace: a
  // reciprocal is cheaper than a store and a load per bound, and the update then reads nothing but w, d, z.
  MPCV_D double dz_of(double slack, double z, double ds) const {
    const double inv = 1.0 / slack;
    return mu * inv - z + z * inv * ds;
  }
  // z reset into [mu / (kappa_Sigma s), kappa_Sigma mu / s], kappa_Sigma = 1e10 (a safeguard that almost
  // never binds: test on z*s, divide only when it does)
  MPCV_D double clamp_z(double z, double s) const {
    const double zs = z * s;
    if (zs > 1e10 * mu) return 1e10 * mu / s;
    if (zs < mu / 1e10) return mu / (1e10 * s);
    return z;
  }

  // mirror blocked controls from their predecessor so outputs read like MPCTools' u trajectory
  MPCV_D void sync_blocked() const {
    if (!(Model::HAS_UPREV && P.ntu > 0)) return;
    if (g.lane == 0)
      for (int k = (P.ntu > 1 ? P.ntu : 1); k < N; ++k) ws[L.w + iu(k, 0)] = ws[L.w + iu(k - 1, 0)];
    g.sync();
  }

  // ---- the solve, phase by phase -----------------------------------------------------------------------
  // The same phase functions are scheduled two ways: solve() runs them back to back inside one
  // kernel (thread / warp layouts, closed loop); the phase-kernel pipeline (mpcv_phase.cuh) runs
  // each phase as its own batch-wide launch with the scalar state parked in ws[L.st..].
  //
  // On entry to start() ws[L.w..] holds the starting point and ws[L.par..] the parameters.

  // push the starting point into the interior (bound_push / bound_frac); fixed variables take
  // their value; z0 = 1, lam0 = 0
  MPCV_DN void start() {
    for (int i = g.lane; i < L.n; i += LANES) {
      const Bnd b = bnd(i);
      double v = ws[L.w + i];
      if (b.fixed && !var_is_blocked_u(i)) v = lbx[i];
      if (b.hasl && b.hasu) {
        const double pl = fmin(P.bound_push * fmax(1.0, fabs(b.lo)), P.bound_frac * (b.hi - b.lo));
        const double pu = fmin(P.bound_push * fmax(1.0, fabs(b.hi)), P.bound_frac * (b.hi - b.lo));
        v = fmin(fmax(v, b.lo + pl), b.hi - pu);
      } else if (b.hasl) {
        v = fmax(v, b.lo + P.bound_push * fmax(1.0, fabs(b.lo)));
      } else if (b.hasu) {
        v = fmin(v, b.hi - P.bound_push * fmax(1.0, fabs(b.hi)));
      }
      ws[L.w + i] = v;
      ws[L.zl + i] = b.hasl ? 1.0 : 0.0;
      ws[L.zu + i] = b.hasu ? 1.0 : 0.0;
    }
    for (int i = g.lane; i < NX * (N + 1); i += LANES) ws[L.lam + i] = 0.0;
    g.sync();
    sync_blocked();
    df = 1.0;
    mu = P.mu_init;
    tau = fmax(0.99, 1.0 - mu);
    nfil = 0;
    theta_max = theta_min = -1.0;
    delta_w_last = 0.0;
    iter = 0;
    acc_count = 0;
    f_last = -1e50;            // IPOPT: last_obj_val_ starts at -1e50
    lg_valid = false;
  }

  // Precondition: eval_derivatives(false) at df = 1.  Gradient-based objective scaling
  // (nlp_scaling_max_gradient) and least-squares multipliers  [I J'; J 0][.; lam] = -[grad f - zl + zu; 0]
  MPCV_DN void init_scaling_and_multipliers() {
    double gmax = 0.0;
    for (int i = g.lane; i < L.n; i += LANES)
      if (!bnd(i).fixed) gmax = fmax(gmax, fabs(ws[L.grad + i]));
    gmax = g.max(gmax);
    if (gmax > P.scal_max_grad) {
      df = fmax(P.scal_max_grad / gmax, 1e-8);
      if (SINGLE) {
        eval_derivatives(false);
      } else {
        // grad and f are linear in df and were evaluated with df = 1: rescaling is bit-identical
        // to a second derivative sweep
        for (int i = g.lane; i < L.n; i += LANES) ws[L.grad + i] = df * ws[L.grad + i];
        f_curr = df * f_curr;
        g.sync();
      }
    }
    if (!SINGLE) {
      if (riccati_factor(0.0, true)) {
        riccati_solve(1, -1);
        double lmax = 0.0;
        for (int i = g.lane; i < L.m; i += LANES) lmax = fmax(lmax, fabs(ws[L.lamp + i]));
        lmax = g.max(lmax);
        if (lmax <= 1e3) {
          for (int i = g.lane; i < L.m; i += LANES) ws[L.lam + i] = ws[L.lamp + i];
        }
        g.sync();
      }
    }
  }

  // convergence test (scaled E_0 <= tol plus the unscaled caps) and monotone barrier update;
  // returns a final IPOPT status or kRunning
  MPCV_DN int check_convergence_update_mu() {
    const Err e = errors();
    const double E0 = Emu(e, 0.0);
    {
      const double dual_u = e.dual * e.sd / df, compl_u = compl_err(e, 0.0) * e.sc / df;
      if (E0 <= P.tol && dual_u <= P.dual_inf_tol && e.prim <= P.constr_viol_tol && compl_u <= P.compl_inf_tol)
        return MPCV_SOLVE_SUCCEEDED;
      // acceptable level (IpOptErrorConvCheck::CurrentIsAcceptable): acceptable_iter consecutive iterates within
      // acceptable_tol (and IPOPT's fixed acceptable_dual_inf / constr_viol / compl_inf tolerances 1e10 / 1e-2 / 1e-2)
      // whose objective moved by no more than acceptable_obj_change_tol.  The scripts set acceptable_tol = tol
      // (single_shooting_v1.py:121-129), so the regular test fires first there.
      const bool acc = P.acceptable_iter > 0 && E0 <= P.acceptable_tol && dual_u <= 1e10 && e.prim <= 1e-2 &&
                       compl_u <= 1e-2 &&
                       fabs(f_curr - f_last) / fmax(1.0, fabs(f_curr)) <= P.acceptable_obj_change_tol;
      f_last = f_curr;
      if (acc) { if (++acc_count >= P.acceptable_iter) return MPCV_SOLVED_TO_ACCEPTABLE_LEVEL; }
      else acc_count = 0;
    }
    if (iter >= P.max_iter) return MPCV_MAXIMUM_ITERATIONS_EXCEEDED;
    if (!(E0 < INFINITY)) return MPCV_INVALID_NUMBER_DETECTED;
    bool done = false;
    while (!done && Emu(e, mu) <= 10.0 * mu) {
      double new_mu = fmin(0.2 * mu, pow(mu, 1.5));
      new_mu = fmax(new_mu, fmin(P.tol, P.compl_inf_tol * df) / 11.0);
      const bool changed = new_mu != mu;
      mu = new_mu;
      tau = fmax(0.99, 1.0 - mu);
      if (!changed) done = true; else nfil = 0;
    }
    return kRunning;
  }

  // search direction with inertia correction, maximal primal step and the merit-function terms
  // of the current iterate.  direction_first() tries delta_w = 0; direction_retry() walks IPOPT's
  // delta_w schedule after a failed first attempt (the phase pipeline runs the retries as a separate,
  // compacted launch so that warps without wrong inertia do not idle through them).
  // factorisation of the Newton system with the backward vector recursion fused in (multiple shooting)
  MPCV_D bool factor_dir(double dw) {
    if (SINGLE) return condensed_factor(dw);
    return riccati_factor_x<true>(dw, false, 0, L.c);
  }
  MPCV_DN bool direction_first() {
    prepare_barrier();
    const bool ok = factor_dir(0.0);
    if (ok) direction_finish(0.0);
    return ok;
  }
  // IPOPT's delta_w schedule: first trial 1e-4 (or a third of the last successful value), then x100
  // (first correction ever / far above the last value) or x8
  MPCV_D double next_delta_w(double dw) const {
    if (dw == 0.0) return (delta_w_last == 0.0) ? 1e-4 : fmax(1e-20, delta_w_last / 3.0);
    return dw * ((delta_w_last == 0.0 || 1e5 * delta_w_last < dw) ? 100.0 : 8.0);
  }
  // returns 0 or MPCV_ERROR_IN_STEP_COMPUTATION
  MPCV_DN int direction_retry() {
    double dw = 0.0;
    bool ok = false;
    while (!ok) {
      dw = next_delta_w(dw);
      if (dw > 1e20) break;
      ok = factor_dir(dw);
    }
    if (!ok) return MPCV_ERROR_IN_STEP_COMPUTATION;
    direction_finish(dw);
    return 0;
  }
  MPCV_D void direction_finish(double dw) {
    if (dw > 0.0) delta_w_last = dw;
    if (SINGLE) condensed_solve(); else riccati_forward(L.c);
    direction_post();
  }
  // maximal primal step (fraction to the boundary) and the merit-function terms of the current iterate
  MPCV_D void direction_post() {
    ls_alpha_max = ftb_primal();
    double theta = 0.0, gBD = 0.0, lg = 0.0;
    lane_loop(L.m, [&](int i) { return V1{ws[L.c + i]}; }, [&](int, const V1& v) { theta += fabs(v.a); });
    const bool need_lg = !lg_valid;
    LogSum ls;
    lane_loop(L.n, [&](int i) { return V3{ws[L.rb + i], ws[L.d + i], ws[L.w + i]}; }, [&](int i, const V3& v) {
      const Bnd b = bnd(i);
      if (!b.fixed) gBD += v.a * v.b;
      if (need_lg) ls.add_var(b, v.c);
    });
    theta = g.sum(theta); gBD = g.sum(gBD);
    lg = need_lg ? g.sum(ls.value()) : lg_curr;
    const double dmp = need_lg ? g.sum(ls.damp) : dmp_curr;
    ls_theta = theta;
    ls_gBD = gBD;
    ls_phi = phi_of(f_curr, lg, dmp);
    if (theta_max < 0.0) { theta_max = 1e4 * fmax(1.0, theta); theta_min = 1e-4 * fmax(1.0, theta); }
  }
  MPCV_D int compute_direction() { return direction_first() ? 0 : direction_retry(); }

  // ---- filter line search --------------------------------------------------------------------------------
  // switching-condition powers of the current iterate, evaluated once per line search
  struct LsPow { double g23, t11; };
  MPCV_D LsPow ls_pows() const {
    LsPow p;
    p.g23 = ls_gBD < 0.0 ? pow(-ls_gBD, 2.3) : 0.0;
    p.t11 = pow(ls_theta, 1.1);
    return p;
  }
  MPCV_D bool ls_is_ftype(double a, const LsPow& pw) const {
    const double eps = 2.220446049250313e-16;
    if (ls_theta == 0.0 && ls_gBD > 0.0 && ls_gBD < 100.0 * eps) return true;
    return ls_gBD < 0.0 && a * pw.g23 > pw.t11;
  }
  MPCV_D bool ls_armijo(double a, double phi_t) const { return compare_le(phi_t - ls_phi, 1e-8 * a * ls_gBD, ls_phi); }
  MPCV_D bool ls_acceptable(double a, double phi_t, double theta_t, const LsPow& pw) const {
    const double theta = ls_theta, phi = ls_phi;
    if (!(phi_t < INFINITY) || !(theta_t < INFINITY) || phi_t != phi_t) return false;
    if (theta_max > 0.0 && theta_t > theta_max) return false;
    bool acc;
    if (a > 0.0 && ls_is_ftype(a, pw) && theta <= theta_min) acc = ls_armijo(a, phi_t);
    else {
      if (phi_t > phi) {
        double basval = 1.0;
        if (fabs(phi) > 10.0) basval = log10(fabs(phi));
        if (log10(phi_t - phi) > 5.0 + basval) return false;
      }
      acc = compare_le(theta_t, (1.0 - 1e-5) * theta, theta) || compare_le(phi_t - phi, -1e-8 * theta, phi);
    }
    if (!acc) return false;
    for (int q = 0; q < nfil; ++q) {
      const double fp = fil_phi(q), ft = fil_th(q);
      if (!(compare_le(phi_t, fp, fp) || compare_le(theta_t, ft, ft))) return false;
    }
    return true;
  }
  // filter augmentation, dual step length and the update of (w, z, lam) for an accepted step
  MPCV_D void ls_accept_step(double alpha, double alpha_test, double phi_acc, const LsPow& pw) {
    ls_filter_augment(alpha_test, phi_acc, pw);
    ls_take_step(alpha);
  }
  // (the two halves are separate so that a lane-group kernel can bring its warp back together between the
  // branchy scalar part and the loops over the variables, see ph_accept_kernel)
  MPCV_D void ls_filter_augment(double alpha_test, double phi_acc, const LsPow& pw) {
    const double theta = ls_theta, phi = ls_phi;
    if (!ls_is_ftype(alpha_test, pw) || !ls_armijo(alpha_test, phi_acc)) {
      filter_add(phi - 1e-8 * theta, (1.0 - 1e-5) * theta);
    }
  }
  MPCV_D void ls_take_step(double alpha) {
    const double alpha_dual = ftb_dual();
    lane_loop(L.n, [&](int i) { const Bnd b = bnd(i); return V4{ws[L.w + i], ws[L.d + i], zl_at(i, b), zu_at(i, b)}; },
              [&](int i, const V4& v) {
      const Bnd b = bnd(i);
      const double wi = v.a + alpha * v.b;
      ws[L.w + i] = wi;
      if (b.hasl) ws[L.zl + i] = clamp_z(v.c + alpha_dual * dz_of(v.a - b.lo, v.c, -v.b), wi - b.lo);
      if (b.hasu) ws[L.zu + i] = clamp_z(v.d + alpha_dual * dz_of(b.hi - v.a, v.d, v.b), b.hi - wi);
    });
    if (!SINGLE)
      lane_loop(L.m, [&](int i) { return V2{ws[L.lam + i], ws[L.lamp + i]}; },
                [&](int i, const V2& v) { ws[L.lam + i] = v.a + alpha * (v.b - v.a); });
    lg_curr = lg_trial;          // the accepting trial evaluation was the last one
    dmp_curr = dmp_trial;
    lg_valid = true;
    g.sync();
    sync_blocked();
    ++iter;
  }

  // Fast path of the phase pipeline: the full step (alpha = ls_alpha_max, stage part already evaluated by
  // trial_stage) is accepted by the filter — no backtracking, no second-order correction.  Returns false
  // without touching the iterate when the slow path (line_search(true)) has to take over.
  struct LsFirst { LsPow pw; double phi_t; bool ok; };
  MPCV_D LsFirst line_search_first_decide() {
    LsFirst r;
    r.pw = ls_pows();
    double f_t, theta_t;
    trial_reduce(ls_alpha_max, L.d, &f_t, &theta_t, &r.phi_t);
    r.ok = ls_acceptable(ls_alpha_max, r.phi_t, theta_t, r.pw);
    return r;
  }
  MPCV_DN bool line_search_first() {
    const LsFirst r = line_search_first_decide();
    if (!r.ok) return false;
    ls_accept_step(ls_alpha_max, ls_alpha_max, r.phi_t, r.pw);
    return true;
  }

  // filter line search with second-order correction, then acceptance of the trial point
  // (w, z, lam updated in place).  have_trial0: the stage part of the first trial (alpha =
  // ls_alpha_max) was already evaluated by trial_stage().  Returns 0 or MPCV_RESTORATION_FAILED.
  MPCV_DN int line_search(bool have_trial0) {
    const LsPow pw = ls_pows();
    const double alpha_max = ls_alpha_max, theta = ls_theta, gBD = ls_gBD;
    double alpha_min = 1e-5;
    if (gBD < 0.0) {
      alpha_min = fmin(1e-5, 1e-8 * theta / (-gBD));
      if (theta <= theta_min) alpha_min = fmin(alpha_min, pw.t11 / pw.g23);
    }
    alpha_min *= 0.05;
    double alpha = alpha_max, alpha_test = alpha_max, phi_acc = 0.0;
    bool accepted = false;
    int nsteps = 0;
    while (alpha > alpha_min || nsteps == 0) {
      double f_t, theta_t, phi_t;
      if (!SINGLE && have_trial0 && nsteps == 0) trial_reduce(alpha, L.d, &f_t, &theta_t, &phi_t);
      else eval_trial(alpha, L.d, !SINGLE && nsteps == 0 && P.max_soc > 0, &f_t, &theta_t, &phi_t);
      alpha_test = alpha;
      if (ls_acceptable(alpha, phi_t, theta_t, pw)) { accepted = true; phi_acc = phi_t; break; }
      // second-order correction on the first rejected trial
      if (!SINGLE && nsteps == 0 && P.max_soc > 0 && theta_t >= theta) {
        // c_soc accumulates in ct: c_soc = ct + alpha_soc * c_soc (c_soc starts as c)
        double theta_soc_old = 0.0, theta_tr = theta_t, alpha_soc = alpha;
        int count = 0;
        bool acc_soc = false;
        // save the plain direction so that a failed SOC can fall back to backtracking
        save_step();
        for (int i = g.lane; i < L.m; i += LANES) ws[L.ct + i] = ws[L.ct + i] + alpha_soc * ws[L.c + i];
        g.sync();
        while (count < P.max_soc && !acc_soc && (count == 0 || theta_tr <= 0.99 * theta_soc_old)) {
          theta_soc_old = theta_tr;
          riccati_solve(0, L.ct);
          alpha_soc = ftb_primal();
          double f_s, theta_s, phi_s;
          eval_trial_soc(alpha_soc, &f_s, &theta_s, &phi_s);
          if (ls_acceptable(alpha, phi_s, theta_s, pw)) { acc_soc = true; alpha = alpha_soc; phi_acc = phi_s; }
          else { ++count; theta_tr = theta_s; }
        }
        if (acc_soc) { accepted = true; break; }
        restore_step();
      }
      alpha *= 0.5;
      ++nsteps;
    }
    if (!accepted) return SINGLE ? (int)MPCV_RESTORATION_FAILED : restoration();
    ls_accept_step(alpha, alpha_test, phi_acc, pw);
    return 0;
  }

  // ---- feasibility restoration ----------------------------------------------------------------------------
  // IPOPT switches to its restoration phase when the line search cannot find an acceptable step (alpha < alpha_min):
  // it leaves the current point in the filter, reduces the constraint violation until the point is acceptable to the
  // filter again with theta <= kappa_resto theta_R (kappa_resto = 0.9), resets the equality multipliers
  // (constr_mult_reset_threshold = 0) and resumes.  IPOPT minimises rho |c|_1 + zeta/2 |D_R (x - x_R)|^2 by another
  // interior-point solve; here the phase is a damped Gauss-Newton iteration on theta: the scaled minimum-norm step
  // d = argmin |d|^2 + sum_i (d_i / s_i)^2  s.t.  J d = -c  (the identity-Hessian Riccati system the multiplier
  // initialisation uses, plus an affine scaling by the bound slacks s_i) with the fraction-to-the-boundary rule and an
  // Armijo test on theta.  Every restoration step counts as an iteration, like
  // IPOPT's.  Returns 0 with the new iterate in place (derivatives evaluated), or MPCV_RESTORATION_FAILED.
  static constexpr int kRestoMaxIter = 40, kRestoMaxBacktrack = 30;
  MPCV_D bool filter_ok(double phi_t, double theta_t) const {
    if (!(phi_t < INFINITY) || !(theta_t < INFINITY)) return false;
    if (theta_max > 0.0 && theta_t > theta_max) return false;
    for (int q = 0; q < nfil; ++q) {
      const double fp = fil_phi(q), ft = fil_th(q);
      if (!(compare_le(phi_t, fp, fp) || compare_le(theta_t, ft, ft))) return false;
    }
    return true;
  }
  MPCV_DN int restoration() {
    const double theta_R = ls_theta, phi_R = ls_phi;
    if (!(theta_R > P.tol)) return MPCV_RESTORATION_FAILED;          // (almost) feasible: nothing to restore
    // the point we leave goes into the filter
    filter_add(phi_R - 1e-8 * theta_R, (1.0 - 1e-5) * theta_R);
    double theta = theta_R;
    bool done = false;
    for (int r = 0; r < kRestoMaxIter && !done; ++r) {
      // affine scaling: the step is measured in |d|^2 + sum_i (d_i / s_i)^2 over the bound slacks s_i, so that a variable
      // sitting at a bound hardly moves and the fraction-to-the-boundary rule does not choke the step
      for (int i = g.lane; i < L.n; i += LANES) {
        const Bnd b = bnd(i);
        double sg = 0.0;
        if (b.hasl) { const double sl = ws[L.w + i] - b.lo; sg += 1.0 / (sl * sl); }
        if (b.hasu) { const double su = b.hi - ws[L.w + i]; sg += 1.0 / (su * su); }
        ws[L.sig + i] = sg;
      }
      g.sync();
      resto_sigma = true;
      const bool fok = riccati_factor(0.0, true);
      resto_sigma = false;
      if (!fok) return MPCV_RESTORATION_FAILED;
      riccati_solve(2, L.c);
      double alpha = ftb_primal();
      double f_t = 0.0, theta_t = 0.0, phi_t = 0.0;
      bool found = false;
      for (int j = 0; j < kRestoMaxBacktrack && !found; ++j) {
        eval_trial(alpha, L.d, false, &f_t, &theta_t, &phi_t);
        if (phi_t < INFINITY && theta_t <= (1.0 - 1e-4 * alpha) * theta) found = true;
        else alpha *= 0.5;
      }
      if (!found) return MPCV_RESTORATION_FAILED;
      for (int i = g.lane; i < L.n; i += LANES) ws[L.w + i] = ws[L.w + i] + alpha * ws[L.d + i];
      g.sync();
      sync_blocked();
      ++iter;
      theta = theta_t;
      eval_derivatives(true);
      if (theta_t <= 0.9 * theta_R && filter_ok(phi_t, theta_t)) done = true;
      if (iter >= P.max_iter) break;
    }
    if (!done) return MPCV_RESTORATION_FAILED;
    // back to the regular algorithm: equality multipliers reset, bound multipliers kept inside their safeguard band
    for (int i = g.lane; i < L.m; i += LANES) ws[L.lam + i] = 0.0;
    for (int i = g.lane; i < L.n; i += LANES) {
      const Bnd b = bnd(i);
      if (b.hasl) ws[L.zl + i] = clamp_z(ws[L.zl + i], ws[L.w + i] - b.lo);
      if (b.hasu) ws[L.zu + i] = clamp_z(ws[L.zu + i], b.hi - ws[L.w + i]);
    }
    g.sync();
    eval_derivatives(true);      // the Hessian of the Lagrangian depends on the multipliers
    lg_valid = false;
    return 0;
  }

  MPCV_DN SolveInfo solve() {
    SolveInfo info;
    start();
    eval_derivatives(false);
    init_scaling_and_multipliers();
    eval_derivatives(true);
    for (;;) {
      int st = check_convergence_update_mu();
      if (st != kRunning) { info.status = st; break; }
      st = compute_direction();
      if (st == 0) st = line_search(false);
      if (st != 0) { info.status = st; break; }
      eval_derivatives(true);
    }
    info.iters = iter;
    info.f = f_curr / df;
    info.df = df;
    return info;
  }
  MPCV_D SolveInfo finish(int status) const {
    SolveInfo info;
    info.status = status;
    info.iters = iter;
    info.f = f_curr / df;
    info.df = df;
    return info;
  }

  // SOC bookkeeping: the plain step (d, lamp) is parked in the (dead until the next iteration)
  // Riccati gain-free region?  No region is dead: K, P are needed for the re-solve.  Park it in
  // grad?  grad is live (rhs).  -> dedicated scratch: reuse `c`-sized `lamp` is live too.  We park in
  // the Hessian block storage, which is not read again until the next factorisation.
  MPCV_D void save_step() const {
    for (int i = g.lane; i < L.n; i += LANES) ws[L.park + i] = ws[L.d + i];
    for (int i = g.lane; i < L.m; i += LANES) ws[L.park + L.n + i] = ws[L.lamp + i];
    g.sync();
  }
  MPCV_D void restore_step() const {
    for (int i = g.lane; i < L.n; i += LANES) ws[L.d + i] = ws[L.park + i];
    for (int i = g.lane; i < L.m; i += LANES) ws[L.lamp + i] = ws[L.park + L.n + i];
    g.sync();
  }
  // trial evaluation for the SOC loop: c_soc <- c(x + alpha_soc d_soc) + alpha_soc * c_soc, in place in ct
  MPCV_DN void eval_trial_soc(double alpha_soc, double* f_out, double* theta_out, double* phi_out) const {
    double fpart = 0.0, thpart = 0.0, logpart = 0.0;
    bool bad = false;
    for (int k = g.lane; k < N; k += LANES) {
      double x[NX], u[NU], xn[NX], q;
      load_xu(k, L.w, alpha_soc, L.d, x, u);
      Model::val(P, x, u, pg(), ps(k), xn, &q);
      fpart += q;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const double r = xn[i] - (ws[L.w + ix(k + 1, i)] + alpha_soc * ws[L.d + ix(k + 1, i)]);
        thpart += fabs(r);
        ws[L.ct + (k + 1) * NX + i] = r + alpha_soc * ws[L.ct + (k + 1) * NX + i];
      }
    }
    if (g.lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const double r = ws[L.par + i] - (ws[L.w + ix(0, i)] + alpha_soc * ws[L.d + ix(0, i)]);
        thpart += fabs(r);
        ws[L.ct + i] = r + alpha_soc * ws[L.ct + i];
      }
    }
    LogSum ls;
    for (int i = g.lane; i < L.n; i += LANES) {
      const Bnd b = bnd(i);
      if (b.hasl || b.hasu) {
        if (!ls.add_var(b, ws[L.w + i] + alpha_soc * ws[L.d + i])) bad = true;
      }
    }
    logpart = ls.value();
    const double f = df * g.sum(fpart);
    const double lg = g.sum(logpart);
    lg_trial = lg;
    dmp_trial = g.sum(ls.damp);
    const bool anybad = g.max(bad ? 1.0 : 0.0) > 0.0;
    *f_out = f;
    *theta_out = g.sum(thpart);
    *phi_out = anybad ? INFINITY : phi_of(f, lg, dmp_trial);
    g.sync();
  }

  // ---- export (projection into the ORIGINAL bounds: honor_original_bounds=yes) -------------------------
  MPCV_D double projected(int i) const {
    double v = ws[L.w + i];
    const double l = lbx[i], u = ubx[i];
    if (l > -kInfBound) v = fmax(v, l);
    if (u < kInfBound) v = fmin(v, u);
    return v;
  }
};

}  // namespace mpcv
