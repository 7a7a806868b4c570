// mpcv_models.cuh — per-model device dynamics, integrators and hand-derived sweeps.
//
// Each model provides, for ONE shooting interval x+ = phi(x,u;p), q = interval cost:
//   val(): value only                      (line-search trial points, rollouts, plant step)
//   der(): value + A = dphi/dx, B = dphi/du, cost gradient and the exact Hessian of
//          df*q + lam' * phi w.r.t. z = (x,u)   (the interior-point Newton system)
// All derivatives are hand-derived forward sweeps over the RK4 stages (no AD, no symbolic
// graph); the adjoint (reverse) sweep over the horizon lives in mpcv_ipm.cuh (Riccati /
// costate recursions).  tests/ check every der() against the oracle's generic AD.
//
// FP64 only.  The file also compiles as plain C++ (MPCV_HD empty) for the CPU-side unit
// harness in tests/hostsim, which exists to debug without a GPU and is never shipped.
#pragma once

#include <math.h>

#include "../../include/mpcv.h"

#if defined(__CUDACC__)
#define MPCV_HD __host__ __device__ __forceinline__
#define MPCV_D __device__ __forceinline__
// heavy phases of the interior-point iteration are real calls: one copy of each in the
// instruction stream keeps the kernel inside the instruction cache
#if defined(MPCV_INLINE_PHASES)
// phase-kernel translation units: every kernel holds one or two phases, inline them
#define MPCV_DN __device__ __forceinline__
#else
#define MPCV_DN __device__ __noinline__
#endif
#else
#define MPCV_HD inline
#define MPCV_D inline
#define MPCV_DN
#endif

namespace mpcv {

// Problem constants broadcast to every thread (kernel argument, lives in constant bank).
struct Params {
  int N, M, ntu;
  double T;
  double Q[4], R[2], R1;
  double extra[4];
  // IPOPT options
  double tol, mu_init, bound_push, bound_frac, bound_relax, scal_max_grad;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol;
  double acceptable_tol, acceptable_obj_change_tol;
  int max_iter, max_soc, acceptable_iter;
  unsigned long long* diag;   // optional diagnostic counters (device memory in the kernels): [0] filter overflows
};

MPCV_HD void sincos_(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(a, s, c);
#else
  *s = sin(a);
  *c = cos(a);
#endif
}

// (sin, cos)(a + d) from (sin, cos)(a) and (sin, cos)(d).  The stage headings of an RK4 interval are th + j (h/2) w,
// j = 0..2M: two sincos and 2M rotations instead of 2M + 1 sincos (FP64 sincos is a ~40-instruction software
// sequence; it was half of the unicycle's instruction count).  Each rotation adds <= 2 ulp: <= 2e-15 after the 8 of
// M = 4, far inside the 1e-12 RK4 budget of SURVEY 8c.
MPCV_HD void rot_(double s, double c, double sd, double cd, double* s1, double* c1) {
  *s1 = s * cd + c * sd;
  *c1 = c * cd - s * sd;
}

MPCV_HD double rsqrt_(double a) {
#if defined(__CUDA_ARCH__)
  return rsqrt(a);
#else
  return 1.0 / sqrt(a);
#endif
}

// packed lower-triangular index of a symmetric matrix, i >= j
MPCV_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ---------------------------------------------------------------------------------------
// Unicycle  xdot = v cos(th), ydot = v sin(th), thdot = w
//   Casadi/single_shooting_v1.py:70-81, Casadi/multiple_shooting_casadi.py:68-73
//
// Structure used by the hand derivation: with u = (v,w) constant over the interval the
// heading at every RK4 stage is th + tau*w exactly, so
//   phi_x = x + v*C0,  phi_y = y + v*S0,  phi_th = th + T*w,
//   C0 = sum_i g_i cos(th + tau_i w),  S0 = sum_i g_i sin(th + tau_i w)
// (Simpson weights g = h/6*(1,4,1) per sub-step: k2 and k3 see the same heading), and every
// RK4 stage position is x + v*(partial trig sum).  Moments C1,S1 (weights g_i tau_i) and
// C2,S2 (g_i tau_i^2) give all w-derivatives.
//
// KIND 0: RK4(M) + cost quadrature  (multiple_shooting_casadi.py:98-114, single_shooting_v2.py:97-113)
// KIND 1: forward Euler + node cost (single_shooting_v1.py:85-91, :100-105)
// KIND 2: RK4(M) + node cost against per-stage references (MPCTools: getCasadiFunc(rk4=True),
//         Trajectory_tracking.py:51-61, mpctools/multiple_shooting_mpctools.py:48-55)
// ---------------------------------------------------------------------------------------
template <int KIND>
struct Unicycle {
  static constexpr int NX = 3, NU = 2, NZ = 5;
  static constexpr int NPG = (KIND == 2) ? 0 : 3;
  static constexpr int NPS = (KIND == 2) ? 5 : 0;
  static constexpr bool HAS_UPREV = false;
  static constexpr bool LTI = false;
  static constexpr int DER_MINB = 4;      // CTAs per SM of the phase pipeline's derivative sweep (register cap 128)
  static constexpr int FAC_MINB = 4;      // ... and of its factor kernel
  static constexpr int MODEL_ID = KIND == 0 ? MPCV_MODEL_UNICYCLE_RK4_QUAD
                                : KIND == 1 ? MPCV_MODEL_UNICYCLE_EULER_NODE
                                            : MPCV_MODEL_UNICYCLE_RK4_NODE;

  struct Sums { double C0, S0, C1, S1, C2, S2; };

  // ---- value only ---------------------------------------------------------------------
  template <class PG, class PS>
  MPCV_HD static void val(const Params& P, const double* x, const double* u, PG pg, PS ps,
                          double* xn, double* q) {
    const double v = u[0], w = u[1], th = x[2];
    if (KIND == 1) {
      double s, c;
      sincos_(th, &s, &c);
      xn[0] = x[0] + v * c * P.T;
      xn[1] = x[1] + v * s * P.T;
      xn[2] = th + w * P.T;
      const double e0 = x[0] - pg[0], e1 = x[1] - pg[1], e2 = th - pg[2];
      *q = P.Q[0] * e0 * e0 + P.Q[1] * e1 * e1 + P.Q[2] * e2 * e2 + P.R[0] * v * v + P.R[1] * w * w;
      return;
    }
    const int M = P.M;
    const double h = P.T / M, hh = 0.5 * h, h6 = h / 6.0;
    double X = x[0], Y = x[1];
    double qa = 0.0;
    double s0, c0, sd, cd;
    sincos_(th, &s0, &c0);
    sincos_(hh * w, &sd, &cd);      // every stage heading is th + j (h/2) w: one rotation per half step (see rot_)
    double xr = 0, yr = 0, tr = 0;
    if (KIND == 0) { xr = pg[0]; yr = pg[1]; tr = pg[2]; }
    const double uc = P.R[0] * v * v + P.R[1] * w * w;
    for (int j = 0; j < M; ++j) {
      const double a0 = th + (2 * j) * hh * w, a1 = th + (2 * j + 1) * hh * w, a2 = th + (2 * j + 2) * hh * w;
      double s1, c1, s2, c2;
      rot_(s0, c0, sd, cd, &s1, &c1);
      rot_(s1, c1, sd, cd, &s2, &c2);
      if (KIND == 0) {
        // L at the four RK4 stage states (k1_q..k4_q of MS:106-112)
        double ex, ey, et, L1, L2, L3, L4;
        ex = X - xr; ey = Y - yr; et = a0 - tr;
        L1 = P.Q[0] * ex * ex + P.Q[1] * ey * ey + P.Q[2] * et * et + uc;
        ex = X + hh * v * c0 - xr; ey = Y + hh * v * s0 - yr; et = a1 - tr;
        L2 = P.Q[0] * ex * ex + P.Q[1] * ey * ey + P.Q[2] * et * et + uc;
        ex = X + hh * v * c1 - xr; ey = Y + hh * v * s1 - yr;
        L3 = P.Q[0] * ex * ex + P.Q[1] * ey * ey + P.Q[2] * et * et + uc;
        ex = X + h * v * c1 - xr; ey = Y + h * v * s1 - yr; et = a2 - tr;
        L4 = P.Q[0] * ex * ex + P.Q[1] * ey * ey + P.Q[2] * et * et + uc;
        qa += h6 * (L1 + 2.0 * L2 + 2.0 * L3 + L4);
      }
      X += h6 * v * (c0 + 4.0 * c1 + c2);
      Y += h6 * v * (s0 + 4.0 * s1 + s2);
      s0 = s2; c0 = c2;
      (void)a0;
    }
    xn[0] = X; xn[1] = Y; xn[2] = th + P.T * w;
    if (KIND == 2) {
      const double e0 = x[0] - ps[0], e1 = x[1] - ps[1], e2 = th - ps[2], d0 = v - ps[3], d1 = w - ps[4];
      qa = P.Q[0] * e0 * e0 + P.Q[1] * e1 * e1 + P.Q[2] * e2 * e2 + P.R[0] * d0 * d0 + P.R[1] * d1 * d1;
    }
    *q = qa;
  }

  // one quadrature point: cost, gradient and Hessian contributions of
  //   wq * ( Qx (ex + v C0)^2 + Qy (ey + v S0)^2 + Qt (th + tau w - tr)^2 )
  MPCV_HD static void quad_point(const Params& P, double wq, double ex, double ey, double et, double tau,
                                 double v, const Sums& s, double* q, double* g, double* H) {
    // x-residual
    {
      const double r = ex + v * s.C0;
      const double gt = -v * s.S0, gv = s.C0, gw = -v * s.S1;    // dr/d(th, v, w); dr/dx = 1
      const double k = 2.0 * wq * P.Q[0];
      const double kr = k * r;
      *q += wq * P.Q[0] * r * r;
      g[0] += kr; g[2] += kr * gt; g[3] += kr * gv; g[4] += kr * gw;
      H[tri(0, 0)] += k;
      H[tri(2, 0)] += k * gt; H[tri(3, 0)] += k * gv; H[tri(4, 0)] += k * gw;
      H[tri(2, 2)] += k * gt * gt + kr * (-v * s.C0);
      H[tri(3, 2)] += k * gv * gt + kr * (-s.S0);
      H[tri(4, 2)] += k * gw * gt + kr * (-v * s.C1);
      H[tri(3, 3)] += k * gv * gv;
      H[tri(4, 3)] += k * gw * gv + kr * (-s.S1);
      H[tri(4, 4)] += k * gw * gw + kr * (-v * s.C2);
    }
    // y-residual
    {
      const double r = ey + v * s.S0;
      const double gt = v * s.C0, gv = s.S0, gw = v * s.C1;
      const double k = 2.0 * wq * P.Q[1];
      const double kr = k * r;
      *q += wq * P.Q[1] * r * r;
      g[1] += kr; g[2] += kr * gt; g[3] += kr * gv; g[4] += kr * gw;
      H[tri(1, 1)] += k;
      H[tri(2, 1)] += k * gt; H[tri(3, 1)] += k * gv; H[tri(4, 1)] += k * gw;
      H[tri(2, 2)] += k * gt * gt + kr * (-v * s.S0);
      H[tri(3, 2)] += k * gv * gt + kr * (s.C0);
      H[tri(4, 2)] += k * gw * gt + kr * (-v * s.S1);
      H[tri(3, 3)] += k * gv * gv;
      H[tri(4, 3)] += k * gw * gv + kr * (s.C1);
      H[tri(4, 4)] += k * gw * gw + kr * (-v * s.S2);
    }
    // heading residual (linear in th, w)
    {
      const double k = 2.0 * wq * P.Q[2];
      *q += wq * P.Q[2] * et * et;
      g[2] += k * et; g[4] += k * et * tau;
      H[tri(2, 2)] += k; H[tri(4, 2)] += k * tau; H[tri(4, 4)] += k * tau * tau;
    }
  }

  // ---- value + derivatives --------------------------------------------------------------
  // A [NX*NX] row-major, B [NX*NU] row-major, g [NZ] = dq/dz (unscaled), W [15] packed lower
  // triangle of  df * d2q/dz2 + sum_i lam_i d2phi_i/dz2.
  template <class PG, class PS>
  MPCV_HD static void der(const Params& P, const double* x, const double* u, PG pg, PS ps,
                          const double* lam, double df, bool want_hess, double* xn, double* A, double* B,
                          double* q, double* g, double* W) {
    const double v = u[0], w = u[1], th = x[2];
    Sums e = {0, 0, 0, 0, 0, 0};   // end-of-interval sums (dynamics)
    double qa = 0.0;
    double Hq[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) Hq[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) g[i] = 0.0;
    if (KIND == 1) {
      double s, c;
      sincos_(th, &s, &c);
      e.C0 = P.T * c; e.S0 = P.T * s;
    } else {
      const int M = P.M;
      const double h = P.T / M, hh = 0.5 * h, h6 = h / 6.0;
      double s0, c0, sd, cd;
      sincos_(th, &s0, &c0);
      sincos_(hh * w, &sd, &cd);
      double xr = 0, yr = 0, tr = 0;
      if (KIND == 0) { xr = pg[0]; yr = pg[1]; tr = pg[2]; }
      const double ex = x[0] - xr, ey = x[1] - yr;
      for (int j = 0; j < M; ++j) {
        const double t0 = (2 * j) * hh, t1 = (2 * j + 1) * hh, t2 = (2 * j + 2) * hh;
        double s1, c1, s2, c2;
        rot_(s0, c0, sd, cd, &s1, &c1);
        rot_(s1, c1, sd, cd, &s2, &c2);
        if (KIND == 0) {
          Sums sp;
          // k1 point: base
          quad_point(P, h6, ex, ey, th + t0 * w - tr, t0, v, e, &qa, g, Hq);
          // k2 point: base + hh*(stage-0 trig)
          sp.C0 = e.C0 + hh * c0; sp.S0 = e.S0 + hh * s0;
          sp.C1 = e.C1 + hh * t0 * c0; sp.S1 = e.S1 + hh * t0 * s0;
          sp.C2 = e.C2 + hh * t0 * t0 * c0; sp.S2 = e.S2 + hh * t0 * t0 * s0;
          quad_point(P, 2.0 * h6, ex, ey, th + t1 * w - tr, t1, v, sp, &qa, g, Hq);
          // k3 point: base + hh*(stage-1 trig)
          sp.C0 = e.C0 + hh * c1; sp.S0 = e.S0 + hh * s1;
          sp.C1 = e.C1 + hh * t1 * c1; sp.S1 = e.S1 + hh * t1 * s1;
          sp.C2 = e.C2 + hh * t1 * t1 * c1; sp.S2 = e.S2 + hh * t1 * t1 * s1;
          quad_point(P, 2.0 * h6, ex, ey, th + t1 * w - tr, t1, v, sp, &qa, g, Hq);
          // k4 point: base + h*(stage-1 trig)
          sp.C0 = e.C0 + h * c1; sp.S0 = e.S0 + h * s1;
          sp.C1 = e.C1 + h * t1 * c1; sp.S1 = e.S1 + h * t1 * s1;
          sp.C2 = e.C2 + h * t1 * t1 * c1; sp.S2 = e.S2 + h * t1 * t1 * s1;
          quad_point(P, h6, ex, ey, th + t2 * w - tr, t2, v, sp, &qa, g, Hq);
        }
        // advance the Simpson sums over this sub-step
        e.C0 += h6 * (c0 + 4.0 * c1 + c2);
        e.S0 += h6 * (s0 + 4.0 * s1 + s2);
        e.C1 += h6 * (t0 * c0 + 4.0 * t1 * c1 + t2 * c2);
        e.S1 += h6 * (t0 * s0 + 4.0 * t1 * s1 + t2 * s2);
        e.C2 += h6 * (t0 * t0 * c0 + 4.0 * t1 * t1 * c1 + t2 * t2 * c2);
        e.S2 += h6 * (t0 * t0 * s0 + 4.0 * t1 * t1 * s1 + t2 * t2 * s2);
        s0 = s2; c0 = c2;
      }
    }
    // control cost / node costs
    if (KIND == 0) {
      qa += P.T * (P.R[0] * v * v + P.R[1] * w * w);
      g[3] += 2.0 * P.T * P.R[0] * v; g[4] += 2.0 * P.T * P.R[1] * w;
      Hq[tri(3, 3)] += 2.0 * P.T * P.R[0]; Hq[tri(4, 4)] += 2.0 * P.T * P.R[1];
    } else {
      double r0, r1, r2, d0, d1;
      if (KIND == 1) { r0 = x[0] - pg[0]; r1 = x[1] - pg[1]; r2 = th - pg[2]; d0 = v; d1 = w; }
      else { r0 = x[0] - ps[0]; r1 = x[1] - ps[1]; r2 = th - ps[2]; d0 = v - ps[3]; d1 = w - ps[4]; }
      qa = P.Q[0] * r0 * r0 + P.Q[1] * r1 * r1 + P.Q[2] * r2 * r2 + P.R[0] * d0 * d0 + P.R[1] * d1 * d1;
      g[0] = 2.0 * P.Q[0] * r0; g[1] = 2.0 * P.Q[1] * r1; g[2] = 2.0 * P.Q[2] * r2;
      g[3] = 2.0 * P.R[0] * d0; g[4] = 2.0 * P.R[1] * d1;
      Hq[tri(0, 0)] = 2.0 * P.Q[0]; Hq[tri(1, 1)] = 2.0 * P.Q[1]; Hq[tri(2, 2)] = 2.0 * P.Q[2];
      Hq[tri(3, 3)] = 2.0 * P.R[0]; Hq[tri(4, 4)] = 2.0 * P.R[1];
    }
    *q = qa;
    xn[0] = x[0] + v * e.C0; xn[1] = x[1] + v * e.S0; xn[2] = th + P.T * w;
    A[0] = 1; A[1] = 0; A[2] = -v * e.S0;
    A[3] = 0; A[4] = 1; A[5] = v * e.C0;
    A[6] = 0; A[7] = 0; A[8] = 1;
    B[0] = e.C0; B[1] = -v * e.S1;
    B[2] = e.S0; B[3] = v * e.C1;
    B[4] = 0;    B[5] = P.T;
    if (want_hess) {
#pragma unroll
      for (int i = 0; i < 15; ++i) W[i] = df * Hq[i];
      const double lx = lam[0], ly = lam[1];
      W[tri(2, 2)] += v * (-lx * e.C0 - ly * e.S0);
      W[tri(4, 2)] += v * (-lx * e.C1 - ly * e.S1);
      W[tri(4, 4)] += v * (-lx * e.C2 - ly * e.S2);
      W[tri(3, 2)] += -lx * e.S0 + ly * e.C0;
      W[tri(4, 3)] += -lx * e.S1 + ly * e.C1;
    }
  }
};

// ---------------------------------------------------------------------------------------
// Linear models x+ = A x + B u, per-problem (A,B) in pg = [A row-major (NXP^2), B (NXP)]
// (the caller discretises: mpc.util.c2d of Inverted_pendulum/...:24,
//  Trajectory_tracking_lateral_error.py:40, Trajectory_tracking_dynamic_model.py:134, or
//  RK4-of-linear).  Node cost sum_i Q_i (x_i - r_i)^2 + R (u - r_u)^2 [+ R1 (u - u_prev)^2],
//  per-stage references ps = [r (NXP), r_u].
//  DU: state augmented with u_prev (MPCTools' Du[t] = u[t]-u[t-1]).
// ---------------------------------------------------------------------------------------
template <int NXP, bool DU>
struct Linear {
  static constexpr int NX = NXP + (DU ? 1 : 0), NU = 1, NZ = NX + 1;
  static constexpr int NPG = NXP * NXP + NXP, NPS = NXP + 1;
  static constexpr bool HAS_UPREV = DU;
  // (A, B) come from the problem's parameters and the cost Hessian from the weights: the same for every stage
  static constexpr bool LTI = true;
  static constexpr int DER_MINB = 4;
  // five states (pendulum with u_prev): at 128 registers the factor kernel spills 1.1 KB per thread; 168 registers
  // (3 CTAs per SM) cost occupancy this HBM-bound configuration does not need — C3 129 - 136 -> 120 ms per step.  Four
  // states (C5): no difference (44.3 / 44.1 ms).
  static constexpr int FAC_MINB = NX >= 5 ? 3 : 4;
  static constexpr int MODEL_ID = NXP == 3 ? (DU ? MPCV_MODEL_LINEAR3_DU : MPCV_MODEL_LINEAR3)
                                           : (DU ? MPCV_MODEL_LINEAR4_DU : MPCV_MODEL_LINEAR4);

  template <class PG, class PS>
  MPCV_HD static void val(const Params& P, const double* x, const double* u, PG pg, PS ps,
                          double* xn, double* q) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < NXP; ++i) {
      double acc = pg[NXP * NXP + i] * u[0];
#pragma unroll
      for (int j = 0; j < NXP; ++j) acc += pg[i * NXP + j] * x[j];
      xn[i] = acc;
      const double e = x[i] - ps[i];
      c += P.Q[i] * e * e;
    }
    const double du = u[0] - ps[NXP];
    c += P.R[0] * du * du;
    if (DU) {
      xn[NXP] = u[0];
      const double d = u[0] - x[NXP];
      c += P.R1 * d * d;
    }
    *q = c;
  }

  template <class PG, class PS>
  MPCV_HD static void der(const Params& P, const double* x, const double* u, PG pg, PS ps,
                          const double* lam, double df, bool want_hess, double* xn, double* A, double* B,
                          double* q, double* g, double* W) {
    (void)lam;
    val(P, x, u, pg, ps, xn, q);
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) A[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) B[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NXP; ++i) {
#pragma unroll
      for (int j = 0; j < NXP; ++j) A[i * NX + j] = pg[i * NXP + j];
      B[i] = pg[NXP * NXP + i];
      g[i] = 2.0 * P.Q[i] * (x[i] - ps[i]);
    }
    g[NX] = 2.0 * P.R[0] * (u[0] - ps[NXP]);
    if (DU) {
      B[NXP] = 1.0;
      const double d = u[0] - x[NXP];
      g[NXP] = -2.0 * P.R1 * d;
      g[NX] += 2.0 * P.R1 * d;
    }
    if (want_hess) {
#pragma unroll
      for (int i = 0; i < NZ * (NZ + 1) / 2; ++i) W[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NXP; ++i) W[tri(i, i)] = df * 2.0 * P.Q[i];
      W[tri(NX, NX)] = df * 2.0 * P.R[0];
      if (DU) {
        W[tri(NXP, NXP)] = df * 2.0 * P.R1;
        W[tri(NX, NX)] += df * 2.0 * P.R1;
        W[tri(NX, NXP)] = -df * 2.0 * P.R1;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
// Frenet kinematic bicycle — Trajectory Tracking/test2.py
//   ode  (:103-112)  ydot = v sin(phi-phit),
//                    phidot = v (tan(delta/L) - kappat/(1-(y-yt) kappat) cos(phi-phit)),  vdot = a
//   cost (:42-51)    (l1 (v-vdes)^2 + l2 (y-yt)^2 + l3 (phi-phit)^2 + l4 a^2 + l5 (tan delta - L kappat)^2)/(Nt+1)
//   RK4, M sub-steps (:118); bounds |delta| <= 0.384, |a| <= 2, |d delta| <= 0.1225 (:31-36,55-59)
// State (y, phi, v, delta_prev), control (d_delta, a) with delta = delta_prev + d_delta: MPCTools' Du bound
// becomes a box on the control and the delta bound a box on the next state.  Stage parameters exactly as the
// code unpacks them: [yt, phit, kappat] = p[:3], vdes = p[3] (the script's builder fills p[2]/p[3] swapped,
// :89-99 — follow the code).  Weights: Q = (l2, l3, l1, l5), R[1] = l4; extra = (L, Nt+1).
//
// Derivatives: the RK4 map depends on z5 = (y, phi, v, delta, a) only.  A second-order forward-mode jet
// (value, gradient, packed Hessian over z5) is pushed through the four stages; the node cost is
// differentiated by hand; the chain rule delta = delta_prev + d_delta duplicates the delta row / column.
// ---------------------------------------------------------------------------------------
struct Jet5 {
  double v, g[5], h[15];
};
MPCV_HD Jet5 jet_const(double c) {
  Jet5 r; r.v = c;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 15; ++i) r.h[i] = 0.0;
  return r;
}
MPCV_HD Jet5 jet_var(double c, int k) { Jet5 r = jet_const(c); r.g[k] = 1.0; return r; }
MPCV_HD Jet5 jet_add(const Jet5& a, const Jet5& b) {
  Jet5 r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = a.g[i] + b.g[i];
#pragma unroll
  for (int i = 0; i < 15; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
MPCV_HD Jet5 jet_axpy(double s, const Jet5& a, const Jet5& b) {   // s*a + b
  Jet5 r; r.v = s * a.v + b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = s * a.g[i] + b.g[i];
#pragma unroll
  for (int i = 0; i < 15; ++i) r.h[i] = s * a.h[i] + b.h[i];
  return r;
}
MPCV_HD Jet5 jet_scale(double s, const Jet5& a) {
  Jet5 r; r.v = s * a.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = s * a.g[i];
#pragma unroll
  for (int i = 0; i < 15; ++i) r.h[i] = s * a.h[i];
  return r;
}
MPCV_HD Jet5 jet_mul(const Jet5& a, const Jet5& b) {
  Jet5 r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j)
      r.h[tri(i, j)] = a.v * b.h[tri(i, j)] + b.v * a.h[tri(i, j)] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  }
  return r;
}
// r = f(a) given f, f', f'' at a.v
MPCV_HD Jet5 jet_fun(const Jet5& a, double f, double f1, double f2) {
  Jet5 r; r.v = f;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.g[i] = f1 * a.g[i];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) r.h[tri(i, j)] = f1 * a.h[tri(i, j)] + f2 * a.g[i] * a.g[j];
  }
  return r;
}

struct FrenetBicycle {
  static constexpr int NX = 4, NU = 2, NZ = 6;
  static constexpr int NPG = 0, NPS = 4;
  static constexpr bool HAS_UPREV = true;
  static constexpr bool LTI = false;
  static constexpr int DER_MINB = 1;
  static constexpr int FAC_MINB = 4;
  static constexpr int MODEL_ID = MPCV_MODEL_FRENET_BICYCLE;

  template <class PS>
  MPCV_HD static void rhs(double L, const double* X, double tdl, double a, PS ps, double* dx) {
    double s, c;
    sincos_(X[1] - ps[1], &s, &c);
    dx[0] = X[2] * s;
    dx[1] = X[2] * (tdl - (ps[2] / (1.0 - (X[0] - ps[0]) * ps[2])) * c);
    dx[2] = a;
  }

  template <class PS>
  MPCV_HD static double node_cost(const Params& P, const double* x, double delta, double a, PS ps) {
    const double ev = x[2] - ps[3], ey = x[0] - ps[0], ep = x[1] - ps[1];
    const double z = tan(delta) - P.extra[0] * ps[2];
    return (ev * ev * P.Q[2] + ey * ey * P.Q[0] + ep * ep * P.Q[1] + a * a * P.R[1] + z * z * P.Q[3]) / P.extra[1];
  }

  template <class PG, class PS>
  MPCV_HD static void val(const Params& P, const double* x, const double* u, PG, PS ps, double* xn, double* q) {
    const int M = P.M;
    const double DT = P.T / M, L = P.extra[0];
    const double delta = x[3] + u[0], a = u[1], tdl = tan(delta / L);
    double X[3] = {x[0], x[1], x[2]};
    for (int j = 0; j < M; ++j) {
      double k1[3], k2[3], k3[3], k4[3], t[3];
      rhs(L, X, tdl, a, ps, k1);
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k1[i];
      rhs(L, t, tdl, a, ps, k2);
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k2[i];
      rhs(L, t, tdl, a, ps, k3);
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT * k3[i];
      rhs(L, t, tdl, a, ps, k4);
#pragma unroll
      for (int i = 0; i < 3; ++i) X[i] = X[i] + DT / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
    xn[0] = X[0]; xn[1] = X[1]; xn[2] = X[2]; xn[3] = delta;
    *q = node_cost(P, x, delta, a, ps);
  }

  // jet of the right-hand side: X = (y, phi, v) jets, tdl = jet of tan(delta/L), aj = jet of a
  template <class PS>
  MPCV_HD static void rhs_jet(const Jet5* X, const Jet5& tdl, const Jet5& aj, PS ps, Jet5* dx) {
    double s, c;
    sincos_(X[1].v - ps[1], &s, &c);
    const Jet5 sj = jet_fun(X[1], s, c, -s), cj = jet_fun(X[1], c, -s, -c);
    dx[0] = jet_mul(X[2], sj);
    // w = kappat / (1 - (y - yt) kappat) as a function of y:  w' = kappat^2 / den^2,  w'' = 2 kappat^3 / den^3
    const double kap = ps[2], den = 1.0 - (X[0].v - ps[0]) * kap, w = kap / den;
    const Jet5 wj = jet_fun(X[0], w, w * w, 2.0 * w * w * w);
    const Jet5 inner = jet_axpy(-1.0, jet_mul(wj, cj), tdl);
    dx[1] = jet_mul(X[2], inner);
    dx[2] = aj;
  }

  template <class PG, class PS>
  MPCV_HD static void der(const Params& P, const double* x, const double* u, PG, PS ps, const double* lam,
                          double df, bool want_hess, double* xn, double* A, double* B, double* q, double* g,
                          double* W) {
    const int M = P.M;
    const double DT = P.T / M, L = P.extra[0];
    const double delta = x[3] + u[0], a = u[1];
    // independent quantities z5 = (y, phi, v, delta, a)
    Jet5 X[3] = {jet_var(x[0], 0), jet_var(x[1], 1), jet_var(x[2], 2)};
    const Jet5 dj = jet_var(delta, 3), aj = jet_var(a, 4);
    const double t = tan(delta / L), sec2 = 1.0 + t * t;
    const Jet5 tdl = jet_fun(dj, t, sec2 / L, 2.0 * t * sec2 / (L * L));
    for (int j = 0; j < M; ++j) {
      Jet5 k1[3], k2[3], k3[3], k4[3], tmp[3];
      rhs_jet(X, tdl, aj, ps, k1);
#pragma unroll
      for (int i = 0; i < 3; ++i) tmp[i] = jet_axpy(DT / 2, k1[i], X[i]);
      rhs_jet(tmp, tdl, aj, ps, k2);
#pragma unroll
      for (int i = 0; i < 3; ++i) tmp[i] = jet_axpy(DT / 2, k2[i], X[i]);
      rhs_jet(tmp, tdl, aj, ps, k3);
#pragma unroll
      for (int i = 0; i < 3; ++i) tmp[i] = jet_axpy(DT, k3[i], X[i]);
      rhs_jet(tmp, tdl, aj, ps, k4);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const Jet5 s1 = jet_add(k1[i], k4[i]), s2 = jet_add(k2[i], k3[i]);
        X[i] = jet_axpy(DT / 6, jet_axpy(2.0, s2, s1), X[i]);
      }
    }
    xn[0] = X[0].v; xn[1] = X[1].v; xn[2] = X[2].v; xn[3] = delta;
    // z6 = (y, phi, v, delta_prev, d_delta, a) -> z5 index
    const int m[6] = {0, 1, 2, 3, 3, 4};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) A[i * 4 + j] = X[i].g[m[j]];
      B[i * 2 + 0] = X[i].g[3];
      B[i * 2 + 1] = X[i].g[4];
    }
    A[12] = 0; A[13] = 0; A[14] = 0; A[15] = 1;
    B[6] = 1; B[7] = 0;
    // node cost, by hand
    const double div = P.extra[1];
    const double ev = x[2] - ps[3], ey = x[0] - ps[0], ep = x[1] - ps[1];
    const double td = tan(delta), sd2 = 1.0 + td * td, z = td - L * ps[2];
    *q = (ev * ev * P.Q[2] + ey * ey * P.Q[0] + ep * ep * P.Q[1] + a * a * P.R[1] + z * z * P.Q[3]) / div;
    const double gd = 2.0 * P.Q[3] * z * sd2 / div;
    g[0] = 2.0 * P.Q[0] * ey / div; g[1] = 2.0 * P.Q[1] * ep / div; g[2] = 2.0 * P.Q[2] * ev / div;
    g[3] = gd; g[4] = gd; g[5] = 2.0 * P.R[1] * a / div;
    if (want_hess) {
      const double hdd = 2.0 * P.Q[3] * (sd2 * sd2 + z * 2.0 * sd2 * td) / div;
      const double hq5[5] = {2.0 * P.Q[0] / div, 2.0 * P.Q[1] / div, 2.0 * P.Q[2] / div, hdd, 2.0 * P.R[1] / div};
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          const int a5 = m[i] >= m[j] ? m[i] : m[j], b5 = m[i] >= m[j] ? m[j] : m[i];
          double v = lam[0] * X[0].h[tri(a5, b5)] + lam[1] * X[1].h[tri(a5, b5)] + lam[2] * X[2].h[tri(a5, b5)];
          if (a5 == b5) v += df * hq5[a5];
          W[tri(i, j)] = v;
        }
      }
    }
  }
};

}  // namespace mpcv
