// mpc_oracle.cpp — CPU oracle for the batched MPC solve.  TEST INFRASTRUCTURE ONLY.
//
// This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may load it.  The shipped path
// (mpc_verde_b200/csrc) is hand-written CUDA and fails loudly without a GPU.
//
// What it restates.  The reference's hot path is `casadi.nlpsol('ipopt')` over a shooting
// rollout.  CasADi (AD), IPOPT (interior point) and MUMPS (LDL^T) are third-party binaries
// that are NOT under /root/reference and are not installable here (no network):
//   CasADi — unpinned by the reference; py3.8 / matplotlib 3.4.3 era => CasADi 3.5.5,
//   which bundles IPOPT 3.12.3 + MUMPS 4.10.  MPCTools — unpinned.
// The restatement therefore follows IPOPT's published algorithm (Waechter & Biegler,
// "On the implementation of an interior-point filter line-search algorithm for large-scale
// nonlinear programming", Math. Prog. 106, 2006) with the IPOPT 3.12 default options, and
// the reference's own call sites for everything problem-specific:
//   models / cost / RK4 / transcription  -> cited at each function below
//   parity pin: the reference's committed data dumps (tests/golden/*.csv, produced by
//   tests/golden/make_golden.py from Casadi/{1,2,3}exemplo.xlsx,
//   Inverted_pendulum/invertpend_data_py.xlsx, Trajectory Tracking/dados*.csv).
//
// It is deliberately a DIFFERENT algorithmic route from the CUDA path so that the two are
// not one implementation checked against itself:
//   derivatives : generic second-order forward-mode AD (Jet<NV>), no hand derivation
//   KKT solve   : null-space condensing + dense Cholesky of the reduced Hessian
//                 (the CUDA path uses hand-derived sweeps + a Riccati recursion)
// "Inertia is correct" in IPOPT <=> reduced Hessian positive definite (the constraint
// Jacobian of a shooting transcription always has full row rank), i.e. Cholesky succeeds.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC (oracle/Makefile).

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/mpcv.h"

namespace {

constexpr double kInf = 1e19;  // IPOPT nlp_lower_bound_inf / nlp_upper_bound_inf

// ---------------------------------------------------------------------------------------
// Second-order forward-mode AD scalar: value, gradient and packed lower-triangular Hessian
// with respect to NV independent variables.  This is what CasADi's SX `hess_l` delivers
// symbolically in the reference (nlpsol generates nlp_grad_f / nlp_jac_g / nlp_hess_l).
// ---------------------------------------------------------------------------------------
template <int NV>
struct Jet {
  static constexpr int NH = NV * (NV + 1) / 2;
  double v;
  double g[NV];
  double h[NH];
  Jet() : v(0) { std::fill(g, g + NV, 0.0); std::fill(h, h + NH, 0.0); }
  Jet(double c) : v(c) { std::fill(g, g + NV, 0.0); std::fill(h, h + NH, 0.0); }
  static Jet var(double val, int idx) { Jet r(val); r.g[idx] = 1.0; return r; }
  double hess(int i, int j) const { return i >= j ? h[i * (i + 1) / 2 + j] : h[j * (j + 1) / 2 + i]; }
};

template <int NV>
Jet<NV> unary(const Jet<NV>& a, double f, double f1, double f2) {
  Jet<NV> r;
  r.v = f;
  for (int i = 0; i < NV; ++i) r.g[i] = f1 * a.g[i];
  int k = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++k) r.h[k] = f1 * a.h[k] + f2 * a.g[i] * a.g[j];
  return r;
}
template <int NV> Jet<NV> operator+(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v + b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int NV> Jet<NV> operator-(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v - b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
  return r;
}
template <int NV> Jet<NV> operator-(const Jet<NV>& a) {
  Jet<NV> r; r.v = -a.v;
  for (int i = 0; i < NV; ++i) r.g[i] = -a.g[i];
  for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = -a.h[i];
  return r;
}
template <int NV> Jet<NV> operator*(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v * b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  int k = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++k)
      r.h[k] = a.v * b.h[k] + b.v * a.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
template <int NV> Jet<NV> operator+(const Jet<NV>& a, double b) { Jet<NV> r = a; r.v += b; return r; }
template <int NV> Jet<NV> operator+(double b, const Jet<NV>& a) { return a + b; }
template <int NV> Jet<NV> operator-(const Jet<NV>& a, double b) { Jet<NV> r = a; r.v -= b; return r; }
template <int NV> Jet<NV> operator-(double b, const Jet<NV>& a) { return (-a) + b; }
template <int NV> Jet<NV> operator*(const Jet<NV>& a, double b) {
  Jet<NV> r; r.v = a.v * b;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] * b;
  for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] * b;
  return r;
}
template <int NV> Jet<NV> operator*(double b, const Jet<NV>& a) { return a * b; }
template <int NV> Jet<NV> inv(const Jet<NV>& a) {
  double f = 1.0 / a.v;
  return unary(a, f, -f * f, 2.0 * f * f * f);
}
template <int NV> Jet<NV> operator/(const Jet<NV>& a, const Jet<NV>& b) { return a * inv(b); }
template <int NV> Jet<NV> operator/(const Jet<NV>& a, double b) { return a * (1.0 / b); }
template <int NV> Jet<NV> operator/(double a, const Jet<NV>& b) { return inv(b) * a; }
template <int NV> Jet<NV> sin(const Jet<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return unary(a, s, c, -s); }
template <int NV> Jet<NV> cos(const Jet<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return unary(a, c, -s, -c); }
template <int NV> Jet<NV> tan(const Jet<NV>& a) {
  double t = std::tan(a.v), d = 1.0 + t * t;
  return unary(a, t, d, 2.0 * t * d);
}
using std::cos;
using std::sin;
using std::tan;

inline double valueof(double a) { return a; }
template <int NV> double valueof(const Jet<NV>& a) { return a.v; }

// ---------------------------------------------------------------------------------------
// Models.  step<T>(): one shooting interval  x_{k+1} = phi(x_k,u_k;p),  q = interval cost.
//   pg = problem-global parameters, ps = parameters of this stage.
// ---------------------------------------------------------------------------------------

// Unicycle rhs = [v cos(theta), v sin(theta), omega]
//   Casadi/single_shooting_v1.py:70-81, Casadi/multiple_shooting_casadi.py:68-73,
//   mpctools/multiple_shooting_mpctools.py:37-42
template <class T>
inline void unicycle_rhs(const T* x, const T* u, T* dx) {
  dx[0] = u[0] * cos(x[2]);
  dx[1] = u[0] * sin(x[2]);
  dx[2] = u[1];
}

// Stage cost L = (x-ref)'Q(x-ref) + u'Ru — multiple_shooting_casadi.py:78-87
template <class T>
inline T unicycle_L(const mpcv_spec& s, const T* x, const T* u, const double* ref) {
  T e0 = x[0] - ref[0], e1 = x[1] - ref[1], e2 = x[2] - ref[2];
  return e0 * e0 * s.Q[0] + e1 * e1 * s.Q[1] + e2 * e2 * s.Q[2] + u[0] * u[0] * s.R[0] + u[1] * u[1] * s.R[1];
}

// RK4 with M sub-steps and cost quadrature: multiple_shooting_casadi.py:98-114
// (identical in single_shooting_v2.py:97-113).  DT = T/M; k1..k4 of both rhs and L.
struct UnicycleRk4Quad {
  static constexpr int NX = 3, NU = 2, NPG = 3, NPS = 0;
  static constexpr bool HAS_UPREV = false;
  template <class T>
  static void step(const mpcv_spec& s, const T* x, const T* u, const double* pg, const double*, T* xn, T& q) {
    const int M = s.M;
    const double DT = s.T / M;
    T X[3] = {x[0], x[1], x[2]};
    T Q(0.0);
    for (int j = 0; j < M; ++j) {
      T k1[3], k2[3], k3[3], k4[3], t[3];
      unicycle_rhs(X, u, k1);
      T k1q = unicycle_L(s, X, u, pg);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k1[i];
      unicycle_rhs(t, u, k2);
      T k2q = unicycle_L(s, t, u, pg);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k2[i];
      unicycle_rhs(t, u, k3);
      T k3q = unicycle_L(s, t, u, pg);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT * k3[i];
      unicycle_rhs(t, u, k4);
      T k4q = unicycle_L(s, t, u, pg);
      for (int i = 0; i < 3; ++i) X[i] = X[i] + DT / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
      Q = Q + DT / 6 * (k1q + 2 * k2q + 2 * k3q + k4q);
    }
    for (int i = 0; i < 3; ++i) xn[i] = X[i];
    q = Q;
  }
};

// Forward Euler + node cost: single_shooting_v1.py:85-91 (st_next = st + f*T) and
// :100-105 (obj += (st-ref)'Q(st-ref) + con'R con for k in range(N)).
struct UnicycleEulerNode {
  static constexpr int NX = 3, NU = 2, NPG = 3, NPS = 0;
  static constexpr bool HAS_UPREV = false;
  template <class T>
  static void step(const mpcv_spec& s, const T* x, const T* u, const double* pg, const double*, T* xn, T& q) {
    T f[3];
    unicycle_rhs(x, u, f);
    for (int i = 0; i < 3; ++i) xn[i] = x[i] + f[i] * s.T;
    q = unicycle_L(s, x, u, pg);
  }
};

// MPCTools: getCasadiFunc(ode, rk4=True, Delta, M) (classic RK4 on the first argument,
// M sub-steps of Delta/M; mpctools/multiple_shooting_mpctools.py:48,
// Trajectory_tracking.py:51) and node cost l(x,u,p) = (x-p[:3])'Q(x-p[:3]) + (u-p[3:5])'R(u-p[3:5])
// (Trajectory_tracking.py:57-58).  multiple_shooting_mpctools.py:53-55 is the special case
// p = (goal,0,0), R = I.
struct UnicycleRk4Node {
  static constexpr int NX = 3, NU = 2, NPG = 0, NPS = 5;
  static constexpr bool HAS_UPREV = false;
  template <class T>
  static void step(const mpcv_spec& s, const T* x, const T* u, const double*, const double* ps, T* xn, T& q) {
    const int M = s.M;
    const double DT = s.T / M;
    T X[3] = {x[0], x[1], x[2]};
    for (int j = 0; j < M; ++j) {
      T k1[3], k2[3], k3[3], k4[3], t[3];
      unicycle_rhs(X, u, k1);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k1[i];
      unicycle_rhs(t, u, k2);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k2[i];
      unicycle_rhs(t, u, k3);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT * k3[i];
      unicycle_rhs(t, u, k4);
      for (int i = 0; i < 3; ++i) X[i] = X[i] + DT / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
    for (int i = 0; i < 3; ++i) xn[i] = X[i];
    T e0 = x[0] - ps[0], e1 = x[1] - ps[1], e2 = x[2] - ps[2], d0 = u[0] - ps[3], d1 = u[1] - ps[4];
    q = e0 * e0 * s.Q[0] + e1 * e1 * s.Q[1] + e2 * e2 * s.Q[2] + d0 * d0 * s.R[0] + d1 * d1 * s.R[1];
  }
};

// Linear models x+ = A x + B u with per-problem (A,B) in pg (row-major A then B), produced by
// the caller with mpc.util.c2d (exact ZOH; Inverted_pendulum/...:24,
// Trajectory_tracking_lateral_error.py:40, Trajectory_tracking_dynamic_model.py:134) or by
// RK4 of the continuous linear model.  Node cost sum_i Q_i (x_i-r_i)^2 + R (u-r_u)^2
// (Trajectory_tracking_lateral_error.py:54-55, Trajectory_tracking_dynamic_model.py:51-55).
// DU variant: state augmented with u_prev, extra cost R1 (u-u_prev)^2 — MPCTools' Du variable
// (Du[t] = u[t]-u[t-1], u[-1] = uprev); the pendulum cost
// (Q00 (x1-xt))^2 + (Q22 x3)^2 + (R1 du)^2 of Inverted_pendulum/...:47-55 is expressed
// with squared weights by the caller.
template <int NXP, bool DU>
struct LinearModel {
  static constexpr int NX = NXP + (DU ? 1 : 0), NU = 1, NPG = NXP * NXP + NXP, NPS = NXP + 1;
  static constexpr bool HAS_UPREV = DU;
  template <class T>
  static void step(const mpcv_spec& s, const T* x, const T* u, const double* pg, const double* ps, T* xn, T& q) {
    const double* A = pg;
    const double* B = pg + NXP * NXP;
    for (int i = 0; i < NXP; ++i) {
      T acc = u[0] * B[i];
      for (int j = 0; j < NXP; ++j) acc = acc + x[j] * A[i * NXP + j];
      xn[i] = acc;
    }
    T c(0.0);
    for (int i = 0; i < NXP; ++i) {
      T e = x[i] - ps[i];
      c = c + e * e * s.Q[i];
    }
    T du = u[0] - ps[NXP];
    c = c + du * du * s.R[0];
    if (DU) {
      xn[NXP] = u[0];
      T d = u[0] - x[NXP];
      c = c + d * d * s.R1;
    }
    q = c;
  }
};

// Frenet kinematic bicycle — Trajectory Tracking/test2.py:103-112 (ode), :42-51 (cost),
// :118 (RK4, M=1), :31-36,55-59 (bounds incl. |d delta| <= 0.1225).  State (y, phi, v,
// delta_prev), control (d_delta, a): delta = delta_prev + d_delta, so MPCTools' Du bound is
// a simple box on the control and the delta bound a box on the next state.
// Parameters exactly as the code unpacks them: [yt, phit, kappat] = p[:3], vdes = p[3]
// (the reference's builder fills p[2]/p[3] swapped, test2.py:89-99 — follow the code).
// weights: Q = (lambda2 y, lambda3 phi, lambda1 v, lambda5 z), R = (-, lambda4 a);
// extra[0] = L, extra[1] = Nt+1 divisor.
struct FrenetBicycle {
  static constexpr int NX = 4, NU = 2, NPG = 0, NPS = 4;
  static constexpr bool HAS_UPREV = true;
  template <class T>
  static void rhs(const mpcv_spec& s, const T* x, const T& delta, const T& a, const double* ps, T* dx) {
    const double L = s.extra[0];
    T dphi = x[1] - ps[1];
    dx[0] = x[2] * sin(dphi);
    dx[1] = x[2] * (tan(delta / L) - (ps[2] / (1.0 - (x[0] - ps[0]) * ps[2])) * cos(dphi));
    dx[2] = a;
  }
  template <class T>
  static void step(const mpcv_spec& s, const T* x, const T* u, const double*, const double* ps, T* xn, T& q) {
    const int M = s.M;
    const double DT = s.T / M, L = s.extra[0];
    T delta = x[3] + u[0];
    T X[3] = {x[0], x[1], x[2]};
    for (int j = 0; j < M; ++j) {
      T k1[3], k2[3], k3[3], k4[3], t[3];
      rhs(s, X, delta, u[1], ps, k1);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k1[i];
      rhs(s, t, delta, u[1], ps, k2);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT / 2 * k2[i];
      rhs(s, t, delta, u[1], ps, k3);
      for (int i = 0; i < 3; ++i) t[i] = X[i] + DT * k3[i];
      rhs(s, t, delta, u[1], ps, k4);
      for (int i = 0; i < 3; ++i) X[i] = X[i] + DT / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
    for (int i = 0; i < 3; ++i) xn[i] = X[i];
    xn[3] = delta;
    T ev = x[2] - ps[3], ey = x[0] - ps[0], ep = x[1] - ps[1];
    T z = tan(delta) - L * ps[2];
    q = (ev * ev * s.Q[2] + ey * ey * s.Q[0] + ep * ep * s.Q[1] + u[1] * u[1] * s.R[1] + z * z * s.Q[3]) / s.extra[1];
  }
};

// ---------------------------------------------------------------------------------------
// Dense helpers (row-major)
// ---------------------------------------------------------------------------------------
// Cholesky A = L L^T in place (lower). Returns false if a pivot is not strictly positive.
bool cholesky(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * n + j];
      for (int k = 0; k < j; ++k) v -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = v / d;
    }
  }
  return true;
}
void chol_solve(const std::vector<double>& L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= L[i * n + k] * b[k];
    b[i] = v / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= L[k * n + i] * b[k];
    b[i] = v / L[i * n + i];
  }
}

// ---------------------------------------------------------------------------------------
// The transcribed NLP and its interior-point solve
// ---------------------------------------------------------------------------------------
struct SolveStats {
  int status = 0, iters = 0;
  int n_inertia_corrections = 0, n_backtracks = 0, n_soc = 0;
  int max_filter = 0, n_resto = 0;   // largest filter of the solve; restoration steps taken
  double f = 0, mu = 0, err = 0, df = 1;
};

template <class Model>
class Ocp {
 public:
  static constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  using J = Jet<NZ>;

  const mpcv_spec& s;
  const int N;
  const bool single;
  const double* xbar;   // p[0..NX)
  const double* pg;     // p[NX .. NX+NPG)
  const double* pst;    // p[NX+NPG ..), N * NPS

  // iterate, stage-wise
  std::vector<double> X, U;          // X: (N+1)*NX, U: N*NU
  // derivative blocks at the current iterate
  std::vector<double> A, Bm, W, gx, gu, phi, qv;   // per stage
  std::vector<double> lam;                          // (N+1)*NX multipliers of g rows

  Ocp(const mpcv_spec& spec, const double* p)
      : s(spec), N(spec.N), single(spec.shooting == MPCV_SHOOTING_SINGLE), xbar(p), pg(p + NX),
        pst(p + NX + Model::NPG),
        X((N + 1) * NX), U(N * NU), A(N * NX * NX), Bm(N * NX * NU), W(N * NZ * NZ), gx(N * NX),
        gu(N * NU), phi(N * NX), qv(N), lam((N + 1) * NX, 0.0) {}

  bool blocked(int k) const { return s.ntu > 0 && k >= s.ntu; }

  const double* ps(int k) const { return pst + (size_t)k * Model::NPS; }

  // value-only interval evaluation
  void step_val(int k, const double* x, const double* u, double* xn, double& q) const {
    double ub[NU];
    for (int i = 0; i < NU; ++i) ub[i] = u[i];
    if (blocked(k)) ub[0] = x[NX - 1];  // move blocking: u_k = u_{k-1} (DU models keep u_prev last)
    Model::template step<double>(s, x, ub, pg, ps(k), xn, q);
  }

  // interval value + first and second derivatives; lamn = multiplier of the defect row k+1
  void step_der(int k, const double* x, const double* u, const double* lamn) {
    J xj[NX], uj[NU], xn[NX], q;
    for (int i = 0; i < NX; ++i) xj[i] = J::var(x[i], i);
    for (int i = 0; i < NU; ++i) uj[i] = J::var(u[i], NX + i);
    if (blocked(k)) uj[0] = xj[NX - 1];
    Model::template step<J>(s, xj, uj, pg, ps(k), xn, q);
    double* Ak = &A[k * NX * NX];
    double* Bk = &Bm[k * NX * NU];
    double* Wk = &W[k * NZ * NZ];
    for (int i = 0; i < NX; ++i) {
      phi[k * NX + i] = xn[i].v;
      for (int j = 0; j < NX; ++j) Ak[i * NX + j] = xn[i].g[j];
      for (int j = 0; j < NU; ++j) Bk[i * NU + j] = xn[i].g[NX + j];
    }
    qv[k] = q.v;
    for (int i = 0; i < NX; ++i) gx[k * NX + i] = q.g[i];
    for (int i = 0; i < NU; ++i) gu[k * NU + i] = q.g[NX + i];
    for (int a = 0; a < NZ; ++a)
      for (int b = 0; b < NZ; ++b) {
        double h = q.hess(a, b);
        for (int i = 0; i < NX; ++i) h += lamn[i] * xn[i].hess(a, b);
        Wk[a * NZ + b] = h;
      }
  }
};

struct IpmOptions {
  // IPOPT 3.12 defaults (option names in comments)
  double tol, mu_init, bound_push, bound_frac, bound_relax_factor, scal_max_grad;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol;
  int max_iter, max_soc;
  // acceptable-level termination (IpOptErrorConvCheck::CurrentIsAcceptable)
  double acceptable_tol = 1e-6, acceptable_obj_change_tol = 1e20, acceptable_dual_inf_tol = 1e10,
         acceptable_constr_viol_tol = 1e-2, acceptable_compl_inf_tol = 1e-2;
  int acceptable_iter = 15;
  double kappa_eps = 10.0;        // barrier_tol_factor
  double kappa_mu = 0.2;          // mu_linear_decrease_factor
  double theta_mu = 1.5;          // mu_superlinear_decrease_power
  double tau_min = 0.99;          // tau_min
  double s_max = 100.0;           // s_max
  double kappa_sigma = 1e10;      // kappa_sigma
  double kappa_d = 1e-4;          // kappa_d (damping of one-sided bounds)
  double kappa_resto = 0.9;       // required_infeasibility_reduction
  int resto_max_iter = 40, resto_max_backtrack = 30;
  double constr_mult_init_max = 1e3;
  // filter line search
  double theta_max_fact = 1e4, theta_min_fact = 1e-4, eta_phi = 1e-8, delta = 1.0, s_phi = 2.3,
         s_theta = 1.1, gamma_phi = 1e-8, gamma_theta = 1e-5, alpha_min_frac = 0.05,
         kappa_soc = 0.99, obj_max_inc = 5.0;
  // inertia correction (IpPDPerturbationHandler)
  double delta_w_init = 1e-4, delta_w_min = 1e-20, delta_w_max = 1e20, kappa_w_minus = 1.0 / 3.0,
         kappa_w_plus = 8.0, kappa_w_plus_first = 100.0;
};

IpmOptions options_from_spec(const mpcv_spec& s) {
  IpmOptions o;
  o.tol = s.tol > 0 ? s.tol : 1e-8;
  o.max_iter = s.max_iter > 0 ? s.max_iter : 3000;
  o.max_soc = s.max_soc >= 0 ? s.max_soc : 4;
  o.mu_init = s.mu_init > 0 ? s.mu_init : 0.1;
  o.bound_push = s.bound_push > 0 ? s.bound_push : 1e-2;
  o.bound_frac = s.bound_frac > 0 ? s.bound_frac : 1e-2;
  o.bound_relax_factor = s.bound_relax_factor >= 0 ? s.bound_relax_factor : 1e-8;
  o.scal_max_grad = s.nlp_scaling_max_gradient > 0 ? s.nlp_scaling_max_gradient : 100.0;
  o.dual_inf_tol = s.dual_inf_tol > 0 ? s.dual_inf_tol : 1.0;
  o.constr_viol_tol = s.constr_viol_tol > 0 ? s.constr_viol_tol : 1e-4;
  o.compl_inf_tol = s.compl_inf_tol > 0 ? s.compl_inf_tol : 1e-4;
  o.acceptable_tol = s.acceptable_tol > 0 ? s.acceptable_tol : 1e-6;
  o.acceptable_iter = s.acceptable_iter != 0 ? s.acceptable_iter : 15;
  o.acceptable_obj_change_tol = s.acceptable_obj_change_tol > 0 ? s.acceptable_obj_change_tol : 1e20;
  return o;
}

// IPOPT's Compare_le(lhs, rhs, BasVal): lhs - rhs <= 10*eps*|BasVal|
inline bool compare_le(double lhs, double rhs, double basval) {
  return lhs - rhs <= 10.0 * std::numeric_limits<double>::epsilon() * std::fabs(basval);
}

template <class Model>
class Ipm {
 public:
  using O = Ocp<Model>;
  static constexpr int NX = O::NX, NU = O::NU, NZ = O::NZ;

  O ocp;
  IpmOptions opt;
  const int N;
  const bool single;
  const int n;   // number of decision variables
  const int m;   // number of equality rows
  // decision vector in the reference layout; index maps
  std::vector<int> ix, iu;  // ix[k*NX+i] -> var index (MS only), iu[k*NU+i] -> var index
  std::vector<double> w, lb, ub, zl, zu;
  std::vector<char> hasl, hasu, fixed;
  std::vector<double> c;     // constraint residuals (m)
  double df = 1.0;           // objective scaling factor (nlp_scaling_method gradient-based)
  double mu, tau;
  SolveStats st;

  Ipm(const mpcv_spec& spec, const double* p)
      : ocp(spec, p), opt(options_from_spec(spec)), N(spec.N),
        single(spec.shooting == MPCV_SHOOTING_SINGLE),
        n(single ? NU * spec.N : NX * (spec.N + 1) + NU * spec.N),
        m(single ? 0 : NX * (spec.N + 1)) {
    ix.assign((N + 1) * NX, -1);
    iu.assign(N * NU, -1);
    if (single) {
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) iu[k * NU + i] = k * NU + i;
    } else {
      // interleaved [X0,U0,X1,U1,...,X_N]: multiple_shooting_casadi.py:116-178
      for (int k = 0; k <= N; ++k)
        for (int i = 0; i < NX; ++i) ix[k * NX + i] = k * NZ + i;
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) iu[k * NU + i] = k * NZ + NX + i;
    }
    w.resize(n); lb.resize(n); ub.resize(n); zl.assign(n, 0.0); zu.assign(n, 0.0);
    hasl.assign(n, 0); hasu.assign(n, 0); fixed.assign(n, 0);
    c.assign(std::max(m, 1), 0.0);
  }

  // ---- unpack decision vector into stage arrays; single shooting rolls the states out ----
  void unpack(const std::vector<double>& v, std::vector<double>& X, std::vector<double>& U) const {
    X.resize((N + 1) * NX); U.resize(N * NU);
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i) U[k * NU + i] = v[iu[k * NU + i]];
    if (single) {
      for (int i = 0; i < NX; ++i) X[i] = ocp.xbar[i];
      double q;
      for (int k = 0; k < N; ++k) ocp.step_val(k, &X[k * NX], &U[k * NU], &X[(k + 1) * NX], q);
    } else {
      for (int k = 0; k <= N; ++k)
        for (int i = 0; i < NX; ++i) X[k * NX + i] = v[ix[k * NX + i]];
    }
  }

  // objective (scaled) and constraint residuals at a trial vector
  double eval_f_c(const std::vector<double>& v, std::vector<double>& cres) const {
    std::vector<double> X, U;
    unpack(v, X, U);
    double f = 0;
    double xn[NX], q;
    if (!single)
      for (int i = 0; i < NX; ++i) cres[i] = ocp.xbar[i] - X[i];   // P[:3]-X0, MS:125-130
    for (int k = 0; k < N; ++k) {
      ocp.step_val(k, &X[k * NX], &U[k * NU], xn, q);
      f += q;
      if (!single)
        for (int i = 0; i < NX; ++i) cres[(k + 1) * NX + i] = xn[i] - X[(k + 1) * NX + i];  // MS:172-175
    }
    return df * f;
  }

  double barrier(const std::vector<double>& v, double f_scaled) const {
    double phi = f_scaled;
    for (int i = 0; i < n; ++i) {
      if (hasl[i]) phi -= mu * std::log(v[i] - lb[i]);
      if (hasu[i]) phi -= mu * std::log(ub[i] - v[i]);
      // linear damping of one-sided bounds (kappa_d, Waechter & Biegler 2006, sec. 3.7)
      if (hasl[i] && !hasu[i]) phi += opt.kappa_d * mu * (v[i] - lb[i]);
      if (hasu[i] && !hasl[i]) phi += opt.kappa_d * mu * (ub[i] - v[i]);
    }
    return phi;
  }

  // ---- derivatives at the current iterate w (uses ocp.lam for the Hessian) ----
  // grad: unscaled-by-barrier objective gradient (scaled by df) in variable order.
  std::vector<double> grad;    // df * grad f  (n)
  double f_curr = 0;           // scaled objective
  void eval_derivatives() {
    unpack(w, ocp.X, ocp.U);
    if (single) {
      // adjoint (costate) sweep provides the multipliers that make the condensed Hessian exact:
      // mu_N = 0, mu_k = q_x + A_k' mu_{k+1}.  Needs A_k first -> two passes.
      static const double zero[NX] = {0};
      for (int k = 0; k < N; ++k) ocp.step_der(k, &ocp.X[k * NX], &ocp.U[k * NU], zero);
      std::fill(ocp.lam.begin(), ocp.lam.end(), 0.0);
      for (int k = N - 1; k >= 0; --k)
        for (int i = 0; i < NX; ++i) {
          double v = df * ocp.gx[k * NX + i];
          for (int j = 0; j < NX; ++j) v += ocp.A[k * NX * NX + j * NX + i] * ocp.lam[(k + 1) * NX + j];
          ocp.lam[k * NX + i] = v;
        }
    }
    f_curr = 0;
    for (int k = 0; k < N; ++k) {
      // Hessian of  df*q_k + lam_{k+1}' phi_k : evaluate with lam/df then scale by df
      double ls[NX];
      for (int i = 0; i < NX; ++i) ls[i] = ocp.lam[(k + 1) * NX + i] / df;
      ocp.step_der(k, &ocp.X[k * NX], &ocp.U[k * NU], ls);
      for (int a = 0; a < NZ * NZ; ++a) ocp.W[k * NZ * NZ + a] *= df;
      f_curr += ocp.qv[k];
    }
    f_curr *= df;
    grad.assign(n, 0.0);
    if (single) {
      // reduced gradient dJ/du_k = q_u + B_k' mu_{k+1}
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) {
          double v = df * ocp.gu[k * NU + i];
          for (int j = 0; j < NX; ++j) v += ocp.Bm[k * NX * NU + j * NU + i] * ocp.lam[(k + 1) * NX + j];
          grad[iu[k * NU + i]] = v;
        }
    } else {
      for (int k = 0; k < N; ++k) {
        for (int i = 0; i < NX; ++i) grad[ix[k * NX + i]] = df * ocp.gx[k * NX + i];
        for (int i = 0; i < NU; ++i) grad[iu[k * NU + i]] = df * ocp.gu[k * NU + i];
      }
      for (int i = 0; i < NX; ++i) c[i] = ocp.xbar[i] - ocp.X[i];
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NX; ++i) c[(k + 1) * NX + i] = ocp.phi[k * NX + i] - ocp.X[(k + 1) * NX + i];
    }
  }

  // J' lam in variable order (multiple shooting): row 0: -I on X0; row k+1: [A_k B_k] on stage k, -I on X_{k+1}
  void jt_lam(const std::vector<double>& lamv, std::vector<double>& out) const {
    out.assign(n, 0.0);
    if (single) return;
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) out[ix[k * NX + i]] -= lamv[k * NX + i];
    for (int k = 0; k < N; ++k) {
      const double* Ak = &ocp.A[k * NX * NX];
      const double* Bk = &ocp.Bm[k * NX * NU];
      for (int j = 0; j < NX; ++j) {
        double l = lamv[(k + 1) * NX + j];
        for (int i = 0; i < NX; ++i) out[ix[k * NX + i]] += Ak[j * NX + i] * l;
        for (int i = 0; i < NU; ++i) out[iu[k * NU + i]] += Bk[j * NU + i] * l;
      }
    }
  }

  // ---- KKT solve by null-space condensing -------------------------------------------
  //   [W+Sigma+dw I, J'; J, 0] [d; lam+] = -[r; cres]
  // with r the gradient of the barrier objective.  `useW=false` replaces W by the identity
  // (least-squares multiplier initialisation).  Returns false when the reduced Hessian is
  // not positive definite (== wrong inertia in IPOPT's terms).
  std::vector<double> Hred, Gam;   // factor of the reduced Hessian; sensitivity blocks
  bool factor_ok = false;
  std::vector<double> Wd;          // stage Hessians incl. Sigma + delta (N * NZ*NZ) + terminal NX*NX
  void build_Wd(const std::vector<double>& sigma, double dw, bool useW) {
    Wd.assign(N * NZ * NZ + NX * NX, 0.0);
    for (int k = 0; k < N; ++k) {
      double* Wk = &Wd[k * NZ * NZ];
      if (useW) for (int a = 0; a < NZ * NZ; ++a) Wk[a] = ocp.W[k * NZ * NZ + a];
      else for (int a = 0; a < NZ; ++a) Wk[a * NZ + a] = 1.0;
      if (!single)
        for (int i = 0; i < NX; ++i) Wk[i * NZ + i] += sigma[ix[k * NX + i]] + dw;
      for (int i = 0; i < NU; ++i) Wk[(NX + i) * NZ + NX + i] += sigma[iu[k * NU + i]] + dw;
    }
    double* WN = &Wd[N * NZ * NZ];
    if (!single)
      for (int i = 0; i < NX; ++i) WN[i * NX + i] = (useW ? 0.0 : 1.0) + sigma[ix[N * NX + i]] + dw;
  }

  bool factor(const std::vector<double>& sigma, double dw, bool useW) {
    build_Wd(sigma, dw, useW);
    const int nu = NU * N;
    // Gam[k] : d x_k / d u  (NX x nu), Gam[0] = 0
    Gam.assign((size_t)(N + 1) * NX * nu, 0.0);
    for (int k = 0; k < N; ++k) {
      const double* Ak = &ocp.A[k * NX * NX];
      const double* Bk = &ocp.Bm[k * NX * NU];
      double* Gn = &Gam[(size_t)(k + 1) * NX * nu];
      const double* Gk = &Gam[(size_t)k * NX * nu];
      for (int i = 0; i < NX; ++i) {
        for (int col = 0; col < NU * k; ++col) {
          double v = 0;
          for (int j = 0; j < NX; ++j) v += Ak[i * NX + j] * Gk[j * nu + col];
          Gn[i * nu + col] = v;
        }
        for (int j = 0; j < NU; ++j) Gn[i * nu + k * NU + j] = Bk[i * NU + j];
      }
    }
    Hred.assign((size_t)nu * nu, 0.0);
    std::vector<double> S((size_t)NZ * nu), WS((size_t)NZ * nu);
    for (int k = 0; k <= N; ++k) {
      const int nz = (k < N) ? NZ : NX;
      const double* Wk = &Wd[k * NZ * NZ];
      const int ldw = (k < N) ? NZ : NX;
      const int ncol = (k < N) ? NU * (k + 1) : nu;
      std::fill(S.begin(), S.end(), 0.0);
      for (int i = 0; i < NX; ++i)
        for (int col = 0; col < ncol; ++col) S[i * nu + col] = Gam[(size_t)k * NX * nu + i * nu + col];
      if (k < N)
        for (int i = 0; i < NU; ++i) S[(NX + i) * nu + k * NU + i] = 1.0;
      for (int a = 0; a < nz; ++a)
        for (int col = 0; col < ncol; ++col) {
          double v = 0;
          for (int b = 0; b < nz; ++b) v += Wk[a * ldw + b] * S[b * nu + col];
          WS[a * nu + col] = v;
        }
      for (int r = 0; r < ncol; ++r)
        for (int col = 0; col <= r; ++col) {
          double v = 0;
          for (int a = 0; a < nz; ++a) v += S[a * nu + r] * WS[a * nu + col];
          Hred[(size_t)r * nu + col] += v;
        }
    }
    for (int r = 0; r < nu; ++r)
      for (int col = r + 1; col < nu; ++col) Hred[(size_t)r * nu + col] = Hred[(size_t)col * nu + r];
    // fixed controls (move blocking / lb == ub): unit row/column
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i)
        if (fixed[iu[k * NU + i]]) {
          int r = k * NU + i;
          for (int q = 0; q < nu; ++q) Hred[(size_t)r * nu + q] = Hred[(size_t)q * nu + r] = 0.0;
          Hred[(size_t)r * nu + r] = 1.0;
        }
    factor_ok = cholesky(Hred, nu);
    return factor_ok;
  }

  // back-solve with the current factor.  r: barrier gradient (n), cres: residuals (m).
  void backsolve(const std::vector<double>& r, const std::vector<double>& cres,
                 std::vector<double>& d, std::vector<double>& lamplus) const {
    const int nu = NU * N;
    std::vector<double> dbar((N + 1) * NX, 0.0);
    if (!single) {
      for (int i = 0; i < NX; ++i) dbar[i] = cres[i];
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NX; ++i) {
          double v = cres[(k + 1) * NX + i];
          for (int j = 0; j < NX; ++j) v += ocp.A[k * NX * NX + i * NX + j] * dbar[k * NX + j];
          dbar[(k + 1) * NX + i] = v;
        }
    }
    // stage gradients
    auto rx = [&](int k, int i) { return single ? 0.0 : r[ix[k * NX + i]]; };
    std::vector<double> g(nu, 0.0);
    for (int k = 0; k <= N; ++k) {
      const int nz = (k < N) ? NZ : NX;
      const int ldw = (k < N) ? NZ : NX;
      const double* Wk = &Wd[k * NZ * NZ];
      double t[NZ];
      for (int a = 0; a < nz; ++a) {
        double v = (a < NX) ? rx(k, a) : r[iu[k * NU + a - NX]];
        for (int b = 0; b < NX; ++b) v += Wk[a * ldw + b] * dbar[k * NX + b];
        t[a] = v;
      }
      const int ncol = (k < N) ? NU * (k + 1) : nu;
      for (int col = 0; col < ncol; ++col) {
        double v = 0;
        for (int i = 0; i < NX; ++i) v += Gam[(size_t)k * NX * nu + i * nu + col] * t[i];
        g[col] += v;
      }
      if (k < N)
        for (int i = 0; i < NU; ++i) g[k * NU + i] += t[NX + i];
    }
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i)
        if (fixed[iu[k * NU + i]]) g[k * NU + i] = 0.0;
    for (int i = 0; i < nu; ++i) g[i] = -g[i];
    chol_solve(Hred, nu, g.data());   // g now holds du
    std::vector<double> dx((N + 1) * NX, 0.0);
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) {
        double v = dbar[k * NX + i];
        const int ncol = (k < N) ? NU * k : nu;
        for (int col = 0; col < ncol; ++col) v += Gam[(size_t)k * NX * nu + i * nu + col] * g[col];
        dx[k * NX + i] = v;
      }
    d.assign(n, 0.0);
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i) d[iu[k * NU + i]] = g[k * NU + i];
    lamplus.assign((N + 1) * NX, 0.0);
    if (single) return;
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) d[ix[k * NX + i]] = dx[k * NX + i];
    // lam+_N = W_N dx_N + r_N ; lam+_k = (W_k [dx;du])_x + r_xk + A_k' lam+_{k+1}
    for (int k = N; k >= 0; --k) {
      const int ldw = (k < N) ? NZ : NX;
      const double* Wk = &Wd[k * NZ * NZ];
      for (int i = 0; i < NX; ++i) {
        double v = r[ix[k * NX + i]];
        for (int b = 0; b < NX; ++b) v += Wk[i * ldw + b] * dx[k * NX + b];
        if (k < N) {
          for (int b = 0; b < NU; ++b) v += Wk[i * ldw + NX + b] * g[k * NU + b];
          for (int j = 0; j < NX; ++j) v += ocp.A[k * NX * NX + j * NX + i] * lamplus[(k + 1) * NX + j];
        }
        lamplus[k * NX + i] = v;
      }
    }
  }

  // ---- error measures ---------------------------------------------------------------
  struct Err { double dual, prim, compl_mu; };
  Err errors(double mu_target, double* sd_out = nullptr, double* sc_out = nullptr) const {
    std::vector<double> jl;
    jt_lam(ocp.lam, jl);
    double dual = 0, prim = 0, comp = 0, zsum = 0, lsum = 0;
    int nz = 0;
    for (int i = 0; i < n; ++i) {
      if (fixed[i]) continue;
      double di = grad[i] + (single ? 0.0 : jl[i]) - zl[i] + zu[i];
      dual = std::max(dual, std::fabs(di));
      if (hasl[i]) { comp = std::max(comp, std::fabs((w[i] - lb[i]) * zl[i] - mu_target)); zsum += std::fabs(zl[i]); ++nz; }
      if (hasu[i]) { comp = std::max(comp, std::fabs((ub[i] - w[i]) * zu[i] - mu_target)); zsum += std::fabs(zu[i]); ++nz; }
    }
    if (!single) {
      for (int i = 0; i < m; ++i) { prim = std::max(prim, std::fabs(c[i])); lsum += std::fabs(ocp.lam[i]); }
    }
    double sd = 1.0, sc = 1.0;
    if (m + nz > 0) sd = std::max(opt.s_max, (lsum + zsum) / (m + nz)) / opt.s_max;
    if (nz > 0) sc = std::max(opt.s_max, zsum / nz) / opt.s_max;
    if (sd_out) *sd_out = sd;
    if (sc_out) *sc_out = sc;
    return Err{dual / sd, prim, comp / sc};
  }
  double E(double mu_target) const {
    Err e = errors(mu_target);
    return std::max(e.dual, std::max(e.prim, e.compl_mu));
  }

  // ---- the solve ----------------------------------------------------------------------
  void solve(const double* x0, const double* lbx, const double* ubx) {
    const double eps = std::numeric_limits<double>::epsilon();
    // bounds: relax by bound_relax_factor*max(1,|b|); |b|>=1e19 is infinite; lb==ub is fixed
    for (int i = 0; i < n; ++i) {
      double l = lbx ? lbx[i] : -INFINITY, u = ubx ? ubx[i] : INFINITY;
      w[i] = x0 ? x0[i] : 0.0;
      hasl[i] = l > -kInf; hasu[i] = u < kInf;
      fixed[i] = (hasl[i] && hasu[i] && l == u);
      if (fixed[i]) { w[i] = l; hasl[i] = hasu[i] = 0; }
      lb[i] = hasl[i] ? l - opt.bound_relax_factor * std::max(1.0, std::fabs(l)) : -INFINITY;
      ub[i] = hasu[i] ? u + opt.bound_relax_factor * std::max(1.0, std::fabs(u)) : INFINITY;
    }
    // move blocking: blocked controls are eliminated (MPCTools Du[t] lb=ub=0 => IPOPT fixed variable)
    for (int k = 0; k < N; ++k)
      if (ocp.blocked(k)) { int v = iu[k * NU]; fixed[v] = 1; hasl[v] = hasu[v] = 0; lb[v] = -INFINITY; ub[v] = INFINITY; }
    // push the starting point into the interior: bound_push / bound_frac
    for (int i = 0; i < n; ++i) {
      if (hasl[i] && hasu[i]) {
        double pl = std::min(opt.bound_push * std::max(1.0, std::fabs(lb[i])), opt.bound_frac * (ub[i] - lb[i]));
        double pu = std::min(opt.bound_push * std::max(1.0, std::fabs(ub[i])), opt.bound_frac * (ub[i] - lb[i]));
        w[i] = std::min(std::max(w[i], lb[i] + pl), ub[i] - pu);
      } else if (hasl[i]) {
        w[i] = std::max(w[i], lb[i] + opt.bound_push * std::max(1.0, std::fabs(lb[i])));
      } else if (hasu[i]) {
        w[i] = std::min(w[i], ub[i] - opt.bound_push * std::max(1.0, std::fabs(ub[i])));
      }
    }
    sync_blocked();
    // gradient-based objective scaling at the starting point
    df = 1.0;
    std::fill(ocp.lam.begin(), ocp.lam.end(), 0.0);
    eval_derivatives();
    {
      double gmax = 0;
      for (int i = 0; i < n; ++i) if (!fixed[i]) gmax = std::max(gmax, std::fabs(grad[i]));
      if (gmax > opt.scal_max_grad) df = std::max(opt.scal_max_grad / gmax, 1e-8);
    }
    for (int i = 0; i < n; ++i) { zl[i] = hasl[i] ? 1.0 : 0.0; zu[i] = hasu[i] ? 1.0 : 0.0; }
    mu = opt.mu_init;
    tau = std::max(opt.tau_min, 1.0 - mu);
    eval_derivatives();
    // least-squares multipliers: [I J'; J 0][.; lam] = -[grad f - zl + zu; 0]
    if (!single) {
      std::vector<double> sig0(n, 0.0), r(n), c0(m, 0.0), d, lp;
      for (int i = 0; i < n; ++i) r[i] = grad[i] - zl[i] + zu[i];
      if (factor(sig0, 0.0, false)) {
        backsolve(r, c0, d, lp);
        double lmax = 0;
        for (double v : lp) lmax = std::max(lmax, std::fabs(v));
        if (lmax <= opt.constr_mult_init_max && std::isfinite(lmax)) ocp.lam = lp;
      }
      eval_derivatives();   // Hessian depends on lam
    }

    std::vector<std::pair<double, double>> filter;   // (phi, theta) entries
    // IPOPT's Filter::AddEntry: entries the new one dominates (both coordinates >= the new ones) leave the filter —
    // whatever they reject, the new entry rejects too
    auto filter_add = [&](double phi_e, double th_e) {
      size_t k = 0;
      for (size_t q = 0; q < filter.size(); ++q)
        if (!(filter[q].first >= phi_e && filter[q].second >= th_e)) filter[k++] = filter[q];
      filter.resize(k);
      filter.emplace_back(phi_e, th_e);
      st.max_filter = std::max(st.max_filter, (int)filter.size());
    };
    double theta_max = -1, theta_min = -1;
    double delta_w_last = 0.0;
    std::vector<double> sigma(n), r(n), d, lamplus, dzl(n), dzu(n), wt(n), ct(std::max(m, 1));

    st = SolveStats();
    st.df = df;
    int acc_count = 0;
    double f_last = -1e50;     // IPOPT: last_obj_val_ starts at -1e50
    for (int iter = 0;; ++iter) {
      st.iters = iter;
      // --- convergence test (IpOptErrorConvCheck) ---
      Err e0 = errors(0.0);
      double E0 = std::max(e0.dual, std::max(e0.prim, e0.compl_mu));
      {
        double sd, sc;
        errors(0.0, &sd, &sc);
        double dual_u = e0.dual * sd / df, compl_u = e0.compl_mu * sc / df;
        if (E0 <= opt.tol && dual_u <= opt.dual_inf_tol && e0.prim <= opt.constr_viol_tol &&
            compl_u <= opt.compl_inf_tol) { st.status = MPCV_SOLVE_SUCCEEDED; break; }
        // acceptable level: acceptable_iter consecutive iterates within the acceptable tolerances
        const bool acc = opt.acceptable_iter > 0 && E0 <= opt.acceptable_tol && dual_u <= opt.acceptable_dual_inf_tol &&
                         e0.prim <= opt.acceptable_constr_viol_tol && compl_u <= opt.acceptable_compl_inf_tol &&
                         std::fabs(f_curr - f_last) / std::max(1.0, std::fabs(f_curr)) <= opt.acceptable_obj_change_tol;
        f_last = f_curr;
        if (acc) { if (++acc_count >= opt.acceptable_iter) { st.status = MPCV_SOLVED_TO_ACCEPTABLE_LEVEL; break; } }
        else acc_count = 0;
      }
      if (iter >= opt.max_iter) { st.status = MPCV_MAXIMUM_ITERATIONS_EXCEEDED; break; }
      if (!std::isfinite(E0)) { st.status = MPCV_INVALID_NUMBER_DETECTED; break; }
      // --- barrier update (IpMonotoneMuUpdate) ---
      {
        bool done = false;
        while (!done && E(mu) <= opt.kappa_eps * mu) {
          double new_mu = std::min(opt.kappa_mu * mu, std::pow(mu, opt.theta_mu));
          new_mu = std::max(new_mu, std::min(opt.tol, opt.compl_inf_tol * df) / (opt.kappa_eps + 1.0));
          bool changed = new_mu != mu;
          mu = new_mu;
          tau = std::max(opt.tau_min, 1.0 - mu);
          if (!changed) done = true;
          else filter.clear();
        }
      }
      // --- search direction with inertia correction (IpPDPerturbationHandler) ---
      for (int i = 0; i < n; ++i) {
        double sg = 0, ri = grad[i];
        if (hasl[i]) { double sl = w[i] - lb[i]; sg += zl[i] / sl; ri -= mu / sl; }
        if (hasu[i]) { double su = ub[i] - w[i]; sg += zu[i] / su; ri += mu / su; }
        if (hasl[i] && !hasu[i]) ri += opt.kappa_d * mu;
        if (hasu[i] && !hasl[i]) ri -= opt.kappa_d * mu;
        sigma[i] = sg; r[i] = ri;
      }
      double dw = 0.0;
      bool ok = factor(sigma, 0.0, true);
      while (!ok) {
        if (dw == 0.0) dw = (delta_w_last == 0.0) ? opt.delta_w_init : std::max(opt.delta_w_min, delta_w_last * opt.kappa_w_minus);
        else dw *= (delta_w_last == 0.0 || 1e5 * delta_w_last < dw) ? opt.kappa_w_plus_first : opt.kappa_w_plus;
        if (dw > opt.delta_w_max) break;
        ++st.n_inertia_corrections;
        ok = factor(sigma, dw, true);
      }
      if (!ok) { st.status = MPCV_ERROR_IN_STEP_COMPUTATION; break; }
      if (dw > 0.0) delta_w_last = dw;
      backsolve(r, c, d, lamplus);
      auto bound_steps = [&](const std::vector<double>& dd) {
        for (int i = 0; i < n; ++i) {
          dzl[i] = dzu[i] = 0;
          if (hasl[i]) { double sl = w[i] - lb[i]; dzl[i] = mu / sl - zl[i] - zl[i] / sl * dd[i]; }
          if (hasu[i]) { double su = ub[i] - w[i]; dzu[i] = mu / su - zu[i] + zu[i] / su * dd[i]; }
        }
      };
      bound_steps(d);
      auto ftb_primal = [&](const std::vector<double>& dd) {
        double a = 1.0;
        for (int i = 0; i < n; ++i) {
          if (hasl[i] && dd[i] < 0) a = std::min(a, -tau * (w[i] - lb[i]) / dd[i]);
          if (hasu[i] && dd[i] > 0) a = std::min(a, tau * (ub[i] - w[i]) / dd[i]);
        }
        return a;
      };
      auto ftb_dual = [&]() {
        double a = 1.0;
        for (int i = 0; i < n; ++i) {
          if (hasl[i] && dzl[i] < 0) a = std::min(a, -tau * zl[i] / dzl[i]);
          if (hasu[i] && dzu[i] < 0) a = std::min(a, -tau * zu[i] / dzu[i]);
        }
        return a;
      };
      double alpha_max = ftb_primal(d);
      // --- filter line search (IpBacktrackingLineSearch + IpFilterLSAcceptor) ---
      double theta = 0;
      for (int i = 0; i < m; ++i) theta += std::fabs(c[i]);
      double phi = barrier(w, f_curr);
      double gBD = 0;
      for (int i = 0; i < n; ++i) gBD += r[i] * d[i];
      if (theta_max < 0) { theta_max = opt.theta_max_fact * std::max(1.0, theta); theta_min = opt.theta_min_fact * std::max(1.0, theta); }
      double alpha_min = opt.gamma_theta;
      if (gBD < 0) {
        alpha_min = std::min(opt.gamma_theta, opt.gamma_phi * theta / (-gBD));
        if (theta <= theta_min) alpha_min = std::min(alpha_min, opt.delta * std::pow(theta, opt.s_theta) / std::pow(-gBD, opt.s_phi));
      }
      alpha_min *= opt.alpha_min_frac;
      auto is_ftype = [&](double a) {
        if (theta == 0.0 && gBD > 0.0 && gBD < 100.0 * eps) return true;
        return gBD < 0.0 && a * std::pow(-gBD, opt.s_phi) > opt.delta * std::pow(theta, opt.s_theta);
      };
      auto armijo = [&](double a, double phi_t) { return compare_le(phi_t - phi, opt.eta_phi * a * gBD, phi); };
      auto acceptable = [&](double a, double phi_t, double theta_t) {
        if (!std::isfinite(phi_t) || !std::isfinite(theta_t)) return false;
        if (theta_max > 0 && theta_t > theta_max) return false;
        bool acc;
        if (a > 0 && is_ftype(a) && theta <= theta_min) acc = armijo(a, phi_t);
        else {
          if (phi_t > phi) {
            double basval = 1.0;
            if (std::fabs(phi) > 10.0) basval = std::log10(std::fabs(phi));
            if (std::log10(phi_t - phi) > opt.obj_max_inc + basval) return false;
          }
          acc = compare_le(theta_t, (1.0 - opt.gamma_theta) * theta, theta) ||
                compare_le(phi_t - phi, -opt.gamma_phi * theta, phi);
        }
        if (!acc) return false;
        for (auto& fe : filter)
          if (!(compare_le(phi_t, fe.first, fe.first) || compare_le(theta_t, fe.second, fe.second))) return false;
        return true;
      };
      auto trial = [&](const std::vector<double>& dd, double a, double& phi_t, double& theta_t) {
        for (int i = 0; i < n; ++i) wt[i] = w[i] + a * dd[i];
        if (ocp.s.ntu > 0) sync_blocked_vec(wt);
        double ft = eval_f_c(wt, ct);
        theta_t = 0;
        for (int i = 0; i < m; ++i) theta_t += std::fabs(ct[i]);
        phi_t = barrier(wt, ft);
      };
      double alpha = alpha_max, alpha_test = alpha_max;
      bool accepted = false;
      std::vector<double> d_used = d, lam_used = lamplus;
      int nsteps = 0;
      while (alpha > alpha_min || nsteps == 0) {
        double phi_t, theta_t;
        trial(d, alpha, phi_t, theta_t);
        alpha_test = alpha;
        if (acceptable(alpha, phi_t, theta_t)) { accepted = true; break; }
        // second-order correction on the first rejected trial
        if (nsteps == 0 && opt.max_soc > 0 && !single && theta_t >= theta) {
          std::vector<double> csoc(c.begin(), c.begin() + m), dsoc, lsoc;
          double theta_soc_old = 0, theta_tr = theta_t, alpha_soc = alpha;
          int count = 0;
          bool acc_soc = false;
          while (count < opt.max_soc && !acc_soc && (count == 0 || theta_tr <= opt.kappa_soc * theta_soc_old)) {
            theta_soc_old = theta_tr;
            for (int i = 0; i < m; ++i) csoc[i] = ct[i] + alpha_soc * csoc[i];
            backsolve(r, csoc, dsoc, lsoc);
            alpha_soc = ftb_primal(dsoc);
            double phi_s, theta_s;
            trial(dsoc, alpha_soc, phi_s, theta_s);
            ++st.n_soc;
            if (acceptable(alpha, phi_s, theta_s)) {
              acc_soc = true; d_used = dsoc; lam_used = lsoc; alpha = alpha_soc;
            } else { ++count; theta_tr = theta_s; }
          }
          if (acc_soc) { accepted = true; bound_steps(d_used); break; }
        }
        alpha *= 0.5;
        ++nsteps;
        ++st.n_backtracks;
      }
      if (!accepted) {
        // --- feasibility restoration (IPOPT: IpRestoMinC_1Nrm; here the simplified phase DESIGN.md states): the point
        // we leave enters the filter; damped minimum-norm Gauss-Newton steps on theta = |c|_1 until the point is
        // acceptable to the filter with theta <= kappa_resto theta_R; equality multipliers reset to zero ---
        bool restored = false;
        int n_resto = 0;
        const bool dbg = getenv("MPCO_DEBUG_RESTO") != nullptr;
        if (dbg) fprintf(stderr, "resto enter iter %d theta %.3e phi %.6e alpha_min %.3e gBD %.3e nfil %zu\n", iter, theta, phi, alpha_min, gBD, filter.size());
        if (!single && theta > opt.tol) {
          filter_add(phi - opt.gamma_phi * theta, (1.0 - opt.gamma_theta) * theta);
          const double theta_R = theta;
          double th = theta_R;
          std::vector<double> sigR(n, 0.0), r0(n, 0.0), dr, lr, cc;
          for (int it = 0; it < opt.resto_max_iter && !restored; ++it) {
            // affine scaling: the step is measured in |d|^2 + sum_i (d_i / s_i)^2 over the bound slacks s_i, so that a
            // variable sitting at a bound hardly moves and the fraction-to-the-boundary rule does not choke the step
            for (int i = 0; i < n; ++i) {
              double sg = 0.0;
              if (hasl[i]) { double sl = w[i] - lb[i]; sg += 1.0 / (sl * sl); }
              if (hasu[i]) { double su = ub[i] - w[i]; sg += 1.0 / (su * su); }
              sigR[i] = sg;
            }
            if (!factor(sigR, 0.0, false)) break;
            cc.assign(c.begin(), c.begin() + m);
            backsolve(r0, cc, dr, lr);
            double a = ftb_primal(dr), phi_t = 0, theta_t = 0;
            bool found = false;
            for (int j = 0; j < opt.resto_max_backtrack && !found; ++j) {
              trial(dr, a, phi_t, theta_t);
              if (std::isfinite(phi_t) && theta_t <= (1.0 - 1e-4 * a) * th) found = true;
              else a *= 0.5;
            }
            if (dbg) fprintf(stderr, "  resto it %d found %d a %.3e theta_t %.3e (th %.3e) phi_t %.6e\n", it, (int)found, a, theta_t, th, phi_t);
            if (!found) break;
            for (int i = 0; i < n; ++i) w[i] = wt[i];
            ++n_resto;
            ++st.n_resto;
            th = theta_t;
            eval_derivatives();
            bool fok = std::isfinite(phi_t) && !(theta_max > 0 && theta_t > theta_max);
            for (auto& fe : filter)
              if (!(compare_le(phi_t, fe.first, fe.first) || compare_le(theta_t, fe.second, fe.second))) fok = false;
            if (theta_t <= opt.kappa_resto * theta_R && fok) restored = true;
            if (iter + n_resto >= opt.max_iter) break;
          }
        }
        if (!restored) { st.status = MPCV_RESTORATION_FAILED; st.iters = iter + n_resto; break; }
        std::fill(ocp.lam.begin(), ocp.lam.end(), 0.0);
        for (int i = 0; i < n; ++i) {
          if (hasl[i]) { double sl = w[i] - lb[i]; zl[i] = std::max(std::min(zl[i], opt.kappa_sigma * mu / sl), mu / (opt.kappa_sigma * sl)); }
          if (hasu[i]) { double su = ub[i] - w[i]; zu[i] = std::max(std::min(zu[i], opt.kappa_sigma * mu / su), mu / (opt.kappa_sigma * su)); }
        }
        eval_derivatives();
        iter += n_resto - 1;      // every restoration step counts as an iteration (the loop adds the last one)
        continue;
      }
      // filter augmentation (h-type step)
      {
        double phi_t, theta_t;
        // (wt, ct already hold the accepted trial)
        phi_t = 0; theta_t = 0; (void)phi_t; (void)theta_t;
        double phi_acc = barrier(wt, eval_f_c(wt, ct));
        if (!is_ftype(alpha_test) || !armijo(alpha_test, phi_acc))
          filter_add(phi - opt.gamma_phi * theta, (1.0 - opt.gamma_theta) * theta);
      }
      double alpha_dual = ftb_dual();
      // accept: primal + equality multipliers with alpha (alpha_for_y=primal), bound multipliers with alpha_dual
      for (int i = 0; i < n; ++i) w[i] = wt[i];
      if (!single)
        for (int i = 0; i < m; ++i) ocp.lam[i] += alpha * (lam_used[i] - ocp.lam[i]);
      for (int i = 0; i < n; ++i) {
        if (hasl[i]) {
          zl[i] += alpha_dual * dzl[i];
          double sl = w[i] - lb[i];
          zl[i] = std::max(std::min(zl[i], opt.kappa_sigma * mu / sl), mu / (opt.kappa_sigma * sl));
        }
        if (hasu[i]) {
          zu[i] += alpha_dual * dzu[i];
          double su = ub[i] - w[i];
          zu[i] = std::max(std::min(zu[i], opt.kappa_sigma * mu / su), mu / (opt.kappa_sigma * su));
        }
      }
      eval_derivatives();
    }
    st.mu = mu;
    st.f = f_curr / df;
    st.err = E(0.0);
  }

  // blocked controls mirror the previous control so that outputs read like MPCTools' u
  void sync_blocked() { if (ocp.s.ntu > 0) sync_blocked_vec(w); }
  void sync_blocked_vec(std::vector<double>& v) const {
    for (int k = 1; k < N; ++k)
      if (ocp.blocked(k)) v[iu[k * NU]] = v[iu[(k - 1) * NU]];
  }

  // final point projected into the ORIGINAL bounds (honor_original_bounds=yes)
  void export_solution(const double* lbx, const double* ubx, double* x, double* f, double* g,
                       double* lam_g, double* lam_x) {
    for (int i = 0; i < n; ++i) {
      double v = w[i];
      if (lbx && lbx[i] > -kInf) v = std::max(v, lbx[i]);
      if (ubx && ubx[i] < kInf) v = std::min(v, ubx[i]);
      x[i] = v;
    }
    if (f) *f = st.f;
    if (g) {
      if (single) { for (int i = 0; i < (N + 1) * NX; ++i) g[i] = ocp.X[i]; }
      else for (int i = 0; i < m; ++i) g[i] = c[i];
    }
    if (lam_g) for (int i = 0; i < (N + 1) * NX; ++i) lam_g[i] = ocp.lam[i] / df;
    if (lam_x) for (int i = 0; i < n; ++i) lam_x[i] = (zu[i] - zl[i]) / df;
  }
};

// ---------------------------------------------------------------------------------------
template <class Model>
void dims_of(const mpcv_spec& s, int* nx, int* nu, int* nvar, int* ng, int* np, int* npg, int* nps) {
  const bool single = s.shooting == MPCV_SHOOTING_SINGLE;
  if (nx) *nx = Model::NX;
  if (nu) *nu = Model::NU;
  if (nvar) *nvar = single ? Model::NU * s.N : Model::NX * (s.N + 1) + Model::NU * s.N;
  if (ng) *ng = Model::NX * (s.N + 1);
  if (np) *np = Model::NX + Model::NPG + s.N * Model::NPS;
  if (npg) *npg = Model::NPG;
  if (nps) *nps = Model::NPS;
}

template <class Model>
void solve_range(const mpcv_spec& s, const double* x0, const double* lbx, const double* ubx,
                 const double* p, double* x, double* f, double* g, double* lam_g, double* lam_x,
                 int32_t* status, int32_t* iters, double* stats, int64_t b0, int64_t b1) {
  int nvar, ng, np;
  dims_of<Model>(s, nullptr, nullptr, &nvar, &ng, &np, nullptr, nullptr);
  for (int64_t b = b0; b < b1; ++b) {
    Ipm<Model> ipm(s, p + b * np);
    ipm.solve(x0 ? x0 + b * nvar : nullptr, lbx, ubx);
    ipm.export_solution(lbx, ubx, x + b * nvar, f ? f + b : nullptr, g ? g + b * ng : nullptr,
                        lam_g ? lam_g + b * ng : nullptr, lam_x ? lam_x + b * nvar : nullptr);
    if (status) status[b] = ipm.st.status;
    if (iters) iters[b] = ipm.st.iters;
    if (stats) {
      double* o = stats + b * 8;
      o[0] = ipm.st.n_inertia_corrections; o[1] = ipm.st.n_backtracks; o[2] = ipm.st.n_soc;
      o[3] = ipm.st.mu; o[4] = ipm.st.err; o[5] = ipm.st.df; o[6] = ipm.st.max_filter; o[7] = ipm.st.n_resto;
    }
  }
}

template <class Model>
int solve_batch(const mpcv_spec& s, const double* x0, const double* lbx, const double* ubx,
                const double* p, double* x, double* f, double* g, double* lam_g, double* lam_x,
                int32_t* status, int32_t* iters, double* stats, int64_t B, int nthreads) {
  nthreads = std::max(1, std::min<int>(nthreads, (int)std::max<int64_t>(B, 1)));
  if (nthreads == 1) {
    solve_range<Model>(s, x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, stats, 0, B);
    return 0;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    int64_t b0 = B * t / nthreads, b1 = B * (t + 1) / nthreads;
    th.emplace_back([=, &s] { solve_range<Model>(s, x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, stats, b0, b1); });
  }
  for (auto& t : th) t.join();
  return 0;
}

template <class Model>
int rollout_batch(const mpcv_spec& s, const double* p, const double* U, double* X, double* q, int64_t B) {
  int np;
  dims_of<Model>(s, nullptr, nullptr, nullptr, nullptr, &np, nullptr, nullptr);
  constexpr int NX = Model::NX, NU = Model::NU;
  for (int64_t b = 0; b < B; ++b) {
    Ocp<Model> o(s, p + b * np);
    double* Xb = X + b * NX * (s.N + 1);
    for (int i = 0; i < NX; ++i) Xb[i] = o.xbar[i];
    double acc = 0, qk;
    for (int k = 0; k < s.N; ++k) {
      o.step_val(k, Xb + k * NX, U + b * NU * s.N + k * NU, Xb + (k + 1) * NX, qk);
      acc += qk;
    }
    if (q) q[b] = acc;
  }
  return 0;
}

template <class Model>
int stage_derivs_batch(const mpcv_spec& s, const double* z, const double* pstage, const double* lam,
                       double* xn, double* A, double* Bm, double* q, double* grad, double* H, int64_t B) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, NPP = Model::NPG + Model::NPS;
  std::vector<double> p(NX + Model::NPG + (size_t)s.N * Model::NPS, 0.0);
  for (int64_t b = 0; b < B; ++b) {
    const double* pp = pstage + b * NPP;
    for (int i = 0; i < Model::NPG; ++i) p[NX + i] = pp[i];
    for (int i = 0; i < Model::NPS; ++i) p[NX + Model::NPG + i] = pp[Model::NPG + i];
    mpcv_spec s2 = s;
    s2.ntu = 0;
    Ocp<Model> o(s2, p.data());
    o.step_der(0, z + b * NZ, z + b * NZ + NX, lam + b * NX);
    for (int i = 0; i < NX; ++i) xn[b * NX + i] = o.phi[i];
    for (int i = 0; i < NX * NX; ++i) A[b * NX * NX + i] = o.A[i];
    for (int i = 0; i < NX * NU; ++i) Bm[b * NX * NU + i] = o.Bm[i];
    q[b] = o.qv[0];
    for (int i = 0; i < NX; ++i) grad[b * NZ + i] = o.gx[i];
    for (int i = 0; i < NU; ++i) grad[b * NZ + NX + i] = o.gu[i];
    for (int i = 0; i < NZ * NZ; ++i) H[b * NZ * NZ + i] = o.W[i];
  }
  return 0;
}

// Closed loop of the scripts: solve -> apply u0 -> plant step with the same discretisation
// (multiple_shooting_casadi.py:273 `state_init = F(args['p'],u[:,0])[0]`,
//  single_shooting_v1.py:17-27 shift_timestep) -> shifted warm start.
//   warm_mode MPCV_WARM_SHIFT     : correctly interleaved shift-by-one
//   warm_mode MPCV_WARM_COLD      : X_k = state, U = 0
//   warm_mode MPCV_WARM_REFERENCE : the scripts' own guess vectors, including the layout
//       mismatch of multiple_shooting_casadi.py:284-287 (w0 = [vec(X0); vec(u0)] fed to an
//       interleaved w) and single_shooting_v1.py:173 (reshape(u0.T) = [v0..v9, w0..w9])
template <class Model>
int closed_loop_batch(const mpcv_spec& s, const double* x_init, const double* pglob, const double* ptraj,
                      const double* lbx, const double* ubx, int n_steps, int warm_mode, double stop_radius,
                      double* out_states, double* out_controls, int32_t* out_steps, int32_t* out_iters,
                      int32_t* out_status, int64_t B) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU;
  const int N = s.N;
  const bool single = s.shooting == MPCV_SHOOTING_SINGLE;
  int nvar, np;
  dims_of<Model>(s, nullptr, nullptr, &nvar, nullptr, &np, nullptr, nullptr);
  for (int64_t b = 0; b < B; ++b) {
    std::vector<double> p(np), w0(nvar, 0.0), x(nvar), Xp((N + 1) * NX), Up(N * NU);
    double state[NX];
    for (int i = 0; i < NX; ++i) state[i] = x_init[b * NX + i];
    for (int i = 0; i < Model::NPG; ++i) p[NX + i] = pglob[b * Model::NPG + i];
    double* os = out_states + b * (size_t)(n_steps + 1) * NX;
    double* oc = out_controls + b * (size_t)n_steps * NU;
    for (int i = 0; i < NX; ++i) os[i] = state[i];
    // first guess: scripts start from w0 = 0 (MS:134,149,169) resp. X0 = repmat(state_init)
    if (!single && warm_mode != MPCV_WARM_REFERENCE)
      for (int k = 0; k <= N; ++k) for (int i = 0; i < NX; ++i) w0[k * NZ + i] = state[i];
    int steps = 0, iters_total = 0, worst = 0;
    for (int t = 0; t < n_steps; ++t) {
      if (stop_radius > 0 && Model::NPG >= NX) {
        double d2 = 0;
        for (int i = 0; i < NX; ++i) d2 += (state[i] - p[NX + i]) * (state[i] - p[NX + i]);
        if (!(std::sqrt(d2) > stop_radius)) break;
      }
      for (int i = 0; i < NX; ++i) p[i] = state[i];
      if (Model::NPS > 0)
        for (int k = 0; k < N; ++k)
          for (int i = 0; i < Model::NPS; ++i)
            p[NX + Model::NPG + k * Model::NPS + i] = ptraj[(b * (size_t)(n_steps + N) + t + k) * Model::NPS + i];
      if (warm_mode == MPCV_WARM_COLD) {
        std::fill(w0.begin(), w0.end(), 0.0);
        if (!single) for (int k = 0; k <= N; ++k) for (int i = 0; i < NX; ++i) w0[k * NZ + i] = state[i];
      }
      Ipm<Model> ipm(s, p.data());
      ipm.solve(w0.data(), lbx, ubx);
      ipm.export_solution(lbx, ubx, x.data(), nullptr, nullptr, nullptr, nullptr);
      iters_total += ipm.st.iters;
      if (ipm.st.status != 0 && worst == 0) worst = ipm.st.status;
      // unpack (MS:243-256)
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) Up[k * NU + i] = single ? x[k * NU + i] : x[k * NZ + NX + i];
      if (!single) for (int k = 0; k <= N; ++k) for (int i = 0; i < NX; ++i) Xp[k * NX + i] = x[k * NZ + i];
      for (int i = 0; i < NU; ++i) oc[t * NU + i] = Up[i];
      // plant step with the same integrator
      {
        Ocp<Model> o(s, p.data());
        double xn[NX], q;
        o.step_val(0, state, Up.data(), xn, q);
        for (int i = 0; i < NX; ++i) state[i] = xn[i];
        // reference quirk: `uprev` is a parameter the scripts never update (stays 0),
        // Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:64
        if (Model::HAS_UPREV && warm_mode == MPCV_WARM_REFERENCE) state[NX - 1] = x_init[b * NX + NX - 1];
      }
      for (int i = 0; i < NX; ++i) os[(t + 1) * NX + i] = state[i];
      ++steps;
      // warm start for the next step
      if (warm_mode == MPCV_WARM_SHIFT) {
        if (single) {
          for (int k = 0; k < N; ++k) for (int i = 0; i < NU; ++i) w0[k * NU + i] = Up[std::min(k + 1, N - 1) * NU + i];
        } else {
          for (int k = 0; k <= N; ++k) for (int i = 0; i < NX; ++i) w0[k * NZ + i] = Xp[std::min(k + 1, N) * NX + i];
          for (int k = 0; k < N; ++k) for (int i = 0; i < NU; ++i) w0[k * NZ + NX + i] = Up[std::min(k + 1, N - 1) * NU + i];
        }
      } else if (warm_mode == MPCV_WARM_REFERENCE) {
        if (single) {
          if (s.model == MPCV_MODEL_UNICYCLE_EULER_NODE) {
            // single_shooting_v1.py:173  reshape(u0.T, 2N, 1), column-major => all v then all omega
            for (int k = 0; k < N; ++k) for (int i = 0; i < NU; ++i) w0[i * N + k] = Up[std::min(k + 1, N - 1) * NU + i];
          } else {
            // single_shooting_v2.py:249  reshape(u0, 2N, 1) => interleaved (correct)
            for (int k = 0; k < N; ++k) for (int i = 0; i < NU; ++i) w0[k * NU + i] = Up[std::min(k + 1, N - 1) * NU + i];
          }
        } else {
          // multiple_shooting_casadi.py:279-287  w0 = [vec(shifted X) ; vec(shifted U)]
          int q = 0;
          for (int k = 0; k <= N; ++k) for (int i = 0; i < NX; ++i) w0[q++] = Xp[std::min(k + 1, N) * NX + i];
          for (int k = 0; k < N; ++k) for (int i = 0; i < NU; ++i) w0[q++] = Up[std::min(k + 1, N - 1) * NU + i];
        }
      }
    }
    if (out_steps) out_steps[b] = steps;
    if (out_iters) out_iters[b] = iters_total;
    if (out_status) out_status[b] = worst;
    for (int t = steps; t < n_steps; ++t) {
      for (int i = 0; i < NX; ++i) os[(t + 1) * NX + i] = state[i];
      for (int i = 0; i < NU; ++i) oc[t * NU + i] = 0.0;
    }
  }
  return 0;
}

#define DISPATCH(s, CALL)                                                         \
  switch ((s).model) {                                                            \
    case MPCV_MODEL_UNICYCLE_RK4_QUAD: return CALL(UnicycleRk4Quad);              \
    case MPCV_MODEL_UNICYCLE_EULER_NODE: return CALL(UnicycleEulerNode);          \
    case MPCV_MODEL_UNICYCLE_RK4_NODE: return CALL(UnicycleRk4Node);              \
    case MPCV_MODEL_LINEAR3: return CALL(LinearModel<3 COMMA false>);             \
    case MPCV_MODEL_LINEAR4: return CALL(LinearModel<4 COMMA false>);             \
    case MPCV_MODEL_LINEAR4_DU: return CALL(LinearModel<4 COMMA true>);           \
    case MPCV_MODEL_LINEAR3_DU: return CALL(LinearModel<3 COMMA true>);           \
    case MPCV_MODEL_FRENET_BICYCLE: return CALL(FrenetBicycle);                   \
    default: return -22;                                                          \
  }
#define COMMA ,

}  // namespace

extern "C" {

int mpco_dims(const mpcv_spec* s, int32_t* nx, int32_t* nu, int32_t* n_var, int32_t* n_g, int32_t* n_p,
              int32_t* npg, int32_t* nps) {
#define CALL(M) (dims_of<M>(*s, nx, nu, n_var, n_g, n_p, npg, nps), 0)
  DISPATCH(*s, CALL)
#undef CALL
}

// Batched solve on the host; stats: [B x 8] = inertia corrections, backtracks, SOC trials, mu, err, df, -, -
int mpco_solve(const mpcv_spec* s, const double* x0, const double* lbx, const double* ubx, const double* p,
               double* x, double* f, double* g, double* lam_g, double* lam_x, int32_t* status,
               int32_t* iters, double* stats, int64_t B, int32_t nthreads) {
#define CALL(M) solve_batch<M>(*s, x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, stats, B, nthreads)
  DISPATCH(*s, CALL)
#undef CALL
}

int mpco_rollout(const mpcv_spec* s, const double* p, const double* U, double* X, double* q, int64_t B) {
#define CALL(M) rollout_batch<M>(*s, p, U, X, q, B)
  DISPATCH(*s, CALL)
#undef CALL
}

int mpco_stage_derivs(const mpcv_spec* s, const double* z, const double* pstage, const double* lam,
                      double* xn, double* A, double* Bm, double* q, double* grad, double* H, int64_t B) {
#define CALL(M) stage_derivs_batch<M>(*s, z, pstage, lam, xn, A, Bm, q, grad, H, B)
  DISPATCH(*s, CALL)
#undef CALL
}

int mpco_closed_loop(const mpcv_spec* s, const double* x_init, const double* pglob, const double* ptraj,
                     const double* lbx, const double* ubx, int32_t n_steps, int32_t warm_mode,
                     double stop_radius, double* out_states, double* out_controls, int32_t* out_steps,
                     int32_t* out_iters, int32_t* out_status, int64_t B) {
#define CALL(M) closed_loop_batch<M>(*s, x_init, pglob, ptraj, lbx, ubx, n_steps, warm_mode, stop_radius, \
                                     out_states, out_controls, out_steps, out_iters, out_status, B)
  DISPATCH(*s, CALL)
#undef CALL
}

}  // extern "C"
