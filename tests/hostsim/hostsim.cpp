// hostsim.cpp — DEVELOPMENT HARNESS, not a product path and not a fallback.
// Compiles the device headers (mpcv_models.cuh / mpcv_ipm.cuh / mpcv_driver.cuh) as plain
// C++ with one lane per problem so that the interior-point logic can be unit-tested against
// the oracle in a container without a GPU.  Nothing in mpc_verde_b200/ loads this library;
// the shipped library (libmpcv.so) contains CUDA kernels only and errors out without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mpc_verde_b200/csrc/mpcv_driver.cuh"
#include "../../mpc_verde_b200/csrc/mpcv_phase.cuh"
#include "../../mpc_verde_b200/csrc/mpcv_params.h"

using namespace mpcv;

static unsigned long long g_diag[4] = {0, 0, 0, 0};

template <class Model, bool SINGLE>
static int hs_solve_t(const mpcv_spec* s, const SolveIO& io, long B) {
  Params P = params_from_spec(*s);
  P.diag = g_diag;
  Layout L = make_layout<Model, SINGLE>(s->N);
  std::vector<double> buf(L.total);
  for (long b = 0; b < B; ++b) {
    std::fill(buf.begin(), buf.end(), 0.0);
    solve_problem<Model, SINGLE, 1, WsDense>(P, L, WsDense{buf.data()}, Grp<1>(0), io, b);
  }
  return 0;
}

// The phase-kernel pipeline (mpcv_phase.cuh) replayed sequentially: the same per-thread bodies in the
// same order as the CUDA graph issues them, including the ping-pong active lists.
template <class Model>
static int hs_solve_phased_t(const mpcv_spec* s, const SolveIO& io, long B, int* sweeps_out) {
  using Ph = Phase<Model, WsDense>;
  Params P = params_from_spec(*s);
  P.diag = g_diag;
  Layout L = make_layout<Model, false>(s->N);
  std::vector<double> slab((size_t)L.total * B, 0.0);
  auto ws = [&](long b) { return WsDense{slab.data() + (size_t)b * L.total}; };
  std::vector<BndEntry> tab(L.n);
  {
    Ipm<Model, false, 1, WsDense> ipm(P, L, WsDense{nullptr}, Grp<1>(0), io.lbx, io.ubx, nullptr);
    for (int i = 0; i < L.n; ++i) tab[i] = ipm.bnd_entry(i);
  }
  std::vector<int> act[2];
  act[0].resize(B); act[1].resize(B);
  int n_act[2] = {(int)B, 0}, sweep = 0;
  for (long b = 0; b < B; ++b) { act[0][b] = (int)b; Ph::init_body(P, L, ws(b), io, b, tab.data(), 0); }
  for (int k = 0; k < L.N; ++k) for (long b = 0; b < B; ++b) Ph::der_body(P, L, ws(b), io, k, false, tab.data());
  for (long b = 0; b < B; ++b) Ph::init2_body(P, L, ws(b), io, tab.data());
  for (int k = 0; k < L.N; ++k) for (long b = 0; b < B; ++b) Ph::der_body(P, L, ws(b), io, k, true, tab.data());
  std::vector<int> retry, slow;
  const Grp<1> g1(0);
  // (the GPU leaves the sweeps at <= kTailBelow problems and finishes them in ph_tail_kernel; the replay keeps
  // sweeping — same phase functions, same results per problem)
  while (n_act[sweep & 1] > 0) {
    const int in = sweep & 1, out = in ^ 1;
    retry.clear(); slow.clear();
    for (int e = 0; e < n_act[in]; ++e) {                                   // pre
      const int b = act[in][e];
      if (Ph::pre_body(P, L, ws(b), io, b, tab.data(), g1, 0)) act[out][n_act[out]++] = b;
    }
    for (int e = 0; e < n_act[out]; ++e)                                    // factor
      if (!Ph::factor_body(P, L, ws(act[out][e]), io, tab.data())) retry.push_back(act[out][e]);
    int won[Ph::kProbe + 1] = {};
    for (int b : retry) {                                                    // probe (kProbe attempts, no stores)
      bool found = false;
      double dsel = 0.0;
      int a = 0;
      for (; a < Ph::kProbe && !found; ++a) {
        double dw;
        found = Ph::probe_body(P, L, ws(b), io, tab.data(), a, &dw);
        dsel = dw;
      }
      won[found ? a - 1 : Ph::kProbe]++;
      if (found) Ph::apply_body(P, L, ws(b), io, tab.data(), dsel);     // the winning lane completes the step
      else ws(b)[L.st + kSlotDwHint] = -dsel;
    }
    for (int b : retry) Ph::retry_body(P, L, ws(b), io, b, tab.data(), 0);  // retry
    for (int e = 0; e < n_act[out]; ++e)                                    // post
      if (Ph::running(L, ws(act[out][e]))) Ph::post_body(P, L, ws(act[out][e]), io, tab.data(), g1);
    for (int k = 0; k < L.N; ++k)                                           // trial
      for (int e = 0; e < n_act[out]; ++e)
        if (Ph::running(L, ws(act[out][e]))) Ph::trial_body(P, L, ws(act[out][e]), io, k, tab.data());
    for (int e = 0; e < n_act[out]; ++e) {                                  // accept
      const int b = act[out][e];
      if (Ph::running(L, ws(b)) && !Ph::accept_body(P, L, ws(b), io, tab.data(), g1)) slow.push_back(b);
    }
    for (int b : slow) Ph::slow_body(P, L, ws(b), io, b, tab.data(), g1, 0);   // slow
    for (int k = 0; k < L.N; ++k)                                           // der
      for (int e = 0; e < n_act[out]; ++e) {
        const int b = act[out][e];
        if (Ph::running(L, ws(b))) Ph::der_body(P, L, ws(b), io, k, true, tab.data());
      }
    if (getenv("HS_STATS")) {
      // schedule statistics the kernel shapes were chosen from: active, wrong inertia (winning probe attempt), slow path
      fprintf(stderr, "sweep %2d active %6d retry %6zu won", sweep, n_act[out], retry.size());
      for (int a = 0; a <= Ph::kProbe; ++a) fprintf(stderr, " %d", won[a]);
      fprintf(stderr, " slow %zu\n", slow.size());
    }
    n_act[in] = 0;
    ++sweep;
  }
  if (sweeps_out) *sweeps_out = sweep;
  return 0;
}

template <class Model, bool SINGLE>
static int hs_loop_t(const mpcv_spec* s, const LoopIO& io, long B) {
  Params P = params_from_spec(*s);
  P.diag = g_diag;
  Layout L = make_layout<Model, SINGLE>(s->N);
  std::vector<double> buf(L.total);
  for (long b = 0; b < B; ++b) {
    std::fill(buf.begin(), buf.end(), 0.0);
    closed_loop_problem<Model, SINGLE, 1, WsDense>(P, L, WsDense{buf.data()}, Grp<1>(0), io, b);
  }
  return 0;
}

template <class Model>
static int hs_der_t(const mpcv_spec* s, const double* z, const double* pstage, const double* lam, double* xn,
                    double* A, double* Bm, double* q, double* grad, double* H, long B) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, NW = NZ * (NZ + 1) / 2;
  Params P = params_from_spec(*s);
  P.diag = g_diag;
  for (long b = 0; b < B; ++b) {
    const double* pp = pstage + b * (Model::NPG + Model::NPS);
    double W[NW], xv[NX], qv;
    Model::der(P, z + b * NZ, z + b * NZ + NX, pp, pp + Model::NPG, lam + b * NX, 1.0, true, xn + b * NX,
               A + b * NX * NX, Bm + b * NX * NU, q + b, grad + b * NZ, W);
    for (int i = 0; i < NZ; ++i)
      for (int j = 0; j < NZ; ++j) H[b * NZ * NZ + i * NZ + j] = i >= j ? W[tri(i, j)] : W[tri(j, i)];
    // value-only path must agree with the derivative path
    Model::val(P, z + b * NZ, z + b * NZ + NX, pp, pp + Model::NPG, xv, &qv);
    for (int i = 0; i < NX; ++i)
      if (xv[i] != xn[b * NX + i] && !(fabs(xv[i] - xn[b * NX + i]) <= 1e-13 * (1 + fabs(xv[i])))) return -1;
    if (!(fabs(qv - q[b]) <= 1e-12 * (1 + fabs(qv)))) return -2;
  }
  return 0;
}

#define HS_DISPATCH(s, SINGLE_OK, CALL)                                                        \
  switch ((s)->model) {                                                                        \
    case MPCV_MODEL_UNICYCLE_RK4_QUAD: return CALL(Unicycle<0>);                               \
    case MPCV_MODEL_UNICYCLE_EULER_NODE: return CALL(Unicycle<1>);                             \
    case MPCV_MODEL_UNICYCLE_RK4_NODE: return CALL(Unicycle<2>);                               \
    case MPCV_MODEL_LINEAR3: return CALL(Linear<3 HS_COMMA false>);                            \
    case MPCV_MODEL_LINEAR4: return CALL(Linear<4 HS_COMMA false>);                            \
    case MPCV_MODEL_LINEAR4_DU: return CALL(Linear<4 HS_COMMA true>);                          \
    case MPCV_MODEL_LINEAR3_DU: return CALL(Linear<3 HS_COMMA true>);                          \
    case MPCV_MODEL_FRENET_BICYCLE: return CALL(FrenetBicycle);                               \
    default: return -22;                                                                       \
  }
#define HS_COMMA ,

extern "C" {

// diagnostic counters of the device headers (Params::diag): [0] filter overflows since the library was loaded
unsigned long long hs_filter_overflows() { return g_diag[0]; }

int hs_solve(const mpcv_spec* s, const double* x0, const double* lbx, const double* ubx, const double* p,
             double* x, double* f, double* g, double* lam_g, double* lam_x, int* status, int* iters, long B) {
  SolveIO io{x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, nullptr};
  if (s->shooting == MPCV_SHOOTING_SINGLE) {
#define CALL(M) hs_solve_t<M, true>(s, io, B)
    HS_DISPATCH(s, 1, CALL)
#undef CALL
  } else {
#define CALL(M) hs_solve_t<M, false>(s, io, B)
    HS_DISPATCH(s, 1, CALL)
#undef CALL
  }
}

int hs_solve_phased(const mpcv_spec* s, const double* x0, const double* lbx, const double* ubx, const double* p,
                    double* x, double* f, double* g, double* lam_g, double* lam_x, int* status, int* iters, long B,
                    int* sweeps) {
  SolveIO io{x0, lbx, ubx, p, x, f, g, lam_g, lam_x, status, iters, nullptr};
  if (s->shooting == MPCV_SHOOTING_SINGLE) return -22;
#define CALL(M) hs_solve_phased_t<M>(s, io, B, sweeps)
  HS_DISPATCH(s, 1, CALL)
#undef CALL
}

int hs_closed_loop(const mpcv_spec* s, const double* x_init, const double* pglob, const double* ptraj,
                   const double* lbx, const double* ubx, int n_steps, int warm_mode, double stop_radius,
                   double* out_states, double* out_controls, int* out_steps, int* out_iters, int* out_status, long B) {
  LoopIO io{x_init, pglob, ptraj, lbx, ubx, out_states, out_controls, out_steps, out_iters, out_status,
            n_steps, warm_mode, stop_radius};
  if (s->shooting == MPCV_SHOOTING_SINGLE) {
#define CALL(M) hs_loop_t<M, true>(s, io, B)
    HS_DISPATCH(s, 1, CALL)
#undef CALL
  } else {
#define CALL(M) hs_loop_t<M, false>(s, io, B)
    HS_DISPATCH(s, 1, CALL)
#undef CALL
  }
}

int hs_stage_derivs(const mpcv_spec* s, const double* z, const double* pstage, const double* lam, double* xn,
                    double* A, double* Bm, double* q, double* grad, double* H, long B) {
#define CALL(M) hs_der_t<M>(s, z, pstage, lam, xn, A, Bm, q, grad, H, B)
  HS_DISPATCH(s, 1, CALL)
#undef CALL
}
}
