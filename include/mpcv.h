/*
 * mpcv.h — C ABI of the B200-native batched nonlinear-MPC solver.
 *
 * Drop-in boundary for ONE hot path of gabrielhaj/mpc-verde: the
 * `casadi.nlpsol('ipopt')` solve over an RK4/Euler/c2d shooting rollout and the
 * shift()/closed loop around it.  The reference binds no FFI of its own (it is
 * pure Python calling CasADi); each entry point below names the reference call
 * site it replaces.  See INTEGRATION.md for the ctypes stub a maintainer of the
 * reference scripts would add.
 *
 *   reference call                                             replaced by
 *   ---------------------------------------------------------  -----------------
 *   ca.nlpsol('solver','ipopt',prob,opts)                      mpcv_create
 *       Casadi/single_shooting_v1.py:121-131
 *       Casadi/single_shooting_v2.py:167-177
 *       Casadi/multiple_shooting_casadi.py:181-197
 *       mpc.nmpc(...)  e.g. Trajectory Tracking/Trajectory_tracking.py:72
 *   sol = solver(x0=,lbx=,ubx=,lbg=,ubg=,p=)                   mpcv_solve / mpcv_solve_host
 *       Casadi/single_shooting_v1.py:174-181
 *       Casadi/single_shooting_v2.py:212-219
 *       Casadi/multiple_shooting_casadi.py:235-242
 *       solver.solve() Trajectory Tracking/Trajectory_tracking.py:107
 *   F(x0=[x;ref], p=u) / ff(U,P) rollout                       mpcv_rollout
 *       Casadi/multiple_shooting_casadi.py:98-114, single_shooting_v1.py:85-95
 *   while ‖state-target‖>0.1: solve, plant step, shift          mpcv_closed_loop
 *       Casadi/single_shooting_v1.py:164-214 (+shift_timestep 17-27)
 *       Casadi/single_shooting_v2.py:201-266
 *       Casadi/multiple_shooting_casadi.py:224-298
 *
 * Conventions
 *   - FP64 throughout.  Row-major, batch dimension leading: problem b's decision
 *     vector is x[b*n_var .. (b+1)*n_var).
 *   - mpcv_solve / mpcv_rollout / mpcv_closed_loop take DEVICE pointers and are
 *     asynchronous on `stream`; the *_host variants take HOST pointers, stage
 *     them through pinned buffers and synchronise before returning.
 *   - Return 0 on success, negative errno-style code on failure;
 *     mpcv_last_error() holds a message.  Nothing throws across the ABI.
 *   - Per-problem solver status mirrors IPOPT's ApplicationReturnStatus.
 *   - One handle per GPU; a handle is not thread-safe, and calls through one handle are ordered: a call
     issued on another stream first waits (on the device) for the previous call of the handle to finish,
     because the workspaces, lists and graphs are owned by the handle.
 *
 * Decision-vector layouts (identical to the reference scripts')
 *   multiple shooting (multiple_shooting_casadi.py:116-178):
 *       w = [X0, U0, X1, U1, ..., U_{N-1}, X_N]            n_var = nx*(N+1)+nu*N
 *       g = [xbar - X0 ; F(X_k,U_k) - X_{k+1}] k=0..N-1     n_g   = nx*(N+1)
 *   single shooting (single_shooting_v1.py:109-111, v2 :127-147):
 *       w = [U0, U1, ..., U_{N-1}]                          n_var = nu*N
 *       g = predicted states vec(X) (inert, bounds +-inf)   n_g   = nx*(N+1)
 *   parameter vector p per problem:
 *       [ xbar (nx) ; problem-global params (npg) ; stage params (N * nps) ]
 *     unicycle point stabilisation: npg=3 (target), nps=0  ->  p = [x;ref] exactly
 *     as `args['p']` in the scripts.
 */
#ifndef MPCV_H_
#define MPCV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- models ------------------------------------ */
enum {
  /* unicycle, RK4 with M sub-steps and cost quadrature (multiple_shooting_casadi.py:68-114,
     single_shooting_v2.py:68-113); npg = 3 (target), nps = 0 */
  MPCV_MODEL_UNICYCLE_RK4_QUAD = 0,
  /* unicycle, forward Euler, node-sum cost (single_shooting_v1.py:70-105); npg=3, nps=0 */
  MPCV_MODEL_UNICYCLE_EULER_NODE = 1,
  /* unicycle, RK4(M) on the state only, node cost against per-stage references
     (mpctools/multiple_shooting_mpctools.py:37-55, Trajectory_tracking.py:40-61); npg=0, nps=5 */
  MPCV_MODEL_UNICYCLE_RK4_NODE = 2,
  /* linear x+ = A x + B u with per-problem (A,B) (c2d or RK4-of-linear done by the caller),
     tracking cost sum_i Q_i (x_i - r_i)^2 + R (u - r_u)^2; nx = 3 (lateral-error bicycle,
     Trajectory_tracking_lateral_error.py:33-55) */
  MPCV_MODEL_LINEAR3 = 3,
  /* same, nx = 4 (dynamic bicycle, Trajectory_tracking_dynamic_model.py:51-55,119-134) */
  MPCV_MODEL_LINEAR4 = 4,
  /* linear nx=4 with input-increment cost R1*(u-u_prev)^2 and move blocking; the state is
     augmented with u_prev (Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:19-64) */
  MPCV_MODEL_LINEAR4_DU = 5,
  /* linear nx=3 with input-increment cost (Trjectory_tracking_le_LTV.py:54-55) */
  MPCV_MODEL_LINEAR3_DU = 6,
  /* Frenet kinematic bicycle, RK4(M) (Trajectory Tracking/test2.py:42-51,103-118); steering
     increment is the control so |d delta| <= bound is a simple box; npg=0, nps=4 */
  MPCV_MODEL_FRENET_BICYCLE = 7,
  MPCV_MODEL_COUNT = 8
};

enum { MPCV_SHOOTING_MULTIPLE = 0, MPCV_SHOOTING_SINGLE = 1 };

/* IPOPT ApplicationReturnStatus values used per problem */
enum {
  MPCV_SOLVE_SUCCEEDED = 0,
  MPCV_SOLVED_TO_ACCEPTABLE_LEVEL = 1,
  MPCV_MAXIMUM_ITERATIONS_EXCEEDED = -1,
  MPCV_RESTORATION_FAILED = -2,
  MPCV_ERROR_IN_STEP_COMPUTATION = -3,
  MPCV_INVALID_NUMBER_DETECTED = -13
};

/* warm-start / closed-loop modes */
enum {
  MPCV_WARM_SHIFT = 0,      /* correctly interleaved shift-by-one guess (default) */
  MPCV_WARM_COLD = 1,       /* X_k = current state, U = 0 every step */
  MPCV_WARM_REFERENCE = 2   /* replay the script's own (scrambled) guess layout:
                               multiple_shooting_casadi.py:284-287, single_shooting_v1.py:173 */
};

/* closed-loop flags (mpcv_loop_args.flags) */
enum {
  MPCV_LOOP_X0_FROM_PREDICTION = 1,  /* next x0 = the solver's predicted x_1 instead of the plant state:
                                        solver.fixvar("x",0,solver.var["x",1]), Trajectory_tracking.py:111-112 */
  MPCV_LOOP_PTRAJ_WINDOWS = 2        /* ptraj holds one full horizon window per MPC step, [B x n_steps x N x nps]
                                        (the par[:, k, t] tables of the scripts), instead of one trajectory
                                        [B x (n_steps+N) x nps] read through a sliding window */
};

/* kernel layout selection */
enum {
  MPCV_LAYOUT_AUTO = 0,
  MPCV_LAYOUT_THREAD = 1,   /* one problem per thread, SoA workspace in HBM/L2 */
  MPCV_LAYOUT_WARP = 2,     /* one warp per problem, stage-parallel, shared-memory workspace */
  MPCV_LAYOUT_PHASED = 3,   /* batch-synchronous phase kernels over the thread layout's workspace: one
                               thread per (problem, interval) for the derivative / trial sweeps, one
                               thread per problem for the Riccati recursion, active-list compaction,
                               one CUDA graph (conditional WHILE) per solve; multiple shooting only */
  MPCV_LAYOUT_RESIDENT = 4  /* persistent CTAs that keep the workspaces of their problems in shared memory and walk
                               the phases of an iteration in lock step; finished problems are replaced from a
                               global queue (continuous batching).  Only x0 / p in and x / f out touch HBM.
                               What MPCV_LAYOUT_AUTO picks when an SM holds >= 8 problems; multiple shooting only */
};

typedef struct mpcv_spec {
  int32_t model;            /* MPCV_MODEL_* */
  int32_t shooting;         /* MPCV_SHOOTING_* */
  int32_t N;                /* horizon length */
  int32_t M;                /* RK4 sub-steps per interval (ignored by Euler / linear) */
  double  T;                /* sampling time (interval length) */
  double  Q[4];             /* state weights (diagonal) */
  double  R[2];             /* control weights (diagonal) */
  double  R1;               /* input-increment weight (DU models) */
  int32_t ntu;              /* move blocking: u free for k < ntu, u_k = u_{k-1} after; 0 = off */
  int32_t layout;           /* MPCV_LAYOUT_* */
  /* IPOPT-named options (defaults of IPOPT 3.12 when 0 is passed; see mpcv_spec_defaults) */
  double  tol;                       /* 1e-8 */
  int32_t max_iter;                  /* 3000 (scripts use 2000) */
  int32_t max_soc;                   /* 4 */
  double  mu_init;                   /* 0.1 */
  double  bound_push;                /* 1e-2 */
  double  bound_frac;                /* 1e-2 */
  double  bound_relax_factor;        /* 1e-8 */
  double  nlp_scaling_max_gradient;  /* 100 */
  double  dual_inf_tol;              /* 1 */
  double  constr_viol_tol;           /* 1e-4 */
  double  compl_inf_tol;             /* 1e-4 */
  double  extra[4];                  /* model constants: FRENET: [L, Nt+1 divisor, -, -] */
  /* acceptable-level termination (opts of Casadi/single_shooting_v1.py:121-129); 0 = IPOPT default */
  double  acceptable_tol;             /* 1e-6 */
  int32_t acceptable_iter;            /* 15; negative switches the test off */
  int32_t reserved_;
  double  acceptable_obj_change_tol;  /* 1e20 */
} mpcv_spec;

typedef struct mpcv_handle mpcv_handle;

/* fill `s` with IPOPT defaults and the unicycle MS N=10 T=0.2 problem of the scripts */
void mpcv_spec_defaults(mpcv_spec* s);

/* sizes implied by a spec; any out pointer may be NULL */
int mpcv_dims(const mpcv_spec* s, int32_t* nx, int32_t* nu, int32_t* n_var, int32_t* n_g,
              int32_t* n_p, int32_t* npg, int32_t* nps);

mpcv_handle* mpcv_create(const mpcv_spec* s);      /* replaces ca.nlpsol(...) */
void         mpcv_destroy(mpcv_handle* h);
const char*  mpcv_last_error(void);

/* Batched solve, device pointers.  lbx/ubx: [n_var] shared by the batch (+-inf allowed,
   |b| >= 1e19 counts as infinite like IPOPT's nlp_lower_bound_inf).  lbg/ubg of the
   reference call are fixed by the transcription (0 for defects, +-inf for inert rows)
   and therefore not arguments.  Nullable outputs: g, lam_g, lam_x, status, iters. */
int mpcv_solve(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx,
               const double* p, double* x, double* f, double* g, double* lam_g, double* lam_x,
               int32_t* status, int32_t* iters, int64_t B, void* stream);

/* Same with host pointers (H2D, solve, D2H, synchronise). */
int mpcv_solve_host(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx,
                    const double* p, double* x, double* f, double* g, double* lam_g,
                    double* lam_x, int32_t* status, int32_t* iters, int64_t B);

/* Shooting rollout of a control sequence: p [B x n_p], U [B x nu*N] ->
   X [B x nx*(N+1)], q [B] (accumulated cost; nullable). Device pointers. */
int mpcv_rollout(mpcv_handle* h, const double* p, const double* U, double* X, double* q,
                 int64_t B, void* stream);

/* Stage derivatives for parity tests of the hand-written forward/adjoint sweeps:
   z [B x (nx+nu)], pstage [B x (npg+nps)], lam [B x nx]  ->
   xn [B x nx], A [B x nx*nx], Bm [B x nx*nu], q [B], grad [B x (nx+nu)],
   H [B x (nx+nu)^2] = Hessian of q + lam' * phi.  Device pointers. */
int mpcv_stage_derivs(mpcv_handle* h, const double* z, const double* pstage, const double* lam,
                      double* xn, double* A, double* Bm, double* q, double* grad, double* H,
                      int64_t B, void* stream);

/* Batched closed loop (solve -> apply u0 -> plant step -> shift), device pointers.
   x_init [B x nx]; pglob [B x npg]; ptraj [B x (n_steps+N) x nps] (NULL when nps = 0);
   out_states [B x (n_steps+1) x nx]; out_controls [B x n_steps x nu];
   out_steps [B] number of MPC iterations actually run; out_iters [B] summed IPM iterations.
   stop_radius > 0: stop a problem when ||x - target||_2 <= stop_radius (scripts: 1e-1). */
int mpcv_closed_loop(mpcv_handle* h, const double* x_init, const double* pglob,
                     const double* ptraj, const double* lbx, const double* ubx,
                     int32_t n_steps, int32_t warm_mode, double stop_radius,
                     double* out_states, double* out_controls, int32_t* out_steps,
                     int32_t* out_iters, int32_t* out_status, int64_t B, void* stream);

/* Extended closed loop: everything mpcv_closed_loop takes, plus per-step model parameters (the LTV scripts
   re-discretise Ac(c[t]) / Ac(vref[t]) every step: Trjectory_tracking_le_LTV.py:126-143,
   Trajectory_tracking_dynamic_model.py:119-141), the x0-from-prediction mode of the MPCTools loops, per-step
   parameter windows, and the histories every script keeps: the predicted horizon of every
   solve (`cat_states`, single_shooting_v1.py:185-188) and the duration of every MPC step (`times`, :209-212).
   Scenarios whose loop has stopped (stop_radius) are no longer solved.  Device pointers; unused ones NULL. */
typedef struct mpcv_loop_args {
  const double* x_init;        /* [B x nx] */
  const double* pglob;         /* [B x npg] constant model parameters */
  const double* pglob_traj;    /* [B x n_steps x npg] model parameters of step t (overrides pglob) */
  const double* ptraj;         /* [B x (n_steps+N) x nps], or [B x n_steps x N x nps] with MPCV_LOOP_PTRAJ_WINDOWS */
  const double* lbx;           /* [n_var] */
  const double* ubx;           /* [n_var] */
  int32_t n_steps, warm_mode, flags, reserved_;
  double stop_radius;
  double* out_states;          /* [B x (n_steps+1) x nx] plant states */
  double* out_controls;        /* [B x n_steps x nu] applied controls */
  int32_t* out_steps;          /* [B] */
  int32_t* out_iters;          /* [B] */
  int32_t* out_status;         /* [B] */
  double* out_horizons;        /* [B x n_steps x (N+1) x nx] predicted states of every solve */
  long long* out_step_ns;      /* [n_steps] device-timer nanoseconds of every MPC step of the batch */
} mpcv_loop_args;
int mpcv_closed_loop_ex(mpcv_handle* h, const mpcv_loop_args* a, int64_t B, void* stream);

/* Batched solve with per-problem bounds: lbx / ubx are [B x n_var] (the reference call takes one lbx / ubx per
   solve, single_shooting_v1.py:174-181, so a batch of calls may carry different boxes).  lbg / ubg of the reference
   call are fixed by the transcription; lam_p is not provided.  Otherwise as mpcv_solve. */
int mpcv_solve_bounds(mpcv_handle* h, const double* x0, const double* lbx, const double* ubx, const double* p,
                      double* x, double* f, double* g, double* lam_g, double* lam_x, int32_t* status,
                      int32_t* iters, int64_t B, void* stream);

/* Batched exact zero-order hold (replaces mpc.util.c2d, Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:24,
   Trajectory Tracking/Trjectory_tracking_le_LTV.py:126-133, Trajectory_tracking_dynamic_model.py:134):
   [A Bd; 0 I] = expm([Ac Bc; 0 0] dt) for B systems.  Ac [B x n x n], Bc [B x n x nu] -> A [B x n x n],
   Bd [B x n x nu], row-major device pointers; n + nu <= 6. */
int mpcv_c2d(int32_t n, int32_t nu, double dt, const double* Ac, const double* Bc, double* A, double* Bd,
             int64_t B, void* stream);

/* ---- reference-trajectory pipeline on the device -------------------------------------
   The scripts build their per-stage parameter tables and per-step models in nested Python loops right before the
   solve; these entry points build them for B scenarios at once, in HBM, in the layouts mpcv_closed_loop_ex reads.
   A scenario is the base path (x, y) [nsim] stretched by scale[b] = (sx, sy) (NULL: unscaled).  Device pointers. */

/* par[:, k, t] of the lateral-error trackers (Trajectory_tracking_lateral_error.py:94-116, Phiref.py:124-155):
   (y_ref, phi_ref, r_ref, delta_ref) -> pwin [B x nsim x Nt x 4]  (MPCV_LOOP_PTRAJ_WINDOWS layout) */
int mpcv_ref_lateral(const double* x, const double* y, int32_t nsim, const double* scale, int32_t Nt, double Delta,
                     double ar, double br, double* pwin, int64_t B, void* stream);
/* p[k, :] of the Frenet bicycle (test2.py:79-100; p[2] / p[3] swapped as in the script) -> pwin [B x n_steps x Nt x 4] */
int mpcv_ref_frenet(const double* x, const double* y, const double* vdes, int32_t nsim, const double* scale, int32_t Nt,
                    double Delta, int32_t n_steps, double* pwin, int64_t B, void* stream);
/* (x, y, theta, v, omega) references of the unicycle tracker cut from a path sampled every dt (central differences,
   v / omega clipped to the control box) -> ptraj [B x T x 5]  (sliding-window layout) */
int mpcv_ref_unicycle_path(const double* x, const double* y, int32_t T, const double* scale, double dt, double vmax,
                           double wmax, double* ptraj, int64_t B, void* stream);
/* the circle reference of Trajectory_tracking.py:84-97 -> ptraj [T x 5] */
int mpcv_ref_circle(int32_t T, double Delta, double* ptraj, void* stream);
/* lane_change.py:5-79: the base path (a, b, c) [n0] extended by arcs and straights -> (xt, yt, c2) [*n_out], the
   columns of out.csv.  a_end / b_end: the last base sample (host values).  xt = NULL only sizes the output. */
int mpcv_path_lane_change_ext(const double* a, const double* b, const double* c, int32_t n0, double a_end, double b_end,
                              double v, double dt, double* xt, double* yt, double* c2, int32_t capacity, int32_t* n_out,
                              void* stream);
/* per-step exact ZOH of the LTV models -> pglob_traj [B x n_steps x npg] (= [A row-major, B]):
   lateral-error bicycle, u_ref = c[t] * spd[b] (Trjectory_tracking_le_LTV.py:126-133), npg = 12;
   dynamic bicycle in v = v[t] (per_scenario: v is [B x T]) with params = (m, a, b, Ca, Jz) or NULL for the script's
   (Trajectory_tracking_dynamic_model.py:37-43,119-134, the A34 precedence as written), npg = 20 */
int mpcv_ltv_lateral(const double* c, int32_t T, const double* spd, double ar, double br, double dt, int32_t n_steps,
                     double* pglob_traj, int64_t B, void* stream);
int mpcv_ltv_dynbike(const double* v, int32_t T, int32_t per_scenario, const double* params, double dt, int32_t n_steps,
                     double* pglob_traj, int64_t B, void* stream);

/* Measured FP64 FMA peak of the current device in TFLOP/s (register-resident DFMA chains);
   the roofline denominator of bench.py. */
int mpcv_fp64_peak(double* tflops, double* ms, void* stream);

/* Optional per-problem latency capture: when `dev_ns` (device pointer, [B] int64) is non-NULL the
   solve kernels stamp %globaltimer at each problem's entry and convergence exit and store the
   difference in nanoseconds (the "p50 solve us" of BASELINE.json). NULL switches it off. */
int mpcv_set_latency_buffer(mpcv_handle* h, long long* dev_ns);

/* number of kernel launches issued through this handle since creation (for the phased layout the
   iteration sweeps run inside a CUDA graph: 12 kernel nodes per sweep, see mpcv_phase_sweeps) */
int64_t mpcv_launch_count(const mpcv_handle* h);

/* phased layout (synchronises `stream`): interior-point sweeps executed by the last mpcv_solve, and the
   number of kernels run through this handle since creation INCLUDING the 12 kernel nodes of every
   graph-driven sweep (which mpcv_launch_count cannot see from the host).  Either out may be NULL. */
int mpcv_phase_sweeps(mpcv_handle* h, int32_t* sweeps, int64_t* kernel_nodes, void* stream);

/* Execution knobs of a handle, to be set before its first solve (-EBUSY afterwards); the MPCV_* environment variables
   of DESIGN.md set the same values for every handle of the process.
     "phase_pipes"     1..8   independent pipes a batch is split into (default 4, never below phase_pipe_min problems
                              per pipe).  A caller that keeps several batches in flight on several handles wants 1: the
                              batches then sit in different phases and share the GPU better than the lock-step pipes of
                              one batch (C2, 4 in flight: 11.5 ms per batch with 1 pipe, 12.0 with 2; one batch alone
                              with 4 pipes: 14.1).
     "phase_pipe_min"  >= 32  smallest share of a pipe
     "tail_below", "tail_shift"   hand-off of the sweeps to the straggler tail at min(B >> shift, below) active problems
     "resident_below"  AUTO layout: batches smaller than this run in the CTA-resident kernel */
int mpcv_set_knob(mpcv_handle* h, const char* name, int64_t value);

/* Diagnostic counters accumulated by the kernels since the handle was created (synchronises the device):
   counters4[0] = filter overflows.  IPOPT's line-search filter (the reference's solver, IpFilter.cpp) is unbounded;
   the device filter holds 16 entries after IPOPT's own pruning of dominated entries and drops the oldest when full.
   A non-zero count says some solve of this handle left IPOPT's iteration path for that reason (never on the
   BASELINE configurations; about 1 in 2,000 of the all-zeros-guess problems of SURVEY Appendix E).
   counters4[1..3] are reserved (0). */
int mpcv_diag(mpcv_handle* h, uint64_t* counters4);

#ifdef __cplusplus
}
#endif
#endif /* MPCV_H_ */
