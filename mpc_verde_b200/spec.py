"""ctypes mirror of `mpcv_spec` (include/mpcv.h) and the problem presets of the reference scripts.

Pure host-side description; nothing here computes.  Each preset cites the script whose
`nlp_prob = {'f','x','g','p'}` it describes.
"""
import ctypes as C
import math

MODEL_UNICYCLE_RK4_QUAD = 0
MODEL_UNICYCLE_EULER_NODE = 1
MODEL_UNICYCLE_RK4_NODE = 2
MODEL_LINEAR3 = 3
MODEL_LINEAR4 = 4
MODEL_LINEAR4_DU = 5
MODEL_LINEAR3_DU = 6
MODEL_FRENET_BICYCLE = 7

SHOOTING_MULTIPLE = 0
SHOOTING_SINGLE = 1

WARM_SHIFT = 0
WARM_COLD = 1
WARM_REFERENCE = 2

LAYOUT_AUTO = 0
LAYOUT_THREAD = 1
LAYOUT_WARP = 2
LAYOUT_PHASED = 3
LAYOUT_RESIDENT = 4

# IPOPT ApplicationReturnStatus names, as CasADi reports them in solver.stats()['return_status']
STATUS_NAMES = {
    0: "Solve_Succeeded",
    1: "Solved_To_Acceptable_Level",
    -1: "Maximum_Iterations_Exceeded",
    -2: "Restoration_Failed",
    -3: "Error_In_Step_Computation",
    -13: "Invalid_Number_Detected",
}

# (nx, nu, npg, nps) per model — must agree with mpcv_dims()
MODEL_DIMS = {
    MODEL_UNICYCLE_RK4_QUAD: (3, 2, 3, 0),
    MODEL_UNICYCLE_EULER_NODE: (3, 2, 3, 0),
    MODEL_UNICYCLE_RK4_NODE: (3, 2, 0, 5),
    MODEL_LINEAR3: (3, 1, 12, 4),
    MODEL_LINEAR4: (4, 1, 20, 5),
    MODEL_LINEAR4_DU: (5, 1, 20, 5),
    MODEL_LINEAR3_DU: (4, 1, 12, 4),
    MODEL_FRENET_BICYCLE: (4, 2, 0, 4),
}


class Spec(C.Structure):
    _fields_ = [
        ("model", C.c_int32),
        ("shooting", C.c_int32),
        ("N", C.c_int32),
        ("M", C.c_int32),
        ("T", C.c_double),
        ("Q", C.c_double * 4),
        ("R", C.c_double * 2),
        ("R1", C.c_double),
        ("ntu", C.c_int32),
        ("layout", C.c_int32),
        ("tol", C.c_double),
        ("max_iter", C.c_int32),
        ("max_soc", C.c_int32),
        ("mu_init", C.c_double),
        ("bound_push", C.c_double),
        ("bound_frac", C.c_double),
        ("bound_relax_factor", C.c_double),
        ("nlp_scaling_max_gradient", C.c_double),
        ("dual_inf_tol", C.c_double),
        ("constr_viol_tol", C.c_double),
        ("compl_inf_tol", C.c_double),
        ("extra", C.c_double * 4),
        ("acceptable_tol", C.c_double),
        ("acceptable_iter", C.c_int32),
        ("reserved_", C.c_int32),
        ("acceptable_obj_change_tol", C.c_double),
    ]

    # -- derived sizes -----------------------------------------------------------------
    @property
    def nx(self):
        return MODEL_DIMS[self.model][0]

    @property
    def nu(self):
        return MODEL_DIMS[self.model][1]

    @property
    def npg(self):
        return MODEL_DIMS[self.model][2]

    @property
    def nps(self):
        return MODEL_DIMS[self.model][3]

    @property
    def single(self):
        return self.shooting == SHOOTING_SINGLE

    @property
    def n_var(self):
        return self.nu * self.N if self.single else self.nx * (self.N + 1) + self.nu * self.N

    @property
    def n_g(self):
        return self.nx * (self.N + 1)

    @property
    def n_p(self):
        return self.nx + self.npg + self.N * self.nps

    def copy(self, **kw):
        s = Spec.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            setattr(s, k, v)
        return s


# options of the scripts' `opts['ipopt']` dicts (Casadi/single_shooting_v1.py:121-129) the solver implements, and
# the ones that do not change the result (printing)
IPOPT_OPTIONS = ("tol", "max_iter", "max_soc", "mu_init", "bound_push", "bound_frac", "bound_relax_factor",
                 "nlp_scaling_max_gradient", "dual_inf_tol", "constr_viol_tol", "compl_inf_tol", "acceptable_tol",
                 "acceptable_iter", "acceptable_obj_change_tol")
IPOPT_INERT = ("print_level", "sb", "print_timing_statistics", "max_cpu_time")


def ipopt_defaults(s, opts=None):
    """IPOPT 3.12 defaults; `opts` is the {'ipopt': {...}} dict of the scripts
    (Casadi/single_shooting_v1.py:121-129)."""
    s.tol = 1e-8
    s.max_iter = 3000
    s.max_soc = 4
    s.mu_init = 0.1
    s.bound_push = 1e-2
    s.bound_frac = 1e-2
    s.bound_relax_factor = 1e-8
    s.nlp_scaling_max_gradient = 100.0
    s.dual_inf_tol = 1.0
    s.constr_viol_tol = 1e-4
    s.compl_inf_tol = 1e-4
    s.acceptable_tol = 1e-6
    s.acceptable_iter = 15
    s.acceptable_obj_change_tol = 1e20
    if opts:
        ip = opts.get("ipopt", opts)
        for k in ("tol", "max_iter", "max_soc", "mu_init", "bound_push", "bound_frac",
                  "bound_relax_factor", "nlp_scaling_max_gradient", "dual_inf_tol",
                  "constr_viol_tol", "compl_inf_tol", "acceptable_tol", "acceptable_iter",
                  "acceptable_obj_change_tol"):
            if k in ip:
                setattr(s, k, ip[k])
        unknown = set(ip) - set(IPOPT_OPTIONS) - set(IPOPT_INERT) - {"ipopt", "layout", "print_time"}
        if unknown:
            import warnings
            warnings.warn("ipopt options ignored by the GPU solver: %s" % sorted(unknown), stacklevel=3)
    return s


def _base(model, shooting, N, T, M, Q, R, R1=0.0, ntu=0, opts=None, extra=()):
    s = Spec()
    s.model, s.shooting, s.N, s.M, s.T = model, shooting, N, M, T
    for i, q in enumerate(Q):
        s.Q[i] = q
    for i, r in enumerate(R):
        s.R[i] = r
    s.R1 = R1
    s.ntu = ntu
    s.layout = LAYOUT_AUTO
    for i, e in enumerate(extra):
        s.extra[i] = e
    return ipopt_defaults(s, opts)


# The scripts' constants: Casadi/single_shooting_v1.py:29-47
UNICYCLE_Q = (1.0, 5.0, 0.1)
UNICYCLE_R = (0.5, 0.05)
V_MAX = 1.0
OMEGA_MAX = math.pi / 4


def unicycle_multiple_shooting(N=10, T=0.2, M=4, Q=UNICYCLE_Q, R=UNICYCLE_R, opts=None):
    """Casadi/multiple_shooting_casadi.py:29-187 — 3+5N variables, 3(N+1) equalities."""
    return _base(MODEL_UNICYCLE_RK4_QUAD, SHOOTING_MULTIPLE, N, T, M, Q, R, opts=opts)


def unicycle_single_shooting_rk4(N=10, T=0.2, M=4, Q=UNICYCLE_Q, R=UNICYCLE_R, opts=None):
    """Casadi/single_shooting_v2.py:29-167 — 2N variables, RK4 M=4 + quadrature cost."""
    return _base(MODEL_UNICYCLE_RK4_QUAD, SHOOTING_SINGLE, N, T, M, Q, R, opts=opts)


def unicycle_single_shooting_euler(N=10, T=0.2, Q=UNICYCLE_Q, R=UNICYCLE_R, opts=None):
    """Casadi/single_shooting_v1.py:29-119 — 2N variables, Euler rollout, node-sum cost."""
    return _base(MODEL_UNICYCLE_EULER_NODE, SHOOTING_SINGLE, N, T, 1, Q, R, opts=opts)


def unicycle_tracking(N=10, T=0.2, M=1, Q=(1.0, 1.0, 0.1), R=UNICYCLE_R, opts=None):
    """Trajectory Tracking/Trajectory_tracking.py:15-72 (MPCTools nmpc, RK4 M=1, per-stage p).
    mpctools/multiple_shooting_mpctools.py:48-64 is Q=(1,5,0.1), R=(1,1), p=(goal,0,0)."""
    return _base(MODEL_UNICYCLE_RK4_NODE, SHOOTING_MULTIPLE, N, T, M, Q, R, opts=opts)


def linear_tracking(nx, N, Q, R, T=0.0, R1=None, ntu=0, opts=None):
    """MPCTools linear trackers: Trajectory_tracking_lateral_error.py:17-75 (nx=3),
    Trajectory_tracking_dynamic_model.py:17-141 (nx=4),
    Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:10-64 (nx=4, Du cost)."""
    if R1 is None and ntu > 0:
        # move blocking needs u_prev in the state (MPCTools: Du[t >= Ntu] = 0): the Du model with a zero rate weight,
        # e.g. Trajectory_tracking_lateral_error.py:17-18 (Ntu = 3, no Du cost).  The plain models would ignore ntu.
        R1 = 0.0
    if R1 is None:
        model = {3: MODEL_LINEAR3, 4: MODEL_LINEAR4}[nx]
        R1 = 0.0
    else:
        model = {3: MODEL_LINEAR3_DU, 4: MODEL_LINEAR4_DU}[nx]
    return _base(model, SHOOTING_MULTIPLE, N, T, 1, Q, (R, 0.0), R1=R1, ntu=ntu, opts=opts)


def frenet_bicycle(N=20, T=0.05, M=1, L=3.5, lam=(2.5, 1.75, 2.5, 0.4, 10.0), opts=None):
    """Trajectory Tracking/test2.py:20-59,103-122; lam = (lambda1..lambda5)."""
    l1, l2, l3, l4, l5 = lam
    return _base(MODEL_FRENET_BICYCLE, SHOOTING_MULTIPLE, N, T, M, (l2, l3, l1, l5), (0.0, l4),
                 opts=opts, extra=(L, N + 1.0))
