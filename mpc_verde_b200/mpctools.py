"""MPCTools-shaped front end: the call surface the reference's `mpctools` scripts use around the same
IPOPT solve (SURVEY.md §8b "secondary boundary", Appendix C), with a leading batch dimension.

    import mpc_verde_b200.mpctools as mpc
    solver = mpc.nmpc(model, N={"x": Nx, "u": Nu, "t": Nt, "p": Np}, x0=x0, lb=lb, ub=ub, p=p, uprev=uprev)
    solver.par["p", k] = vec          # Trajectory Tracking/Trajectory_tracking.py:105-106
    solver.fixvar("x", 0, x0)         # Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:73
    solver.solve()                    # Trajectory_tracking.py:107
    solver.stats["status"]            # :110
    solver.saveguess()                # :111  (guess := solution shifted by one stage)
    solver.var["x", 1], solver.var["u", 0, :], solver.var["x", :, :]     # :112-118
    sol = mpc.callSolver(solver)      # {"x": (Nt+1, Nx), "u": (Nt, Nu), "Du", "status", "obj"}

What differs from MPCTools, and why: `nmpc` takes a structural `model` (one of `problems.*`, which name the
reference script and lines they transcribe) instead of CasADi functions f and l — the GPU solver is a
fixed-structure solver for the scripts' shooting transcriptions and cannot consume a symbolic graph.
Everything else keeps MPCTools' meaning: the objective sums the stage cost over t = 0..Nt-1 with no terminal
term, x[0] is fixed to x0, `lb`/`ub` boxes apply to every stage, `Du[t] = u[t] - u[t-1]` with u[-1] = `uprev`
(a parameter the scripts never update), `lb = ub = 0` on `Du[t >= Ntu]` is move blocking, failures are reported
in `stats["status"]` and never raised.  Every array may carry a leading batch dimension B (x0 of shape (B, Nx),
parameters (B, Np)); without it results are unbatched, exactly like the scripts.
"""
import numpy as np

from . import problems as _problems
from . import spec as _S

c2d = _problems.c2d                      # mpc.util.c2d (Inverted_pendulum/...:24)


class util:                              # noqa: N801  (MPCTools spells it mpc.util.c2d)
    c2d = staticmethod(_problems.c2d)


class _Par:
    def __init__(self, owner):
        self._o = owner

    def __setitem__(self, key, val):
        name, k = key
        if name != "p":
            raise KeyError(name)
        v = np.asarray(val, dtype=np.float64)
        o = self._o
        if v.ndim == 2 and v.shape[-1] == 1:
            v = v[:, 0]                   # CasADi column vector
        o._stage[..., k, :] = v            # broadcasts over the batch

    def __getitem__(self, key):
        name, k = key
        if name != "p":
            raise KeyError(name)
        return self._o._unb(self._o._stage[..., k, :])


class _Var:
    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, key):
        o = self._o
        if o._sol is None:
            raise RuntimeError("solve() first")
        name, idx = key[0], key[1:]
        arr = {"x": o._sol["x"], "u": o._sol["u"], "Du": o._sol["Du"]}[name]
        out = arr[(slice(None),) + tuple(idx)]
        return o._unb(out)


class ControlSolver:
    """What `mpc.nmpc(...)` returns in the scripts (the subset of its surface they use)."""

    def __init__(self, prob, nx_user, x0, lb, ub, p, uprev, pglob, opts):
        from . import solver as _solver          # nlpsol raises without a CUDA device: there is no CPU path
        self._solver = _solver.nlpsol("solver", "ipopt", prob, opts or {"ipopt": {"print_level": 0}})
        sp = self.spec = self._solver.spec
        self._du = sp.nx != nx_user                    # Du models carry u_prev as an extra state
        self._nxu = nx_user
        x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
        self._batched = x0.ndim == 2
        x0 = np.atleast_2d(x0)
        B = self._B = x0.shape[0]
        self._x0 = x0.copy()
        up = np.zeros(1) if uprev is None else np.asarray(uprev, dtype=np.float64).reshape(-1)
        self._uprev = np.broadcast_to(up.reshape(-1, 1) if up.size == B else up.reshape(1, 1), (B, 1)).copy()
        self._pglob = np.zeros((B, 0)) if sp.npg == 0 else np.broadcast_to(np.asarray(pglob, dtype=np.float64), (B, sp.npg)).copy()
        self._stage = np.zeros((B, sp.N, max(sp.nps, 1)))[:, :, :sp.nps]
        if p is not None and sp.nps:
            self._stage[:] = np.broadcast_to(np.asarray(p, dtype=np.float64), self._stage.shape)
        self._lbx, self._ubx = self._bounds(lb or {}, ub or {})
        self._guess = np.zeros((B, sp.n_var))          # MPCTools' default guess: zeros except x[0]
        self._sol = None
        self.par = _Par(self)
        self.var = _Var(self)
        self.stats = {"status": None}

    # -- helpers -----------------------------------------------------------------------------------
    def _unb(self, a):
        return a if self._batched else a[0]

    def _bounds(self, lb, ub):
        """lb/ub dicts with keys "x", "u", "Du" as in the scripts (Inverted_pendulum/...:34-42,
        Trajectory_tracking.py:64-67, test2.py:55-59)."""
        sp = self.spec
        nz = sp.nx + sp.nu
        lo = np.full(sp.n_var, -np.inf)
        hi = np.full(sp.n_var, np.inf)
        for d, dst, sign in ((lb, lo, -1), (ub, hi, +1)):
            if "x" in d:
                v = np.asarray(d["x"], dtype=np.float64).reshape(-1)
                for k in range(sp.N + 1):
                    dst[k * nz:k * nz + self._nxu] = v
            if "u" in d:
                v = np.asarray(d["u"], dtype=np.float64).reshape(-1)
                for k in range(sp.N):
                    dst[k * nz + sp.nx:k * nz + sp.nx + sp.nu] = v
        # Du bounds.  The scripts use them two ways: 'free, then pinned to 0' = move blocking
        # (Inverted_pendulum/...:34-42, Trajectory_tracking_lateral_error.py) and a finite rate box on a model whose
        # control IS the increment (test2.py:55-59, problems.frenet_bounds).  Anything else would be silently
        # dropped by a fixed-structure solver, so it is refused.
        if ("Du" in lb) != ("Du" in ub):
            raise NotImplementedError("Du bounds must be given in both lb and ub")
        if "Du" in lb:
            if sp.nu != 1:
                raise NotImplementedError("Du bounds are supported for single-input models only (Frenet bicycle: "
                                          "use problems.frenet_bounds, where the increment is the control)")
            dl = np.asarray(lb["Du"], dtype=np.float64).reshape(-1)
            du = np.asarray(ub["Du"], dtype=np.float64).reshape(-1)
            if dl.size != sp.N * sp.nu or du.size != sp.N * sp.nu:
                raise ValueError("Du bounds must have Nt*Nu = %d entries" % (sp.N * sp.nu))
            pinned = (dl == 0) & (du == 0)
            free = np.isneginf(dl) & np.isposinf(du)
            if not np.all(pinned | free):
                raise NotImplementedError("finite Du rate bounds are outside the hot path (only 'free, then pinned "
                                          "to 0' = move blocking)")
            ntu = int(np.argmax(pinned)) if pinned.any() else 0
            if pinned.any() and not pinned[ntu:].all():
                raise NotImplementedError("Du bounds other than 'free, then pinned to 0' are outside the hot path")
            if ntu != self.spec.ntu:
                raise ValueError("model was built with ntu=%d but the Du bounds pin from t=%d" % (self.spec.ntu, ntu))
        elif self.spec.ntu > 0:
            raise ValueError("model was built with ntu=%d but no Du bounds pin the later moves" % self.spec.ntu)
        return lo, hi

    def _x_aug(self):
        return np.concatenate([self._x0, self._uprev], 1) if self._du else self._x0

    # -- the MPCTools surface ------------------------------------------------------------------------
    def fixvar(self, name, t, val):
        """lb = ub = guess = val; the scripts only ever fix x[0]."""
        if name != "x" or t != 0:
            raise NotImplementedError("only fixvar('x', 0, value) is on the hot path")
        self._x0[:] = np.asarray(val, dtype=np.float64).reshape(self._x0.shape if np.ndim(val) == 2 else (1, -1))

    def saveguess(self, toffset=1):
        """Copy the solution into the guess shifted by `toffset` stages (MPCTools default 1)."""
        if self._sol is None:
            return
        sp = self.spec
        nz = sp.nx + sp.nu
        w = self._sol["w"]
        g = w.copy()
        for k in range(sp.N + 1):
            ks = min(k + toffset, sp.N)
            g[:, k * nz:k * nz + sp.nx] = w[:, ks * nz:ks * nz + sp.nx]
            if k < sp.N:
                ku = min(k + toffset, sp.N - 1)
                g[:, k * nz + sp.nx:(k + 1) * nz] = w[:, ku * nz + sp.nx:(ku + 1) * nz]
        self._guess = g

    def solve(self):
        sp = self.spec
        nz = sp.nx + sp.nu
        xa = self._x_aug()
        g = self._guess.copy()
        g[:, :sp.nx] = xa                                  # x[0] is fixed
        p = np.concatenate([xa, self._pglob, self._stage.reshape(self._B, -1)], 1)
        sol = self._solver(x0=g, lbx=self._lbx, ubx=self._ubx, p=p, outputs=("x", "f"))
        st = self._solver.stats()
        w = np.atleast_2d(sol["x"])
        X = np.stack([w[:, k * nz:k * nz + self._nxu] for k in range(sp.N + 1)], 1)
        U = np.stack([w[:, k * nz + sp.nx:(k + 1) * nz] for k in range(sp.N)], 1)
        uprev = self._uprev[:, None, :] if sp.nu == 1 else np.zeros((self._B, 1, sp.nu))
        Du = np.diff(np.concatenate([np.broadcast_to(uprev, (self._B, 1, sp.nu)), U], 1), axis=1)
        self._sol = {"w": w, "x": X, "u": U, "Du": Du, "f": np.atleast_1d(sol["f"])}
        names = st["return_status"]
        self.stats = {"status": names if self._batched else (names if isinstance(names, str) else names[0]),
                      "iter_count": st["iter_count"], "success": st["success"]}
        return self.stats["status"]


def nmpc(model=None, l=None, N=None, x0=None, lb=None, ub=None, p=None, uprev=None, pglob=None, opts=None, **kw):
    """`mpc.nmpc(f, l, N, x0, lb, ub, p=..., uprev=..., funcargs=..., isQP=..., verbosity=...)` of the scripts
    (Inverted_pendulum/...:64, Trajectory_tracking.py:72, Phiref.py:174) with `model` — a `problems.*`
    template — in place of the CasADi functions f and l.  `pglob`: per-problem model constants (the (A, B) of a
    linear model, row-major A then B).  MPCTools-only keywords (funcargs, inferargs, isQP, verbosity,
    timelimit) are accepted and ignored: they do not change the optimum."""
    if not (isinstance(model, dict) and "spec" in model):
        raise TypeError("nmpc(model=...) takes one of mpc_verde_b200.problems.* (the GPU solver is fixed-structure)")
    sp = model["spec"]
    if N is not None:
        if int(N["t"]) != sp.N:
            raise ValueError("N['t']=%d but the model was built with N=%d" % (N["t"], sp.N))
        nx_user = int(N["x"])
    else:
        nx_user = sp.nx
    return ControlSolver(model, nx_user, x0, lb, ub, p, uprev, pglob, opts)


def callSolver(solver):                   # noqa: N802  (MPCTools' spelling)
    """`sol = mpc.callSolver(solver)` (Inverted_pendulum/...:74-77)."""
    status = solver.solve()
    s = solver._sol
    un = solver._unb
    return {"x": un(s["x"]), "u": un(s["u"]), "Du": un(s["Du"]), "status": status, "obj": un(s["f"])}
