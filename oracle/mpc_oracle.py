"""ctypes loader for the CPU oracle (oracle/libmpc_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
The product package (mpc_verde_b200) never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmpc_oracle.so")
    src = os.path.join(_HERE, "mpc_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "mpcv.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmpc_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def solve(spec, x0, lbx, ubx, p, nthreads=1, want_stats=False):
    """Batched oracle solve. x0 [B,n_var] or None, lbx/ubx [n_var] or None, p [B,n_p]."""
    p = _f64(p)
    if p.ndim == 1:
        p = p[None, :]
    B = p.shape[0]
    n, ng = spec.n_var, spec.n_g
    assert p.shape[1] == spec.n_p, (p.shape, spec.n_p)
    x0 = _f64(np.broadcast_to(np.zeros(n) if x0 is None else x0, (B, n)))
    lbx = _f64(np.broadcast_to(-np.inf if lbx is None else lbx, (n,)))
    ubx = _f64(np.broadcast_to(np.inf if ubx is None else ubx, (n,)))
    x = np.empty((B, n)); f = np.empty(B); g = np.empty((B, ng)); lam_g = np.empty((B, ng))
    lam_x = np.empty((B, n))
    status = np.empty(B, np.int32); iters = np.empty(B, np.int32); stats = np.zeros((B, 8))
    rc = lib().mpco_solve(C.byref(spec), _p(x0), _p(lbx), _p(ubx), _p(p), _p(x), _p(f), _p(g), _p(lam_g),
                          _p(lam_x), _p(status, C.c_int32), _p(iters, C.c_int32), _p(stats),
                          C.c_int64(B), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError("mpco_solve failed: %d" % rc)
    out = {"x": x, "f": f, "g": g, "lam_g": lam_g, "lam_x": lam_x, "status": status, "iters": iters}
    if want_stats:
        out["stats"] = stats
    return out


def rollout(spec, p, U):
    p = _f64(p); U = _f64(U)
    if p.ndim == 1:
        p, U = p[None, :], U[None, :]
    B = p.shape[0]
    X = np.empty((B, spec.nx * (spec.N + 1))); q = np.empty(B)
    rc = lib().mpco_rollout(C.byref(spec), _p(p), _p(U), _p(X), _p(q), C.c_int64(B))
    if rc != 0:
        raise RuntimeError("mpco_rollout failed: %d" % rc)
    return X.reshape(B, spec.N + 1, spec.nx), q


def stage_derivs(spec, z, pstage, lam):
    z = _f64(z); lam = _f64(lam)
    B = z.shape[0]
    nx, nu = spec.nx, spec.nu
    nz = nx + nu
    pstage = _f64(pstage if pstage is not None else np.zeros((B, max(spec.npg + spec.nps, 1))))
    xn = np.empty((B, nx)); A = np.empty((B, nx, nx)); Bm = np.empty((B, nx, nu)); q = np.empty(B)
    grad = np.empty((B, nz)); H = np.empty((B, nz, nz))
    rc = lib().mpco_stage_derivs(C.byref(spec), _p(z), _p(pstage), _p(lam), _p(xn), _p(A), _p(Bm), _p(q),
                                 _p(grad), _p(H), C.c_int64(B))
    if rc != 0:
        raise RuntimeError("mpco_stage_derivs failed: %d" % rc)
    return {"xn": xn, "A": A, "B": Bm, "q": q, "grad": grad, "H": H}


def closed_loop(spec, x_init, pglob, ptraj, lbx, ubx, n_steps, warm_mode=0, stop_radius=0.0):
    x_init = _f64(x_init)
    if x_init.ndim == 1:
        x_init = x_init[None, :]
    B = x_init.shape[0]
    nx, nu, n = spec.nx, spec.nu, spec.n_var
    pglob = _f64(np.zeros((B, 1)) if pglob is None else np.broadcast_to(pglob, (B, max(spec.npg, 1))))
    ptraj = None if ptraj is None else _f64(np.broadcast_to(ptraj, (B, n_steps + spec.N, spec.nps)))
    lbx = _f64(np.broadcast_to(-np.inf if lbx is None else lbx, (n,)))
    ubx = _f64(np.broadcast_to(np.inf if ubx is None else ubx, (n,)))
    states = np.empty((B, n_steps + 1, nx)); controls = np.empty((B, n_steps, nu))
    steps = np.empty(B, np.int32); iters = np.empty(B, np.int32); status = np.empty(B, np.int32)
    rc = lib().mpco_closed_loop(C.byref(spec), _p(x_init), _p(pglob), _p(ptraj), _p(lbx), _p(ubx),
                                C.c_int32(n_steps), C.c_int32(warm_mode), C.c_double(stop_radius),
                                _p(states), _p(controls), _p(steps, C.c_int32), _p(iters, C.c_int32),
                                _p(status, C.c_int32), C.c_int64(B))
    if rc != 0:
        raise RuntimeError("mpco_closed_loop failed: %d" % rc)
    return {"states": states, "controls": controls, "steps": steps, "iters": iters, "status": status}
