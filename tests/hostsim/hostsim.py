"""ctypes loader for the host build of the device headers (development harness; see hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "mpc_verde_b200", "csrc")
_LIB = None


def build(lanes=False):
    """lanes=True: the lane-parallel Riccati factorisation (riccati_factor_lanes) replayed with one lane."""
    so = os.path.join(_HERE, "libhostsim_lanes.so" if lanes else "libhostsim.so")
    deps = [os.path.join(_HERE, "hostsim.cpp")] + [
        os.path.join(_CSRC, f) for f in ("mpcv_models.cuh", "mpcv_ipm.cuh", "mpcv_driver.cuh", "mpcv_phase.cuh",
                                         "mpcv_params.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-ffp-contract=off"] + (["-DMPCV_HOST_LANE_RICCATI"] if lanes else []) +
                              ["-x", "c++", os.path.join(_HERE, "hostsim.cpp"), "-o", so])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def solve(spec, x0, lbx, ubx, p, phased=False, libpath=None):
    """phased=True replays the phase-kernel pipeline's schedule (mpcv_phase.cuh) instead of the
    one-kernel solve; libpath loads another build of the harness (A/B checks of refactors)."""
    L = lib() if libpath is None else C.CDLL(libpath)
    p = _f64(p)
    if p.ndim == 1:
        p = p[None, :]
    B = p.shape[0]
    n, ng = spec.n_var, spec.n_g
    x0 = _f64(np.broadcast_to(np.zeros(n) if x0 is None else x0, (B, n)))
    lbx = _f64(np.broadcast_to(-np.inf if lbx is None else lbx, (n,)))
    ubx = _f64(np.broadcast_to(np.inf if ubx is None else ubx, (n,)))
    x = np.empty((B, n)); f = np.empty(B); g = np.empty((B, ng)); lam_g = np.empty((B, ng)); lam_x = np.empty((B, n))
    status = np.empty(B, np.int32); iters = np.empty(B, np.int32)
    args = (C.byref(spec), _p(x0), _p(lbx), _p(ubx), _p(p), _p(x), _p(f), _p(g), _p(lam_g), _p(lam_x),
            _p(status, C.c_int32), _p(iters, C.c_int32), C.c_long(B))
    out = {"x": x, "f": f, "g": g, "lam_g": lam_g, "lam_x": lam_x, "status": status, "iters": iters}
    if phased:
        sweeps = C.c_int(0)
        rc = L.hs_solve_phased(*args, C.byref(sweeps))
        out["sweeps"] = sweeps.value
    else:
        rc = L.hs_solve(*args)
    assert rc == 0, rc
    return out


def filter_overflows():
    """times the 16-entry filter of the device headers dropped its oldest entry since the harness was loaded"""
    f = lib().hs_filter_overflows
    f.restype = C.c_ulonglong
    return int(f())


def stage_derivs(spec, z, pstage, lam):
    z = _f64(z); lam = _f64(lam)
    B = z.shape[0]
    nx, nu = spec.nx, spec.nu
    nz = nx + nu
    pstage = _f64(pstage if pstage is not None else np.zeros((B, max(spec.npg + spec.nps, 1))))
    xn = np.empty((B, nx)); A = np.empty((B, nx, nx)); Bm = np.empty((B, nx, nu)); q = np.empty(B)
    grad = np.empty((B, nz)); H = np.empty((B, nz, nz))
    rc = lib().hs_stage_derivs(C.byref(spec), _p(z), _p(pstage), _p(lam), _p(xn), _p(A), _p(Bm), _p(q), _p(grad),
                               _p(H), C.c_long(B))
    assert rc == 0, rc
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
    rc = lib().hs_closed_loop(C.byref(spec), _p(x_init), _p(pglob), _p(ptraj), _p(lbx), _p(ubx), C.c_int(n_steps),
                              C.c_int(warm_mode), C.c_double(stop_radius), _p(states), _p(controls),
                              _p(steps, C.c_int32), _p(iters, C.c_int32), _p(status, C.c_int32), C.c_long(B))
    assert rc == 0, rc
    return {"states": states, "controls": controls, "steps": steps, "iters": iters, "status": status}
