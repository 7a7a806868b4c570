"""Reference-trajectory pipeline on the device: the tables the scripts build in nested
Python loops right before the solve — per-stage parameters `par[:, k, t]`, the extended lane-change path, the
per-step (A, B) of the LTV models — computed by batched kernels (csrc/mpcv_ref.cu) for B scenarios at once and left
in HBM in the layouts `NlpSolver.closed_loop` reads (`windows=True` tables, sliding-window trajectories,
`pglob_traj`).  A scenario is the base path stretched by `scale[b] = (sx, sy)` (the CSV path scaled in
speed and laterally).

    par   = reference.lateral_windows(x, y, Nt, Delta)                       # Phiref.py:124-155
    pgt   = reference.ltv_lateral(c, Delta, n_steps)                         # Trjectory_tracking_le_LTV.py:126-133
    out   = solver.closed_loop(x0, pglob_traj=pgt, ptraj=par, windows=True, ...)

There is no CPU path: everything here needs a CUDA device.  Inputs may be numpy arrays or CUDA tensors; outputs are
CUDA tensors (pass them on to `closed_loop` without a host round trip).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .problems import LATERAL_AR, LATERAL_BR


def _dev(a, device):
    if a is None:
        return None
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    return t.to(device=device, dtype=torch.float64).contiguous()


def _device(device):
    if not torch.cuda.is_available():
        raise _lib.MpcvError("mpc_verde_b200.reference needs a CUDA device; there is no CPU fallback")
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _batch(scale):
    return 1 if scale is None else int(np.shape(scale)[0])


def lateral_windows(x, y, Nt, Delta, scale=None, ar=LATERAL_AR, br=LATERAL_BR, device=None):
    """(y_ref, phi_ref, r_ref, delta_ref) of the lateral-error trackers, the scripts' `par[:, k, t]`
    (Trajectory_tracking_lateral_error.py:94-116, Phiref.py:124-155) -> [B, Nsim, Nt, 4] (`windows=True` layout)."""
    dev = _device(device)
    with torch.cuda.device(dev):
        xd, yd, sc = _dev(x, dev), _dev(y, dev), _dev(scale, dev)
        B, nsim = _batch(scale), xd.numel()
        out = torch.empty((B, nsim, Nt, 4), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ref_lateral(_p(xd), _p(yd), nsim, _p(sc), Nt, float(Delta), float(ar), float(br),
                                               _p(out), C.c_int64(B), _stream(dev)), "mpcv_ref_lateral")
    return out


def frenet_windows(x, y, vdes, Nt, Delta, n_steps, scale=None, device=None):
    """Per-stage parameters of the Frenet bicycle as test2.py:79-100 fills them — p[2] = vdes and p[3] = |(xdd, ydd)|,
    which the model then unpacks as (kappa_t, vdes): swapped in the script and kept -> [B, n_steps, Nt, 4]."""
    dev = _device(device)
    with torch.cuda.device(dev):
        xd, yd, vd, sc = _dev(x, dev), _dev(y, dev), _dev(vdes, dev), _dev(scale, dev)
        B, nsim = _batch(scale), xd.numel()
        out = torch.empty((B, n_steps, Nt, 4), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ref_frenet(_p(xd), _p(yd), _p(vd), nsim, _p(sc), Nt, float(Delta), n_steps, _p(out),
                                              C.c_int64(B), _stream(dev)), "mpcv_ref_frenet")
    return out


def unicycle_path_reference(x, y, dt, scale=None, vmax=1.0, wmax=np.pi / 4, device=None):
    """(x, y, theta, v, omega) references of the unicycle tracker (Trajectory_tracking.py:54-61 parameter layout) cut
    from a path sampled every dt -> [B, T, 5], the sliding-window layout (`ptraj[b, t:t+N]` is step t's window)."""
    dev = _device(device)
    with torch.cuda.device(dev):
        xd, yd, sc = _dev(x, dev), _dev(y, dev), _dev(scale, dev)
        B, T = _batch(scale), xd.numel()
        out = torch.empty((B, T, 5), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ref_unicycle_path(_p(xd), _p(yd), T, _p(sc), float(dt), float(vmax), float(wmax),
                                                     _p(out), C.c_int64(B), _stream(dev)), "mpcv_ref_unicycle_path")
    return out


def circle_reference(T, Delta, device=None):
    """Trajectory_tracking.py:84-97: unit circle at 0.1 rad/s, v_ref = omega_ref = 1 -> [T, 5] (sliding window)."""
    dev = _device(device)
    with torch.cuda.device(dev):
        out = torch.empty((T, 5), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ref_circle(T, float(Delta), _p(out), _stream(dev)), "mpcv_ref_circle")
    return out


def lane_change_extended(a, b, c, v=0.6, dt=0.05, device=None):
    """lane_change.py:5-79: the lane change (a, b, c) = (x, y, uref) followed by a half turn, a straight, an S of two
    half-radius half turns, the straight back and a closing half turn; uref = v on the appended part.
    Returns (x_t, y_t, c2) — the columns of out.csv — as CUDA tensors."""
    dev = _device(device)
    a_h = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b_h = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    with torch.cuda.device(dev):
        ad, bd, cd = _dev(a, dev), _dev(b, dev), _dev(c, dev)
        n0 = ad.numel()
        n = C.c_int32(0)
        args = (n0, float(a_h[-1]), float(b_h[-1]), float(v), float(dt))
        lib = _lib.lib()
        _lib.check(lib.mpcv_path_lane_change_ext(None, None, None, *args, None, None, None, 0, C.byref(n), None),
                   "mpcv_path_lane_change_ext")
        xt, yt, c2 = (torch.empty((n.value,), dtype=torch.float64, device=dev) for _ in range(3))
        _lib.check(lib.mpcv_path_lane_change_ext(_p(ad), _p(bd), _p(cd), *args, _p(xt), _p(yt), _p(c2), n.value,
                                                 C.byref(n), _stream(dev)), "mpcv_path_lane_change_ext")
    return xt, yt, c2


def ltv_lateral(c, Delta, n_steps, speed_scale=None, ar=LATERAL_AR, br=LATERAL_BR, device=None):
    """Exact ZOH of Ac(u_ref = c[t]) for every step (Trjectory_tracking_le_LTV.py:126-133) -> pglob_traj
    [B, n_steps, 12] = [A row-major, B] of MODEL_LINEAR3 / LINEAR3_DU."""
    dev = _device(device)
    with torch.cuda.device(dev):
        cd, sp = _dev(c, dev), _dev(speed_scale, dev)
        B = 1 if speed_scale is None else sp.numel()
        out = torch.empty((B, n_steps, 12), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ltv_lateral(_p(cd), cd.numel(), _p(sp), float(ar), float(br), float(Delta), n_steps,
                                               _p(out), C.c_int64(B), _stream(dev)), "mpcv_ltv_lateral")
    return out


def ltv_dynamic_bicycle(v, Delta, n_steps, params=None, device=None):
    """Exact ZOH of the dynamic bicycle linearised at v = v[t] (Trajectory_tracking_dynamic_model.py:37-43,119-134;
    A34 with the operator precedence as written) -> pglob_traj [B, n_steps, 20].  v: [T] shared or [B, T]."""
    dev = _device(device)
    with torch.cuda.device(dev):
        vd = _dev(v, dev)
        per = vd.dim() == 2
        B, T = (vd.shape[0], vd.shape[1]) if per else (1, vd.numel())
        ph = None if params is None else (C.c_double * 5)(*[float(q) for q in np.asarray(params).reshape(-1)[:5]])
        out = torch.empty((B, n_steps, 20), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mpcv_ltv_dynbike(_p(vd), T, 1 if per else 0, ph, float(Delta), n_steps, _p(out),
                                               C.c_int64(B), _stream(dev)), "mpcv_ltv_dynbike")
    return out
