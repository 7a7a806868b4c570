"""Host-side mirror of the reference's solver interface, with a leading batch dimension.

Reference call shape kept (Casadi/multiple_shooting_casadi.py:181-242):

    solver = nlpsol('solver', 'ipopt', nlp_prob, opts)          # nlp_prob = {'f','x','g','p'}
    sol = solver(x0=..., lbx=..., ubx=..., lbg=..., ubg=..., p=...)
    sol['x'], sol['f']                                           # (+ 'g','lam_x','lam_g')
    solver.stats()['return_status']

`nlp_prob` is a structural template (problems.py) instead of a CasADi SX graph: the GPU
solver is a fixed-structure solver for the shooting transcriptions of the scripts.  Every
array argument takes a leading batch dimension B; a call without one solves a single problem
and returns unbatched results, exactly like the scripts.

PyTorch is used for device memory and streams only; all compute is in libmpcv.so.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .spec import (LAYOUT_AUTO, SHOOTING_SINGLE, STATUS_NAMES, WARM_SHIFT, Spec, ipopt_defaults)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _addr(t):
    return None if t is None else t.data_ptr()


class _Handle:
    def __init__(self, spec):
        self.lib = _lib.lib()
        self.spec = spec
        self.h = self.lib.mpcv_create(C.byref(spec))
        if not self.h:
            msg = self.lib.mpcv_last_error()
            raise _lib.MpcvError("mpcv_create failed: %s" % (msg.decode() if msg else "?"))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.mpcv_destroy(self.h)
                self.h = None
        except Exception:
            pass


class NlpSolver:
    """Callable returned by `nlpsol`; mirrors the CasADi `Function` the scripts call.

    One call at a time per solver: the workspaces, lists and CUDA graphs belong to the handle, so a call issued
    on another stream waits on the device for the solver's previous call (see include/mpcv.h)."""

    def __init__(self, name, prob, opts=None, device=None):
        if not (isinstance(prob, dict) and "spec" in prob):
            raise TypeError("nlp_prob must come from mpc_verde_b200.problems (keys 'f','x','g','p','spec')")
        self.name = name
        self.prob = prob
        spec = prob["spec"].copy()
        ipopt_defaults(spec, opts)
        if opts and "layout" in opts:
            spec.layout = opts["layout"]
        self.spec = spec
        if not torch.cuda.is_available():
            raise _lib.MpcvError("mpc_verde_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if spec.ntu > 0 and spec.model not in (5, 6, 7):
            raise ValueError("ntu > 0 (move blocking) needs a model with u_prev in its state: build it with "
                             "linear_tracking(..., ntu=K) / R1=..., not the plain linear models")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self._handle = _Handle(spec)
        # execution knobs (include/mpcv.h, mpcv_set_knob): e.g. {"pipes": 1} for a caller that keeps several batches in
        # flight on several solvers
        for key, knob in (("pipes", "phase_pipes"), ("pipe_min", "phase_pipe_min"), ("tail_below", "tail_below"),
                          ("tail_shift", "tail_shift"), ("resident_below", "resident_below")):
            if opts and key in opts:
                self.set_knob(knob, opts[key])
        self._last = None
        self._latency = None

    def set_knob(self, name, value):
        """`mpcv_set_knob`: execution knobs of this solver's handle; before its first solve."""
        _lib.check(self._handle.lib.mpcv_set_knob(self._handle.h, name.encode(), int(value)), "mpcv_set_knob")

    # -- sizes ---------------------------------------------------------------------------
    @property
    def n_var(self):
        return self.spec.n_var

    @property
    def n_g(self):
        return self.spec.n_g

    @property
    def n_p(self):
        return self.spec.n_p

    def launch_count(self):
        return int(self._handle.lib.mpcv_launch_count(self._handle.h))

    def diagnostics(self):
        """Counters the kernels keep since the solver was created: `filter_overflows` (the 16-entry device filter had
        to drop its oldest entry; IPOPT's filter is unbounded)."""
        import ctypes as C
        c = (C.c_uint64 * 4)()
        with torch.cuda.device(self.device):
            _lib.check(self._handle.lib.mpcv_diag(self._handle.h, c), "mpcv_diag")
        return {"filter_overflows": int(c[0])}

    def _phase_counts(self):
        import ctypes as C
        n, k = C.c_int32(0), C.c_int64(0)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(self._handle.lib.mpcv_phase_sweeps(self._handle.h, C.byref(n), C.byref(k), C.c_void_p(st)),
                       "mpcv_phase_sweeps")
        return int(n.value), int(k.value)

    def phase_sweeps(self):
        """Interior-point sweeps the last phased-layout solve ran on the device (each sweep is twelve
        kernel nodes of the solve's CUDA graph).  Synchronises the current stream."""
        return self._phase_counts()[0]

    def kernel_count(self):
        """Kernels run through this handle since creation, graph-driven sweeps included."""
        return self._phase_counts()[1]

    # -- argument normalisation ---------------------------------------------------------------
    @staticmethod
    def _is_per_problem(a):
        return a is not None and not np.isscalar(a) and getattr(a, "ndim", np.ndim(a)) == 2 and np.shape(a)[0] > 1

    def _vec(self, a, n, fill, name, device):
        """bounds: scalar-broadcast allowed (e.g. lbg=-inf, single_shooting_v1.py:141)."""
        if a is None:
            a = fill
        if isinstance(a, torch.Tensor):
            t = a.to(dtype=torch.float64).reshape(-1)
        else:
            t = torch.as_tensor(np.asarray(a, dtype=np.float64).reshape(-1))
        if t.numel() == 1:
            t = t.expand(n)
        if t.numel() != n:
            raise ValueError("%s has %d entries, expected %d" % (name, t.numel(), n))
        return t.contiguous().to(device)

    def _batched(self, a, n, name):
        if isinstance(a, torch.Tensor):
            t = a.to(dtype=torch.float64)
        else:
            t = torch.as_tensor(np.asarray(a, dtype=np.float64))
        if t.dim() == 2 and t.shape[1] == 1 and n != 1:
            t = t.reshape(-1)           # CasADi column vector
        if t.dim() <= 1:
            if t.numel() == 1 and n != 1:
                t = t.reshape(1).expand(n)
            t = t.reshape(1, -1)
            unbatched = True
        else:
            unbatched = False
        if t.shape[-1] != n:
            raise ValueError("%s has trailing size %d, expected %d" % (name, t.shape[-1], n))
        return t.reshape(-1, n), unbatched

    def _check_g_bounds(self, lbg, ubg):
        """The transcription fixes the g-bounds: 0 for the defect rows of multiple shooting
        (multiple_shooting_casadi.py:135-136,174-175), +-inf for the inert rows of single
        shooting (single_shooting_v1.py:141-142).  Anything else is outside the hot path."""
        for b, sign, nm in ((lbg, -1, "lbg"), (ubg, +1, "ubg")):
            if b is None:
                continue
            v = np.asarray(b.cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64).reshape(-1)
            if self.spec.single:
                ok = np.all(v * sign >= 1e19)
            else:
                ok = np.all(v == 0.0)
            if not ok:
                raise NotImplementedError(
                    "%s: only the scripts' own constraint bounds are supported (0 for multiple-shooting defects, "
                    "+-inf for single-shooting rows)" % nm)

    # -- the solve -------------------------------------------------------------------------------
    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, p=None, lam_x0=None, lam_g0=None,
                 outputs=("x", "f", "g", "lam_g", "lam_x")):
        """`outputs` selects which entries of the result dict are produced ('x' and 'f' always are;
        the scripts read only sol['x']); skipping the rest saves their device->host copies."""
        if p is None:
            raise ValueError("p is required")
        if lam_x0 is not None or lam_g0 is not None:
            import warnings
            warnings.warn("lam_x0 / lam_g0 are ignored: like IPOPT without warm_start_init_point the solver starts "
                          "from z = 1 and least-squares constraint multipliers", stacklevel=2)
        self._check_g_bounds(lbg, ubg)
        n, ng, npar = self.spec.n_var, self.spec.n_g, self.spec.n_p
        on_device = isinstance(p, torch.Tensor) and p.is_cuda
        pt, unb = self._batched(p, npar, "p")
        B = pt.shape[0]
        if x0 is None:
            x0t = None
        else:
            x0t, _ = self._batched(x0, n, "x0")
            if x0t.shape[0] == 1 and B > 1:
                x0t = x0t.expand(B, n)
            if x0t.shape[0] != B:
                raise ValueError("x0 batch %d != p batch %d" % (x0t.shape[0], B))
        lib, h = self._handle.lib, self._handle.h
        # one lbx / ubx per problem ([B, n]: a batch of reference calls with different boxes) or one for the batch
        per_problem = self._is_per_problem(lbx) or self._is_per_problem(ubx)
        if isinstance(p, torch.Tensor) and p.is_cuda and p.device != self.device:
            raise ValueError("p lives on %s but the solver was created on %s" % (p.device, self.device))
        if per_problem and not on_device:
            pt, on_device = pt.to(self.device), True          # (the host-pointer C entry takes shared bounds only)
            to_host = True
        else:
            to_host = False
        if on_device:
            dev = pt.device
            with torch.cuda.device(dev):
                pt = pt.contiguous()
                x0d = None if x0t is None else x0t.to(dev).contiguous()
                if per_problem:
                    def rows(a, fill, name):
                        if self._is_per_problem(a):
                            t, _ = self._batched(a, n, name)
                            if t.shape[0] != B:
                                raise ValueError("%s batch %d != p batch %d" % (name, t.shape[0], B))
                            return t.to(dev).contiguous()
                        return self._vec(a, n, fill, name, dev).reshape(1, n).expand(B, n).contiguous()
                    lb, ub = rows(lbx, -math.inf, "lbx"), rows(ubx, math.inf, "ubx")
                else:
                    lb = self._vec(lbx, n, -math.inf, "lbx", dev)
                    ub = self._vec(ubx, n, math.inf, "ubx", dev)
                out = {
                    "x": torch.empty((B, n), dtype=torch.float64, device=dev),
                    "f": torch.empty((B,), dtype=torch.float64, device=dev),
                    "g": torch.empty((B, ng), dtype=torch.float64, device=dev),
                    "lam_g": torch.empty((B, ng), dtype=torch.float64, device=dev),
                    "lam_x": torch.empty((B, n), dtype=torch.float64, device=dev),
                }
                status = torch.empty((B,), dtype=torch.int32, device=dev)
                iters = torch.empty((B,), dtype=torch.int32, device=dev)
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                for k in ("g", "lam_g", "lam_x"):
                    if k not in outputs:
                        del out[k]
                entry = lib.mpcv_solve_bounds if per_problem else lib.mpcv_solve
                rc = entry(h, _ptr(x0d), _ptr(lb), _ptr(ub), _ptr(pt), _ptr(out["x"]), _ptr(out["f"]),
                           _ptr(out.get("g")), _ptr(out.get("lam_g")), _ptr(out.get("lam_x")), _ptr(status),
                           _ptr(iters), C.c_int64(B), stream)
                _lib.check(rc, "mpcv_solve")
                if to_host:
                    out = {k: v.cpu().numpy() for k, v in out.items()}
                    status, iters = status.cpu().numpy(), iters.cpu().numpy()
        else:
            # host buffers: one C-ABI call does H2D, solve, D2H.  Page-locked inputs (torch pin_memory) are
            # copied straight from the caller's memory; the results then land in page-locked arrays cached
            # on the solver (valid until the next call), otherwise in fresh pageable arrays.
            pinned = isinstance(p, torch.Tensor) and p.is_pinned()
            pt = np.ascontiguousarray(pt.numpy())
            x0h = None if x0t is None else np.ascontiguousarray(x0t.numpy())
            lb = np.ascontiguousarray(self._vec(lbx, n, -math.inf, "lbx", "cpu").numpy())
            ub = np.ascontiguousarray(self._vec(ubx, n, math.inf, "ubx", "cpu").numpy())
            shapes = {"x": (B, n), "f": (B,), "g": (B, ng), "lam_g": (B, ng), "lam_x": (B, n)}
            keys = ["x", "f"] + [k for k in ("g", "lam_g", "lam_x") if k in outputs]
            if pinned:
                cache = self.__dict__.setdefault("_pinned_out", {})
                def mk(key, shape, dtype):
                    t = cache.get((key, shape))
                    if t is None:
                        t = cache[(key, shape)] = torch.empty(shape, dtype=dtype).pin_memory()
                    return t.numpy()
                out = {k: mk(k, shapes[k], torch.float64) for k in keys}
                status, iters = mk("status", (B,), torch.int32), mk("iters", (B,), torch.int32)
            else:
                out = {k: np.empty(shapes[k]) for k in keys}
                status = np.empty((B,), np.int32)
                iters = np.empty((B,), np.int32)
            hp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
            with torch.cuda.device(self.device):
                rc = lib.mpcv_solve_host(h, hp(x0h), hp(lb), hp(ub), hp(pt), hp(out["x"]), hp(out["f"]),
                                         hp(out.get("g")), hp(out.get("lam_g")), hp(out.get("lam_x")), hp(status),
                                         hp(iters), C.c_int64(B))
            _lib.check(rc, "mpcv_solve_host")
        self._last = (status, iters)
        if unb:
            out = {k: v[0] for k, v in out.items()}
        return out

    def stats(self):
        """Like CasADi's solver.stats(): per-problem `return_status`, `iter_count`, `success`."""
        if self._last is None:
            return {}
        status, iters = self._last
        st = status.cpu().numpy() if isinstance(status, torch.Tensor) else status
        it = iters.cpu().numpy() if isinstance(iters, torch.Tensor) else iters
        names = [STATUS_NAMES.get(int(s), "Internal_Error") for s in st]
        return {"return_status": names if len(names) > 1 else names[0], "status_code": st,
                "iter_count": it if len(it) > 1 else int(it[0]), "success": bool(np.all(st == 0))}

    # -- latency capture (p50 solve us) ---------------------------------------------------------------
    def enable_latency(self, B):
        self._latency = torch.zeros((B,), dtype=torch.int64, device=self.device)
        _lib.check(self._handle.lib.mpcv_set_latency_buffer(self._handle.h, _ptr(self._latency)), "set_latency")
        return self._latency

    def disable_latency(self):
        _lib.check(self._handle.lib.mpcv_set_latency_buffer(self._handle.h, None), "set_latency")
        self._latency = None

    # -- rollout F / ff -------------------------------------------------------------------------------
    def rollout(self, p, U):
        """Shooting rollout of a control sequence (ff(U,P), single_shooting_v1.py:95; the F loop
        of single_shooting_v2.py:222-224).  Returns X [B, N+1, nx] and the accumulated cost q [B]."""
        s = self.spec
        pt, unb = self._batched(p, s.n_p, "p")
        Ut, _ = self._batched(U, s.nu * s.N, "U")
        B = pt.shape[0]
        dev = self.device
        with torch.cuda.device(dev):
            pd, Ud = pt.to(dev).contiguous(), Ut.to(dev).contiguous()
            X = torch.empty((B, s.N + 1, s.nx), dtype=torch.float64, device=dev)
            q = torch.empty((B,), dtype=torch.float64, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self._handle.lib.mpcv_rollout(self._handle.h, _ptr(pd), _ptr(Ud), _ptr(X), _ptr(q),
                                                     C.c_int64(B), stream), "mpcv_rollout")
        if not (isinstance(p, torch.Tensor) and p.is_cuda):
            X, q = X.cpu().numpy(), q.cpu().numpy()
        return (X[0], q[0]) if unb else (X, q)

    def stage_derivs(self, z, pstage, lam):
        """Hand-written stage sweeps (value, A, B, gradient, Hessian of q + lam'phi) for parity tests."""
        s = self.spec
        nz = s.nx + s.nu
        dev = self.device
        zt = torch.as_tensor(np.asarray(z, dtype=np.float64)).reshape(-1, nz).to(dev).contiguous()
        B = zt.shape[0]
        npp = max(s.npg + s.nps, 1)
        pt = torch.zeros((B, npp), dtype=torch.float64) if pstage is None else torch.as_tensor(np.asarray(pstage, dtype=np.float64)).reshape(B, -1)
        pt = pt.to(dev).contiguous()
        lt = torch.as_tensor(np.asarray(lam, dtype=np.float64)).reshape(B, s.nx).to(dev).contiguous()
        mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
        out = {"xn": mk(B, s.nx), "A": mk(B, s.nx, s.nx), "B": mk(B, s.nx, s.nu), "q": mk(B), "grad": mk(B, nz),
               "H": mk(B, nz, nz)}
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self._handle.lib.mpcv_stage_derivs(self._handle.h, _ptr(zt), _ptr(pt), _ptr(lt), _ptr(out["xn"]),
                                                          _ptr(out["A"]), _ptr(out["B"]), _ptr(out["q"]), _ptr(out["grad"]),
                                                          _ptr(out["H"]), C.c_int64(B), stream), "mpcv_stage_derivs")
        return {k: v.cpu().numpy() for k, v in out.items()}

    # -- batched closed loop ---------------------------------------------------------------------------
    def closed_loop(self, x_init, pglob=None, ptraj=None, lbx=None, ubx=None, n_steps=100, warm_mode=WARM_SHIFT,
                    stop_radius=0.0, pglob_traj=None, x0_from_prediction=False, windows=False, horizons=False,
                    step_times=False):
        """The scripts' MPC loop, batched and on device: solve -> apply u0 -> plant step with the
        same discretisation -> shifted guess (multiple_shooting_casadi.py:224-298,
        single_shooting_v1.py:164-214).  Returns the histories the scripts keep
        (states [B, n_steps+1, nx], applied controls [B, n_steps, nu]) plus per-problem
        step / iteration counts and the first non-zero solver status.

        pglob_traj [B, n_steps, npg]: the model of every step (the LTV scripts re-discretise Ac(c[t]) per step,
        Trjectory_tracking_le_LTV.py:126-143).  x0_from_prediction: the next x0 is the solver's predicted x_1
        (`solver.fixvar("x",0,solver.var["x",1])`, Trajectory_tracking.py:111-112) while the plant is simulated
        on.  windows: ptraj is [B, n_steps, N, nps], one horizon window per step (the scripts' par[:, k, t]).
        horizons / step_times: also return the predicted horizon of every solve (`cat_states`,
        single_shooting_v1.py:185-188) and the device-timed duration of every MPC step in ms (`times`, :209-212)."""
        s = self.spec
        dev = self.device
        to_np = not (isinstance(x_init, torch.Tensor) and x_init.is_cuda)
        xi, unb = self._batched(x_init, s.nx, "x_init")
        B = xi.shape[0]

        def dev_t(a):
            t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64))
            return t.to(dtype=torch.float64)

        with torch.cuda.device(dev):
            xi = xi.to(dev).contiguous()
            pg = pgt = None
            if s.npg > 0 and pglob_traj is not None:
                pgt = dev_t(pglob_traj)
                if pgt.dim() == 2:
                    pgt = pgt.unsqueeze(0)
                if pgt.shape[1] != n_steps or pgt.shape[2] != s.npg:
                    raise ValueError("pglob_traj must be [B, n_steps, npg] = [*, %d, %d]" % (n_steps, s.npg))
                pgt = pgt.expand(B, -1, -1).to(dev).contiguous()
            elif s.npg > 0:
                pg, _ = self._batched(pglob, s.npg, "pglob")
                pg = pg.expand(B, s.npg).to(dev).contiguous()
            pt = None
            if s.nps > 0:
                pt = dev_t(ptraj)
                want = (n_steps, s.N, s.nps) if windows else (n_steps + s.N, s.nps)
                if pt.dim() == len(want):
                    pt = pt.unsqueeze(0)
                if tuple(pt.shape[1:]) != want:
                    raise ValueError("ptraj must be [B, %s]" % ", ".join(str(v) for v in want))
                pt = pt.expand(B, *want).to(dev).contiguous()
            lb = self._vec(lbx, s.n_var, -math.inf, "lbx", dev)
            ub = self._vec(ubx, s.n_var, math.inf, "ubx", dev)
            states = torch.empty((B, n_steps + 1, s.nx), dtype=torch.float64, device=dev)
            controls = torch.empty((B, n_steps, s.nu), dtype=torch.float64, device=dev)
            steps = torch.empty((B,), dtype=torch.int32, device=dev)
            iters = torch.empty((B,), dtype=torch.int32, device=dev)
            status = torch.empty((B,), dtype=torch.int32, device=dev)
            hor = torch.zeros((B, n_steps, s.N + 1, s.nx), dtype=torch.float64, device=dev) if horizons else None
            tns = torch.zeros((n_steps,), dtype=torch.int64, device=dev) if step_times else None
            a = _lib.LoopArgs()
            a.x_init, a.pglob, a.pglob_traj, a.ptraj = xi.data_ptr(), _addr(pg), _addr(pgt), _addr(pt)
            a.lbx, a.ubx = lb.data_ptr(), ub.data_ptr()
            a.n_steps, a.warm_mode, a.stop_radius = n_steps, warm_mode, stop_radius
            a.flags = (1 if x0_from_prediction else 0) | (2 if windows else 0)
            a.out_states, a.out_controls = states.data_ptr(), controls.data_ptr()
            a.out_steps, a.out_iters, a.out_status = steps.data_ptr(), iters.data_ptr(), status.data_ptr()
            a.out_horizons, a.out_step_ns = _addr(hor), _addr(tns)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = self._handle.lib.mpcv_closed_loop_ex(self._handle.h, C.byref(a), C.c_int64(B), stream)
            _lib.check(rc, "mpcv_closed_loop_ex")
        out = {"states": states, "controls": controls, "steps": steps, "iters": iters, "status": status}
        if horizons:
            out["horizons"] = hor
        if step_times:
            out["step_ms"] = tns.to(torch.float64) / 1e6
        if to_np:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out


def nlpsol(name, plugin, nlp_prob, opts=None, device=None):
    """Drop-in for `ca.nlpsol(name, 'ipopt', nlp_prob, opts)` (single_shooting_v1.py:131,
    single_shooting_v2.py:177, multiple_shooting_casadi.py:197)."""
    if plugin != "ipopt":
        raise ValueError("only the 'ipopt' plugin of the reference scripts is provided (got %r)" % (plugin,))
    return NlpSolver(name, nlp_prob, opts, device)


def fp64_peak(device=None):
    """Measured FP64 FMA peak (TFLOP/s) of the current GPU: the roofline denominator."""
    lib = _lib.lib()
    t, ms = C.c_double(0), C.c_double(0)
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.mpcv_fp64_peak(C.byref(t), C.byref(ms), stream), "mpcv_fp64_peak")
    return t.value, ms.value


def c2d(Ac, Bc, dt):
    """Batched exact zero-order hold on the GPU (`mpc.util.c2d` for B systems at once): Ac [B, n, n],
    Bc [B, n, nu] (numpy or CUDA tensors) -> (A, Bd) of the same shapes and kind."""
    on_dev = isinstance(Ac, torch.Tensor) and Ac.is_cuda
    At = torch.as_tensor(Ac, dtype=torch.float64)
    Bt = torch.as_tensor(Bc, dtype=torch.float64)
    unb = At.dim() == 2
    if unb:
        At, Bt = At[None], Bt.reshape(1, At.shape[-1], -1)
    n, nu, B = At.shape[-1], Bt.shape[-1], At.shape[0]
    dev = At.device if on_dev else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        Ad, Bd_in = At.to(dev).contiguous(), Bt.to(dev).contiguous()
        A = torch.empty_like(Ad)
        Bd = torch.empty_like(Bd_in)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().mpcv_c2d(n, nu, float(dt), _ptr(Ad), _ptr(Bd_in), _ptr(A), _ptr(Bd), C.c_int64(B), stream),
                   "mpcv_c2d")
    if not on_dev:
        A, Bd = A.cpu().numpy(), Bd.cpu().numpy()
    return (A[0], Bd[0]) if unb else (A, Bd)
