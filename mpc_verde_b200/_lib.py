"""ctypes binding of libmpcv.so (the C ABI of include/mpcv.h).

There is no CPU path: if the CUDA library is missing or no GPU is usable, every entry point
raises.  The library is built in-tree by `__graft_entry__.build()` / `make -C mpc_verde_b200/csrc`.
"""
import ctypes as C
import os

from .spec import Spec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCV_LIB") or os.path.join(_HERE, "libmpcv.so")   # MPCV_LIB: tuning builds
_LIB = None

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)
_V = C.c_void_p

# every symbol include/mpcv.h declares, with its signature
SIGNATURES = {
    "mpcv_spec_defaults": (None, [C.POINTER(Spec)]),
    "mpcv_dims": (C.c_int, [C.POINTER(Spec)] + [_I] * 7),
    "mpcv_create": (_V, [C.POINTER(Spec)]),
    "mpcv_destroy": (None, [_V]),
    "mpcv_last_error": (C.c_char_p, []),
    "mpcv_solve": (C.c_int, [_V] + [_V] * 4 + [_V] * 5 + [_V, _V, C.c_int64, _V]),
    "mpcv_solve_host": (C.c_int, [_V] + [_V] * 4 + [_V] * 5 + [_V, _V, C.c_int64]),
    "mpcv_rollout": (C.c_int, [_V, _V, _V, _V, _V, C.c_int64, _V]),
    "mpcv_stage_derivs": (C.c_int, [_V] + [_V] * 9 + [C.c_int64, _V]),
    "mpcv_closed_loop": (C.c_int, [_V] + [_V] * 5 + [C.c_int32, C.c_int32, C.c_double] + [_V] * 5 + [C.c_int64, _V]),
    "mpcv_fp64_peak": (C.c_int, [_D, _D, _V]),
    "mpcv_set_latency_buffer": (C.c_int, [_V, _V]),
    "mpcv_launch_count": (C.c_int64, [_V]),
    "mpcv_diag": (C.c_int, [_V, C.POINTER(C.c_uint64)]),
    "mpcv_set_knob": (C.c_int, [_V, C.c_char_p, C.c_int64]),
    "mpcv_c2d": (C.c_int, [C.c_int32, C.c_int32, C.c_double, _V, _V, _V, _V, C.c_int64, _V]),
    "mpcv_phase_sweeps": (C.c_int, [_V, C.POINTER(C.c_int32), C.POINTER(C.c_int64), _V]),
    "mpcv_closed_loop_ex": (C.c_int, [_V, _V, C.c_int64, _V]),
    "mpcv_solve_bounds": (C.c_int, [_V] + [_V] * 4 + [_V] * 5 + [_V, _V, C.c_int64, _V]),
    "mpcv_ref_lateral": (C.c_int, [_V, _V, C.c_int32, _V, C.c_int32, C.c_double, C.c_double, C.c_double, _V, C.c_int64, _V]),
    "mpcv_ref_frenet": (C.c_int, [_V, _V, _V, C.c_int32, _V, C.c_int32, C.c_double, C.c_int32, _V, C.c_int64, _V]),
    "mpcv_ref_unicycle_path": (C.c_int, [_V, _V, C.c_int32, _V, C.c_double, C.c_double, C.c_double, _V, C.c_int64, _V]),
    "mpcv_ref_circle": (C.c_int, [C.c_int32, C.c_double, _V, _V]),
    "mpcv_path_lane_change_ext": (C.c_int, [_V, _V, _V, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, _V, _V, _V,
                                            C.c_int32, C.POINTER(C.c_int32), _V]),
    "mpcv_ltv_lateral": (C.c_int, [_V, C.c_int32, _V, C.c_double, C.c_double, C.c_double, C.c_int32, _V, C.c_int64, _V]),
    "mpcv_ltv_dynbike": (C.c_int, [_V, C.c_int32, C.c_int32, _V, C.c_double, C.c_int32, _V, C.c_int64, _V]),
}


class LoopArgs(C.Structure):
    """ctypes mirror of `mpcv_loop_args` (include/mpcv.h)."""
    _fields_ = [("x_init", _V), ("pglob", _V), ("pglob_traj", _V), ("ptraj", _V), ("lbx", _V), ("ubx", _V),
                ("n_steps", C.c_int32), ("warm_mode", C.c_int32), ("flags", C.c_int32), ("reserved_", C.c_int32),
                ("stop_radius", C.c_double), ("out_states", _V), ("out_controls", _V), ("out_steps", _V),
                ("out_iters", _V), ("out_status", _V), ("out_horizons", _V), ("out_step_ns", _V)]


class MpcvError(RuntimeError):
    pass


def lib():
    """Load libmpcv.so; raise (never fall back) when it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MpcvError(
                "mpc_verde_b200: CUDA library %s not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C mpc_verde_b200/csrc`); there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the ABI and the header ever diverge
            fn.restype = res
            fn.argtypes = args
        _LIB = l
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().mpcv_last_error()
        raise MpcvError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
