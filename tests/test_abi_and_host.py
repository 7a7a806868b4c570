"""CPU-side checks: the C-ABI library loads and exports every symbol include/mpcv.h declares
(no compute calls without a GPU), host-side sizes/layouts, and the product has no CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

from mpc_verde_b200 import _lib, problems
from mpc_verde_b200 import spec as S
from oracle import mpc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpcv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _declared_symbols()
    assert len(names) >= 12
    l = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(l, n), n
    assert sorted(_lib.SIGNATURES) == names          # the Python binding covers exactly the header


def test_dims_agree_between_spec_header_library_and_oracle():
    l = _lib.lib()
    for mk in (S.unicycle_multiple_shooting, S.unicycle_single_shooting_rk4, S.unicycle_single_shooting_euler,
               lambda: S.unicycle_tracking(N=20), lambda: S.linear_tracking(3, 20, (1, 1, 1, 0), 1.0),
               lambda: S.linear_tracking(4, 50, (1, 1, 1, 1), 1.0), lambda: S.linear_tracking(4, 40, (1, 0, 1, 0), 0.0, R1=1e-4),
               lambda: S.linear_tracking(3, 5, (1, 1, 0, 0), 0.01, R1=0.1)):
        sp = mk()
        v = [ctypes.c_int32() for _ in range(7)]
        assert l.mpcv_dims(ctypes.byref(sp), *[ctypes.byref(x) for x in v]) == 0
        got = tuple(x.value for x in v)
        assert got == (sp.nx, sp.nu, sp.n_var, sp.n_g, sp.n_p, sp.npg, sp.nps)
        w = [ctypes.c_int32() for _ in range(7)]
        assert O.lib().mpco_dims(ctypes.byref(sp), *[ctypes.byref(x) for x in w]) == 0
        assert tuple(x.value for x in w) == got


def test_spec_defaults_are_the_scripts_problem():
    l = _lib.lib()
    sp = S.Spec()
    l.mpcv_spec_defaults(ctypes.byref(sp))
    ref = S.unicycle_multiple_shooting()
    for f in ("model", "shooting", "N", "M", "T", "tol", "mu_init", "bound_push", "bound_frac"):
        assert getattr(sp, f) == getattr(ref, f), f
    assert list(sp.Q)[:3] == [1.0, 5.0, 0.1] and list(sp.R) == [0.5, 0.05]


def test_no_cpu_path_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mpc_verde_b200 as mv
    with pytest.raises(_lib.MpcvError):
        mv.nlpsol("solver", "ipopt", problems.unicycle_multiple_shooting())
    sp = S.unicycle_multiple_shooting()
    l = _lib.lib()
    assert not l.mpcv_create(ctypes.byref(sp))
    assert b"no CPU path" in l.mpcv_last_error()
    t = ctypes.c_double()
    assert l.mpcv_fp64_peak(ctypes.byref(t), None, None) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc_verde_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                # nothing shipped may import, load, link or include the checker or the CPU harness
                assert "libmpc_oracle" not in src and "libhostsim" not in src, fn
                for line in src.splitlines():
                    code = line.split("#")[0] if fn.endswith(".py") else line.split("//")[0]
                    if re.search(r"\b(import|from|include|CDLL)\b", code):
                        assert not re.search(r"oracle|hostsim", code), (fn, line)


def test_bounds_and_guess_helpers_follow_the_script_layout():
    sp = S.unicycle_multiple_shooting()
    lb, ub = problems.unicycle_bounds(sp)
    assert lb.shape == (53,) and np.isinf(lb[:3]).all() and lb[3] == -1 and ub[4] == np.pi / 4 and np.isinf(ub[50:]).all()
    w0 = problems.cold_start(sp, [1.0, 2.0, 3.0])
    assert w0.shape == (1, 53) and list(w0[0, 5:8]) == [1, 2, 3] and w0[0, 3] == 0 and list(w0[0, 50:]) == [1, 2, 3]
    ss = S.unicycle_single_shooting_euler()
    lb, ub = problems.unicycle_bounds(ss)
    assert lb.shape == (20,) and lb[0] == -1 and lb[1] == -np.pi / 4
    A, B = problems.rk4_linear(problems.PENDULUM_AC, problems.PENDULUM_BC, 0.01)
    Ae, Be = problems.c2d(problems.PENDULUM_AC, problems.PENDULUM_BC, 0.01)
    assert np.abs(A - Ae).max() < 1e-5 and np.abs(B - Be).max() < 1e-5   # RK4 is the 4th-order Taylor of expm
    Ac, Bc = problems.dynamic_bicycle_matrices(0.8)
    ev = np.sort(np.linalg.eigvals(Ac).real)
    assert abs(ev[0] + 641.0) < 1.0 and abs(ev[1] + 224.7) < 1.0      # stiff: why c2d, not RK4
