"""mpc_verde_b200 — B200-native batched nonlinear-MPC solver (drop-in for the reference's
`casadi.nlpsol('ipopt')` shooting solve and closed loop)."""
from . import problems, spec  # noqa: F401
from .spec import (WARM_COLD, WARM_REFERENCE, WARM_SHIFT, LAYOUT_AUTO, LAYOUT_THREAD, LAYOUT_WARP, LAYOUT_PHASED, LAYOUT_RESIDENT,  # noqa: F401
                   STATUS_NAMES, Spec)


def __getattr__(name):
    # solver.py imports torch and binds the CUDA library; keep `import mpc_verde_b200.spec` light
    import importlib
    if name in ("nlpsol", "NlpSolver", "fp64_peak", "c2d"):
        return getattr(importlib.import_module(__name__ + ".solver"), name)
    if name in ("dist", "solver", "mpctools", "reference", "sinks"):
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
