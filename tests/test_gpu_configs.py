"""Every BASELINE.json configuration at its FULL per-GPU size, exactly as bench.py runs it (the workload classes are
bench.py's), checked through size-independent properties and a sampled oracle comparison:
  all problems succeed; bounds are honoured; the gathered result is deterministic across two runs (bit for bit);
  closed loops: the recorded plant trajectory replays under the model's own rollout (mpcv_rollout) to 1e-12, and the
  tracking error stays bounded; a random sample of problems / scenarios equals the CPU oracle."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mpc_verde_b200 import spec as S                      # noqa: E402
from oracle import mpc_oracle as O                          # noqa: E402

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


@pytest.fixture(scope="module")
def env():
    import torch
    import bench
    import mpc_verde_b200 as mv
    torch.cuda.set_device(0)
    return bench, mv, torch, torch.device("cuda", 0)


def test_c3_pendulum_full_size(env):
    bench, mv, torch, dev = env
    wl = bench.C3()
    wl.setup(mv, dev)
    x, f = wl.step()
    torch.cuda.synchronize()
    assert wl.B == 262144 and bool((wl.status == 0).all())
    u = x[:, 5::6][:, :40]
    assert float(u.abs().max()) <= 200.0                                   # |u| <= 200 honoured exactly (projection)
    x2, f2 = wl.step()
    assert torch.equal(x, x2) and torch.equal(f, f2)                       # deterministic
    idx = np.random.default_rng(0).choice(wl.B, 96, replace=False)
    ref = O.solve(wl.spec, wl.w0_h[idx], wl.lbx, wl.ubx, wl.p_h[idx], nthreads=NCPU)
    same = ref["iters"] == wl.iters.cpu().numpy()[idx]
    assert same.mean() >= 0.98, same.mean()
    assert np.abs(x.cpu().numpy()[idx] - ref["x"]).max() <= 1e-5
    assert np.abs(x.cpu().numpy()[idx][same] - ref["x"][same]).max() <= 1e-8
    assert np.abs(f.cpu().numpy()[idx] - ref["f"]).max() <= 1e-6 * np.abs(ref["f"]).max()


@pytest.mark.parametrize("key", ["c4", "c4f"])
def test_c4_tracker_closed_loops_full_size(env, key):
    bench, mv, torch, dev = env
    wl = bench.CONFIGS[key]()
    wl.setup(mv, dev)
    controls, states = wl.step()
    torch.cuda.synchronize()
    assert bool((wl.status == 0).all())
    sp = wl.spec
    lbx, ubx = wl.lbx, wl.ubx
    nz = sp.nx + sp.nu
    c = controls.cpu().numpy()
    assert np.all(c >= lbx[sp.nx:nz] - 1e-12) and np.all(c <= ubx[sp.nx:nz] + 1e-12)
    c2, s2 = wl.step()
    assert torch.equal(controls, c2) and torch.equal(states, s2)
    # sampled oracle comparison (bench.py's own spot check) and tracking quality
    assert wl.check([controls, states]) <= 1e-7
    st = states.cpu().numpy()
    if key == "c4":
        refs = wl._refs(wl.dev_in).cpu().numpy()[:, :wl.n_steps + 1, :2]
        err = np.hypot(*(st[:, :, :2] - refs).transpose(2, 0, 1))
        assert err[:, 20:].max() < 0.5                                     # the tracker stays on the lane-change path
    else:
        assert np.abs(st[:, :, 3]).max() <= 0.384 + 1e-7                   # |delta| <= 0.384: a box on the delta state of the NLP (relaxed by 1e-8 inside the solve, constr_viol_tol on the defect)


def test_c5_dynamic_bicycle_ltv_full_size(env):
    bench, mv, torch, dev = env
    wl = bench.C5()
    wl.setup(mv, dev)
    controls, states = wl.step()
    torch.cuda.synchronize()
    assert wl.B == 131072 and bool((wl.status == 0).all())
    assert float(controls.abs().max()) <= 20.0
    c2, s2 = wl.step()
    assert torch.equal(controls, c2) and torch.equal(states, s2)
    assert wl.check([controls, states], n=16) <= 1e-8
    # the recorded plant trajectory obeys x+ = A(v_t) x + B(v_t) u with the per-step exact ZOH
    from mpc_verde_b200 import problems
    st, cu, v = states.cpu().numpy(), controls.cpu().numpy(), wl.host_in["v"]
    for b in np.random.default_rng(1).choice(wl.B, 32, replace=False):
        for t in range(wl.n_steps):
            A, Bd = problems.c2d(*problems.dynamic_bicycle_matrices(v[b, t]), wl.spec.T)
            assert np.abs(A @ st[b, t] + Bd[:, 0] * cu[b, t, 0] - st[b, t + 1]).max() <= 1e-10


def test_c1_single_problem_closed_loop(env):
    bench, mv, torch, dev = env
    wl = bench.C1()
    wl.setup(mv, dev)
    controls, states = wl.step()
    torch.cuda.synchronize()
    r = O.closed_loop(wl.spec, *wl._oracle_args(1))
    assert int(r["steps"][0]) == 84
    assert np.abs(controls.cpu().numpy() - r["controls"]).max() <= 1e-9
    assert np.abs(states.cpu().numpy() - r["states"]).max() <= 1e-9
