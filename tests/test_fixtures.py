"""Self-generated fixtures (tests/golden/self_generated/, reference-unpinned: see make_fixtures.py).

CPU: the oracle still reproduces them bit for bit (guards the checker itself).
GPU: the CUDA kernels reproduce them through the Python mirror / C ABI -- `vmap_step` and `vmap_sim_ahead` in both
memory layouts -- within the north star's tolerances (fp32 1e-6 per step and 1e-4 over a trajectory, fp64 1e-12 / 1e-9,
relative to the field's normalisation half-span; flags bit exact)."""
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import field_scales, scaled_err, traj_err
from oracle import oracle as O

FIX = np.load(Path(__file__).resolve().parent / "golden" / "self_generated" / "fixtures.npz")
sys.path.insert(0, str(Path(__file__).resolve().parent / "golden" / "self_generated"))
from make_fixtures import ENVS, make_oracle  # noqa: E402

CASES = [(e, s, d) for e in ENVS for s in ("euler", "tsit5") for d in ("f32", "f64")]
STEP_TOL = {"f32": 1e-6, "f64": 1e-12}
TRAJ_TOL = {"f32": 1e-4, "f64": 1e-9}


def _inputs(key, n_state):
    return [FIX[f"{key}/state{i}"] for i in range(n_state)], FIX[f"{key}/actions"]


def _orc(env, sname, dname, **kw):
    return make_oracle(env, O.TSIT5 if sname == "tsit5" else O.EULER, np.float32 if dname == "f32" else np.float64, **kw)


@pytest.mark.parametrize("env,sname,dname", CASES)
def test_oracle_reproduces_self_generated_fixtures(env, sname, dname):
    orc = _orc(env, sname, dname)
    key = f"{env}/{sname}/{dname}"
    st, acts = _inputs(key, orc.D)
    r = orc.step(st, acts[:, 0], epilogue=True)
    assert np.array_equal(r["obs"], FIX[f"{key}/step/obs"], equal_nan=True)
    s = orc.sim_ahead(st, acts, orc.tau, orc.tau, epilogue=True)
    assert np.array_equal(s["obs"], FIX[f"{key}/sim/obs"], equal_nan=True)
    for i, x in enumerate(s["traj"]):
        assert np.array_equal(x, FIX[f"{key}/sim/traj{i}"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("env,sname,dname", CASES)
def test_kernels_reproduce_self_generated_fixtures(excenvs, env, sname, dname):
    import torch

    from gpu_helpers import make_env, np_, to_state

    orc = _orc(env, sname, dname)
    key = f"{env}/{sname}/{dname}"
    st, acts = _inputs(key, orc.D)
    B = st[0].shape[0]
    e = make_env(excenvs, orc, B)
    sc = field_scales(orc)
    # vmap_step
    obs, new, rew, trunc, term = e.vmap_step_with_epilogue(to_state(e, st), torch.as_tensor(acts[:, 0], device="cuda"))
    for i, (name, s_) in enumerate(zip(e.PHYS_FIELDS, sc)):
        assert scaled_err(np_(getattr(new.physical_state, name)), FIX[f"{key}/step/state{i}"], s_) <= STEP_TOL[dname], name
    assert scaled_err(np_(obs), FIX[f"{key}/step/obs"], 1.0) <= STEP_TOL[dname]
    assert np.array_equal(np_(trunc).astype(np.uint8), FIX[f"{key}/step/truncated"])
    # vmap_sim_ahead, both layouts
    for layout in ("batch_major", "time_major"):
        o, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(
            to_state(e, st), torch.as_tensor(acts, device="cuda"), orc.tau, orc.tau, layout=layout)
        for i, (name, s_) in enumerate(zip(e.PHYS_FIELDS, sc)):
            err = traj_err(np_(getattr(states.physical_state, name)), FIX[f"{key}/sim/traj{i}"], s_)
            assert err <= TRAJ_TOL[dname], (layout, name, err)
        ref_obs = FIX[f"{key}/sim/obs"]
        assert traj_err(np_(o).reshape(B, -1), ref_obs.reshape(B, -1), 1.0) <= TRAJ_TOL[dname], layout
        tr, tr_ref = np_(trunc).astype(np.uint8), FIX[f"{key}/sim/truncated"]
        if orc.n_trunc == orc.O:
            safe = np.abs(np.abs(ref_obs) - 1.0) > 1e-3
            assert np.array_equal(tr[safe], tr_ref[safe]), layout
        else:
            assert np.array_equal(tr, tr_ref), layout


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
@pytest.mark.parametrize("dname", ["f32", "f64"])
@pytest.mark.parametrize("name", ["half", "third", "coarse"])
def test_kernels_reproduce_zoh_quirk_fixtures(excenvs, name, dname, layout):
    import torch

    from gpu_helpers import make_env, np_, to_state

    key = f"quirk_{name}/{dname}"
    obs_dt, act_dt, n_steps = FIX[f"{key}/steps"]
    orc = O.Oracle("mass_spring_damper", solver=O.TSIT5, dtype=np.float32 if dname == "f32" else np.float64, tau=obs_dt)
    st, acts = [FIX[f"{key}/state0"], FIX[f"{key}/state1"]], FIX[f"{key}/actions"]
    e = make_env(excenvs, orc, st[0].shape[0])
    o, _, _ = e.vmap_sim_ahead(to_state(e, st), torch.as_tensor(acts, device="cuda"), float(obs_dt), float(act_dt),
                               layout=layout)
    ref = FIX[f"{key}/obs"]
    assert o.shape == ref.shape
    assert e.sim_ahead_shape(acts.shape[1], float(obs_dt), float(act_dt)) == (ref.shape[1], int(n_steps))
    assert np.max(np.abs(np_(o) - ref)) <= TRAJ_TOL[dname]


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
def test_kernels_reproduce_tracking_fixture(excenvs, layout):
    import torch

    from gpu_helpers import make_env, np_, to_state

    key = "tracking_cart_pole/f32"
    base = O.Oracle("cart_pole", dtype=np.float32)
    static = {k: FIX[f"{key}/static/{k}"] for k in base.static}
    orc = O.Oracle("cart_pole", solver=O.TSIT5, dtype=np.float32, static_params=static, tau=1e-3,
                   control_state=["theta", "velocity"])
    st, refl = [FIX[f"{key}/state{i}"] for i in range(4)], [FIX[f"{key}/ref{i}"] for i in range(4)]
    e = make_env(excenvs, orc, st[0].shape[0])
    o, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(
        to_state(e, st, refl), torch.as_tensor(FIX[f"{key}/actions"], device="cuda"), 1e-3, 1e-3, layout=layout)
    assert np.max(np.abs(np_(o) - FIX[f"{key}/obs"])) <= 1e-4
    assert np.max(np.abs(np_(rew)[..., 0] - FIX[f"{key}/reward"])) <= 1e-4
    assert np.array_equal(np_(term)[..., 0].astype(np.uint8), FIX[f"{key}/terminated"])
