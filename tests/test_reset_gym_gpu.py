"""`-m gpu`: vmap_reset state packing, observation <-> state round trip, GymWrapper, and the reference's own
shape/structure tests (reference tests/envs/test_core_functions.py, tests/test_gym_wrapper.py) on the CUDA path."""
import numpy as np
import pytest
import torch

from gpu_helpers import np_
from oracle import oracle as O

pytestmark = pytest.mark.gpu

NAME = {"CART_POLE": "cart_pole", "MASS_SPRING_DAMPER": "mass_spring_damper", "PENDULUM": "pendulum",
        "FLUID_TANK": "fluid_tank", "PMSM": "pmsm", "ACROBOT": "acrobot"}


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_default_reset_matches_oracle_and_round_trips(excenvs, dtype):
    """reference test_core_functions.py:55-77: obs == generate_observation(generate_state_from_observation(obs))
    exactly, for single and vmapped calls."""
    for reg in excenvs.EnvironmentRegistry:
        e = reg.make(batch_size=4, dtype=dtype)
        orc = O.Oracle(NAME[reg.name], dtype=np.float64 if dtype == torch.float64 else np.float32)
        ref = orc.reset(4)
        obs, state = e.vmap_reset()
        assert np.array_equal(np_(obs), ref["obs"]), reg
        for n, b in zip(e.PHYS_FIELDS, ref["state"]):
            assert np.array_equal(np_(getattr(state.physical_state, n)), b), (reg, n)
        assert torch.equal(obs, e.vmap_generate_observation(state))
        assert torch.equal(obs, e.vmap_generate_observation(e.vmap_generate_state_from_observation(obs)))
        # single-environment API
        o1, s1 = e.reset(e.env_properties)
        assert o1.shape == e.obs_description.shape
        assert torch.equal(o1, e.generate_observation(s1, e.env_properties))
        assert torch.equal(o1, e.generate_observation(e.generate_state_from_observation(o1, e.env_properties),
                                                      e.env_properties))
        assert bool(torch.isnan(state.PRNGKey).all()) and not bool(state.additions.active_solver_state.any())
        assert bool(torch.isnan(state.reference.__dict__[e.PHYS_FIELDS[0]]).all())


def test_random_reset_distribution(excenvs):
    for reg in excenvs.EnvironmentRegistry:
        e = reg.make(batch_size=4096)
        obs, state = e.vmap_reset(rng=1234)
        assert obs.shape == (4096, len(e.obs_description))
        lim = 1.0 + 1e-5
        if reg.name != "PMSM":
            assert bool((obs.abs() <= lim).all())
            assert obs.std().item() > 0.2
        else:
            assert bool((obs[:, :2].abs() <= lim).all())  # currents inside their box
        assert type(state) == e.State


@pytest.mark.parametrize("reg_name", list(NAME))
def test_step_and_sim_ahead_shapes_and_structure(excenvs, reg_name):
    """reference test_core_functions.py:80-131 (test_step, test_simulate_ahead)."""
    ex = excenvs
    e = ex.EnvironmentRegistry[reg_name].make(batch_size=4)
    init_obs, state = e.reset(e.env_properties)
    struct = ex.tree_structure(state)
    for _ in range(20):
        obs, state = e.step(state, torch.ones(e.action_dim), e.env_properties)
    assert init_obs.shape == obs.shape and struct == ex.tree_structure(state)
    init_obs, state = e.vmap_reset()
    struct = ex.tree_structure(state)
    for _ in range(20):
        obs, state = e.vmap_step(state, torch.ones((4, e.action_dim), device="cuda"))
    assert init_obs.shape == obs.shape and struct == ex.tree_structure(state)
    obs, init_state = e.reset(e.env_properties)
    obs, states, last_state = e.sim_ahead(init_state, torch.ones((10, e.action_dim)), e.env_properties, e.tau, e.tau)
    assert obs.shape == (11, len(e.obs_description))
    assert ex.tree_structure(init_state) == ex.tree_structure(last_state)
    obs, init_state = e.vmap_reset()
    obs, states, last_state = e.vmap_sim_ahead(init_state, torch.ones((4, 10, e.action_dim), device="cuda"), e.tau, e.tau)
    assert obs.shape == (4, 11, len(e.obs_description))
    assert ex.tree_structure(init_state) == ex.tree_structure(last_state)


@pytest.mark.parametrize("reg_name", list(NAME))
def test_gym_wrapper(excenvs, reg_name):
    """reference tests/test_gym_wrapper.py:19-48: gym obs bit-identical to vmap_step obs; output shapes; reference
    generation switches on."""
    ex = excenvs
    e = ex.EnvironmentRegistry[reg_name].make(batch_size=4)
    gym = ex.GymWrapper(env=e)
    action = torch.ones((4, e.action_dim), device="cuda")
    _, state = e.vmap_reset()
    new_obs, state = e.vmap_step(state, action)
    gym.reset()
    new_obs_gym, reward, terminated, truncated = gym.step(action)
    assert torch.equal(new_obs, new_obs_gym)
    assert reward.shape == (4, 1) and terminated.shape == (4, 1)
    gym2 = ex.GymWrapper(env=ex.EnvironmentRegistry[reg_name].make(batch_size=4),
                         control_state=[e.PHYS_FIELDS[-1] if reg_name != "PMSM" else "i_d"])
    gym2.reset(rng_env=0, rng_ref=1)
    assert gym2.ref_gen is True and gym2.reference_hold_steps.shape == (4, 1)
    for _ in range(3):
        o, r, te, tr = gym2.step(action)
    assert o.shape == (4, len(gym2.env.obs_description)) and torch.isfinite(r).all()


@pytest.mark.parametrize("reg_name,field", [("PENDULUM", "theta"), ("MASS_SPRING_DAMPER", "velocity"),
                                            ("FLUID_TANK", "height"), ("PMSM", "i_q")])
def test_gym_wrapper_reference_generator_order_and_distribution(excenvs, reg_name, field):
    """Reference gym_wrapper.py:109-126 runs update_ref AFTER vmap_step and BEFORE reward / terminated / truncated: on a
    step where a hold counter expires the reward must be computed against the NEW reference while the returned
    observation still shows the old one.  hold_steps in [1, 2): every counter expires on every step from the second on.  The
    new references come from the device-side counter RNG: every environment draws its own, uniformly over the field's
    range (init_state's distribution), and the draws differ from step to step."""
    ex = excenvs
    B = 8192
    env = ex.EnvironmentRegistry[reg_name].make(batch_size=B)
    gym = ex.GymWrapper(env=env, control_state=[field], ref_params={"hold_steps_min": 1, "hold_steps_max": 2})
    gym.reset(rng_env=0, rng_ref=1)
    assert gym.ref_gen and bool((gym.reference_hold_steps == 1).all())
    fi = env.PHYS_FIELDS.index(field)
    lo, hi = (float(getattr(env.env_properties.physical_normalizations, field).min),
              float(getattr(env.env_properties.physical_normalizations, field).max))
    action = torch.zeros((B, env.action_dim), device="cuda")
    seen = []
    for step in range(4):
        st_before = ex.tree_unflatten(gym.state_tree_struct, gym.state)
        ref_before = getattr(st_before.reference, field).clone()
        hold_before = gym.reference_hold_steps.clone()
        obs, reward, terminated, truncated = gym.step(action)
        st_after = ex.tree_unflatten(gym.state_tree_struct, gym.state)
        ref_after = getattr(st_after.reference, field)
        expired = hold_before[:, 0] == 0
        # counters: expired ones were redrawn in [1, 2) and decremented -> 0; the others were decremented
        assert bool((gym.reference_hold_steps[:, 0] == torch.where(expired, 0, hold_before[:, 0] - 1)).all())
        assert bool((ref_after[~expired] == ref_before[~expired]).all())
        if expired.any():
            assert bool((ref_after[expired] != ref_before[expired]).float().mean() > 0.999)
            seen.append(ref_after.clone())
        # the observation of the step carries the OLD reference ...
        pn = getattr(env.env_properties.physical_normalizations, field)
        assert torch.allclose(obs[:, -1], pn.normalize(ref_before).to(obs.dtype), rtol=0, atol=1e-6)
        # ... reward / terminated / truncated the NEW one
        r2, tr2, te2 = env.vmap_generate_reward_truncated_terminated(st_after)
        assert torch.equal(reward, r2) and torch.equal(truncated, tr2) and torch.equal(terminated, te2)
    assert len(seen) == 3 and bool((seen[0] != seen[1]).float().mean() > 0.999)  # fresh draws every time
    x = seen[-1].double()
    if reg_name == "PMSM":  # currents: disc of radius i_max reflected into the box (pmsm_env.py:404-427)
        assert x.min() >= lo and x.max() <= hi
    else:  # uniform over the range (fluid tank: uniform(0, 1) of the normalised height = the upper half of the range)
        a, b = ((lo + hi) / 2, hi) if reg_name == "FLUID_TANK" else (lo, hi)
        assert x.min() >= a and x.max() <= b
        assert abs(x.mean().item() - (a + b) / 2) < 0.02 * (b - a)
        assert abs(x.std().item() - (b - a) / 12 ** 0.5) < 0.02 * (b - a)
    # neighbouring environments draw independently
    c = torch.corrcoef(torch.stack([x[:-1], x[1:]]))[0, 1].abs().item()
    assert c < 0.05


def test_single_environment_epilogue_hooks(excenvs):
    """generate_reward / generate_truncated / generate_terminated of ONE environment (pendulum_env.py:297-309, 381-390),
    as GymWrapper reads them off the env object (gym_wrapper.py:53-58)."""
    ex = excenvs
    env = ex.Pendulum(batch_size=4, control_state=["theta"])
    obs, st = env.reset(env.env_properties)
    st.reference.theta = torch.tensor(0.5, device="cuda")
    r = env.generate_reward(st, torch.zeros(1), env.env_properties)
    tr = env.generate_truncated(st, env.env_properties)
    te = env.generate_terminated(st, r, env.env_properties)
    assert tuple(r.shape) == (1,) and tuple(tr.shape) == (3,) and tuple(te.shape) == (1,)
    import math
    want = -((math.sin(math.pi) - math.sin(0.5)) ** 2 + (math.cos(math.pi) - math.cos(0.5)) ** 2)
    assert abs(r.item() - want) < 1e-5 and not bool(te.item()) and not bool(tr.any())
    tank = ex.FluidTank(batch_size=2)
    _, ts = tank.reset(tank.env_properties)
    assert tuple(tank.generate_truncated(ts, tank.env_properties).shape) == (1,)
    assert not bool(tank.generate_terminated(ts, None, tank.env_properties).item())
