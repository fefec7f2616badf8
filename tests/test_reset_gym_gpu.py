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
