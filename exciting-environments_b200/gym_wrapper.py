"""Gym-style stateful adapter with the reference's surface (reference gym_wrapper.py:14-207): holds the flattened
state, ``step(action) -> (obs, reward, terminated, truncated)``, ``reset(rng_env, rng_ref, initial_state)`` and an
optional random reference generator with hold steps.

Where the reference issues ``env.vmap_step`` followed by three more vmapped passes for reward / terminated /
truncated inside one jit (gym_wrapper.py:109-126), this wrapper makes ONE kernel launch when the reference generator
is off: the epilogue is fused into the step kernel (``CoreEnvironment.vmap_step_with_epilogue``).  With the generator
on, the step kernel is followed by ``exenv_gym_epilogue``: per-environment hold counters, new references from a
counter-based device RNG (Philox4x32-10, one stream per environment) where a counter ran out, and the epilogue against
the UPDATED reference -- the reference's order of operations.  JAX's threefry bit stream is not reproducible without
JAX: the generator has the reference's distributions, not its bit pattern.
"""
from __future__ import annotations

import torch

from .registration import EnvironmentRegistry
from .utils import tree_flatten, tree_structure, tree_unflatten


class GymWrapper:
    def __init__(self, env, control_state=None, generate_reward=None, generate_terminated=None,
                 generate_truncated=None, ref_params=None):
        self.env = env
        if control_state is None:
            print(f"No chosen control state in the GymWrapper. Control state is set to {self.env.control_state}.")
            self.control_state = self.env.control_state
        else:
            assert type(control_state) == list, f"Control state has to be a list."
            for i in control_state:
                assert i in list(
                    self.env.PHYS_FIELDS
                ), f"Given control state {i} is no valid physical state {list(self.env.PHYS_FIELDS)}."
            self.control_state = control_state
            self.env.control_state = control_state
        self.ref_gen = False
        _, init_state = self.env.vmap_reset()
        self.ref_params = ref_params or {"hold_steps_min": 10, "hold_steps_max": 1000}
        self.reference_hold_steps = torch.zeros((self.env.batch_size, 1), dtype=torch.int32, device=self.env.device)
        self.state, self.state_tree_struct = tree_flatten(init_state)
        # user-supplied epilogue callables replace the fused one (they receive the batched state)
        self.generate_reward = generate_reward
        self.generate_truncated = generate_truncated
        self.generate_terminated = generate_terminated

    @classmethod
    def from_env(cls, env_type: EnvironmentRegistry, **env_kwargs):
        """Creates GymWrapper with environment from EnvironmentRegistry."""
        return cls(env_type.make(**env_kwargs))

    def step(self, action):
        """One simulation step for all batches.

        Returns ``(observation (B, O'), reward (B, 1), terminated (B, 1), truncated (B, n_trunc))``."""
        obs, reward, terminated, truncated, self.state, self.reference_hold_steps = self.gym_step(
            action, self.state, self.reference_hold_steps
        )
        return obs, reward, terminated, truncated

    def gym_step(self, action, state, reference_hold_steps):
        """``vmap_step``, then ``update_ref``, then reward / terminated / truncated of the new state against the
        UPDATED reference -- the reference's order (gym_wrapper.py:109-126); the observation returned by the step
        still carries the previous reference, as upstream.  Without the reference generator the epilogue is fused
        into the step kernel (one launch); with it, the step kernel is followed by one small kernel
        (``exenv_gym_epilogue``: hold counters, counter-based random references, epilogue)."""
        state = tree_unflatten(self.state_tree_struct, state)
        env = self.env
        if len(self.control_state) and self.ref_gen:
            obs, state = env.vmap_step(state, action)
            state, reference_hold_steps, reward, truncated, terminated = self._ref_epilogue(
                state, reference_hold_steps, force_new=False)
        else:
            obs, state, reward, truncated, terminated = env.vmap_step_with_epilogue(state, action)
        if self.generate_reward is not None:
            reward = self.generate_reward(state, action, env.env_properties)
        if self.generate_terminated is not None:
            terminated = self.generate_terminated(state, reward, env.env_properties)
        if self.generate_truncated is not None:
            truncated = self.generate_truncated(state, env.env_properties)
        return obs, reward, terminated, truncated, tree_flatten(state)[0], reference_hold_steps

    def _ref_epilogue(self, state, hold_steps, force_new):
        """One launch of ``exenv_gym_epilogue`` (reference gym_wrapper.py:111-126 / 170-192).  Functional like the
        reference: the tracked reference leaves, the hold counters and the per-environment RNG state are new tensors."""
        import ctypes as C

        from . import _lib

        env = self.env
        B, dev, dt = env.batch_size, env.device, env.dtype
        phys = env._phys_leaves(state.physical_state, B)
        refs = []
        for n in env.PHYS_FIELDS:
            r = env._as(getattr(state.reference, n), (B,))
            refs.append(r.clone() if n in self.control_state else r)
        hold = hold_steps.reshape(B).to(torch.int32).clone()
        rng = state.PRNGKey.clone()
        assert rng.dtype == torch.int64 and tuple(rng.shape) == (B, 2), "reference generator state: (B, 2) int64"
        a = _lib.ExenvGymArgs()
        a.batch = B
        for i in range(env.physical_state_dim):
            a.state[i], a.ref[i] = phys[i].data_ptr(), refs[i].data_ptr()
        a.hold_steps, a.rng = hold.data_ptr(), rng.data_ptr()
        a.hold_steps_min, a.hold_steps_max = int(self.ref_params["hold_steps_min"]), int(self.ref_params["hold_steps_max"])
        a.force_new = 1 if force_new else 0
        rew = torch.empty((B, 1), dtype=dt, device=dev)
        trunc = torch.empty((B, env.n_trunc), dtype=torch.bool, device=dev)
        term = torch.empty((B, 1), dtype=torch.bool, device=dev)
        a.reward, a.truncated, a.terminated = rew.data_ptr(), trunc.data_ptr(), term.data_ptr()
        _lib.check(_lib.lib().exenv_gym_epilogue(env._plan().handle, C.byref(a), env._stream()))
        new_state = env.State(physical_state=state.physical_state, PRNGKey=rng, additions=state.additions,
                              reference=env.PhysicalState(*refs))
        return new_state, hold.reshape(B, 1), rew, trunc, term

    def reset(self, rng_env=None, rng_ref=None, initial_state=None):
        """Resets the environment to a random or passed initial state; ``rng_ref`` switches the reference generator
        on (reference gym_wrapper.py:132-168)."""
        if initial_state is not None:
            obs, state = self.env.vmap_reset(initial_state=tree_unflatten(self.state_tree_struct, initial_state))
        else:
            _, state = self.env.vmap_reset(rng_env)
        B, dev = self.env.batch_size, self.env.device
        if rng_ref is not None:
            # per-environment generator state: (key, draw counter) -- the analogue of jax.random.split(rng_ref, B)
            state.PRNGKey = _ref_keys(rng_ref, B, dev)
            self.ref_gen = True
            if len(self.control_state):
                state, self.reference_hold_steps, _, _, _ = self._ref_epilogue(
                    state, torch.zeros((B, 1), dtype=torch.int32, device=dev), force_new=True)
            else:
                self.reference_hold_steps = torch.zeros((B, 1), dtype=torch.int32, device=dev)
        else:
            self.ref_gen = False
            print("Since no PRNGKey for reference was provided, reference generation is deactivated.")
        self.state, self.state_tree_struct = tree_flatten(state)
        obs = self.env.vmap_generate_observation(state)
        return obs, {}

    def update_ref(self, state, hold_steps):
        """New references where the hold counter ran out, then count down (reference gym_wrapper.py:170-175)."""
        state, hold_steps, _, _, _ = self._ref_epilogue(state, hold_steps, force_new=False)
        return state, hold_steps

    def generate_new_ref(self, state, hold_steps=None):
        """Random reference for the tracked fields of EVERY batch and a new random hold time (reference
        gym_wrapper.py:177-192 vmapped over the batch)."""
        B = self.env.batch_size
        if hold_steps is None:
            hold_steps = torch.zeros((B, 1), dtype=torch.int32, device=self.env.device)
        state, hold_steps, _, _, _ = self._ref_epilogue(state, hold_steps, force_new=True)
        return state, hold_steps

    def render(self, *_, **__):
        raise NotImplementedError("To be implemented!")

    def close(self):
        raise NotImplementedError("To be implemented!")


def _ref_keys(rng_ref, B, device):
    """(B, 2) int64 generator state of the device-side reference generator: column 0 = per-environment Philox key
    (SplitMix64 of seed + index, so neighbouring environments get unrelated keys), column 1 = draw counter (0)."""
    if isinstance(rng_ref, torch.Tensor) and rng_ref.dtype == torch.int64 and tuple(rng_ref.shape) == (B, 2):
        return rng_ref.to(device).clone()
    if isinstance(rng_ref, torch.Generator):
        seed = rng_ref.initial_seed()
    elif isinstance(rng_ref, torch.Tensor):
        seed = int(rng_ref.reshape(-1)[-1].item()) if rng_ref.numel() else 0
    else:
        seed = int(rng_ref)
    import numpy as np

    with np.errstate(over="ignore"):
        z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.arange(1, B + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    keys = torch.from_numpy(z.view(np.int64).copy())
    return torch.stack([keys, torch.zeros(B, dtype=torch.int64)], dim=1).to(device)
