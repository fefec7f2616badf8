"""Gym-style stateful adapter with the reference's surface (reference gym_wrapper.py:14-207): holds the flattened
state, ``step(action) -> (obs, reward, terminated, truncated)``, ``reset(rng_env, rng_ref, initial_state)`` and an
optional random reference generator with hold steps.

Where the reference issues ``env.vmap_step`` followed by three more vmapped passes for reward / terminated /
truncated inside one jit (gym_wrapper.py:109-126), this wrapper makes ONE kernel launch: the epilogue is fused into
the step kernel (``CoreEnvironment.vmap_step_with_epilogue``).  The reference generator draws from a
``torch.Generator`` (JAX's threefry stream is not reproducible without JAX): statistical, not bit-wise, parity.
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
        self.reference_hold_steps = torch.zeros((self.env.batch_size, 1), device=self.env.device)
        self.state, self.state_tree_struct = tree_flatten(init_state)
        # user-supplied epilogue callables replace the fused one (they receive the batched state)
        self.generate_reward = generate_reward
        self.generate_truncated = generate_truncated
        self.generate_terminated = generate_terminated
        self._ref_rng = None

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
        state = tree_unflatten(self.state_tree_struct, state)
        obs, state, reward, truncated, terminated = self.env.vmap_step_with_epilogue(state, action)
        if len(self.control_state) and self.ref_gen:
            state, reference_hold_steps = self.update_ref(state, reference_hold_steps)
        if self.generate_reward is not None:
            reward = self.generate_reward(state, action, self.env.env_properties)
        if self.generate_terminated is not None:
            terminated = self.generate_terminated(state, reward, self.env.env_properties)
        if self.generate_truncated is not None:
            truncated = self.generate_truncated(state, self.env.env_properties)
        return obs, reward, terminated, truncated, tree_flatten(state)[0], reference_hold_steps

    def reset(self, rng_env=None, rng_ref=None, initial_state=None):
        """Resets the environment to a random or passed initial state; ``rng_ref`` switches the reference generator
        on (reference gym_wrapper.py:132-168)."""
        if initial_state is not None:
            obs, state = self.env.vmap_reset(initial_state=tree_unflatten(self.state_tree_struct, initial_state))
        else:
            _, state = self.env.vmap_reset(rng_env)
        if rng_ref is not None:
            from .core_env import _as_generator, _key_tensor

            self._ref_rng = _as_generator(rng_ref, self.env.device)
            state.PRNGKey = _key_tensor(rng_ref, self.env.batch_size, self.env.device)
            self.ref_gen = True
            state, self.reference_hold_steps = self.generate_new_ref(
                state, torch.ones(self.env.batch_size, dtype=torch.bool, device=self.env.device),
                torch.zeros((self.env.batch_size, 1), device=self.env.device))
        else:
            self.ref_gen = False
            print("Since no PRNGKey for reference was provided, reference generation is deactivated.")
        self.state, self.state_tree_struct = tree_flatten(state)
        obs = self.env.vmap_generate_observation(state)
        return obs, {}

    def update_ref(self, state, hold_steps):
        """New references where the hold counter ran out, then count down (reference gym_wrapper.py:170-175)."""
        expired = hold_steps[:, 0] == 0
        state, hold_steps = self.generate_new_ref(state, expired, hold_steps)
        return state, hold_steps - 1

    def generate_new_ref(self, state, mask, hold_steps):
        """Random reference for the tracked fields of the masked batches (reference gym_wrapper.py:177-192): the
        physical state of a freshly drawn random initial state, and a new random hold time."""
        env = self.env
        B = env.batch_size
        _, fresh = env.vmap_reset(self._ref_rng)
        for name in self.control_state:
            cur = getattr(state.reference, name)
            new = getattr(fresh.physical_state, name)
            setattr(state.reference, name, torch.where(mask, new, torch.as_tensor(cur, device=env.device).to(new.dtype)))
        draw = torch.randint(self.ref_params["hold_steps_min"], self.ref_params["hold_steps_max"], (B, 1),
                             generator=self._ref_rng, device=env.device).to(hold_steps.dtype)
        hold_steps = torch.where(mask[:, None], draw, hold_steps)
        return state, hold_steps

    def render(self, *_, **__):
        raise NotImplementedError("To be implemented!")

    def close(self):
        raise NotImplementedError("To be implemented!")
