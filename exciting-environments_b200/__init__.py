"""exciting-environments_b200: the batched dynamics hot path of ExcitingSystems/exciting-environments
(``vmap_step`` / ``vmap_sim_ahead`` / ``vmap_reset`` state packing for pendulum, cart-pole, acrobot,
mass-spring-damper, fluid tank and PMSM) as hand-written sm_100a CUDA kernels behind a C ABI, with the reference's
Python class surface on top.

    import exciting_environments_b200 as excenvs          # root-level alias of this package
    env = excenvs.EnvironmentRegistry.PENDULUM.make(batch_size=2**20, solver=excenvs.Tsit5(), tau=1e-4)
    obs, state = env.vmap_reset()
    obs, state = env.vmap_step(state, action)
    observations, states, last_state = env.vmap_sim_ahead(state, actions, env.tau, env.tau)
"""
from . import _lib
from ._lib import ExenvError, build, launch_count
from .core_env import CoreEnvironment, config
from .envs import PMSM, Acrobot, CartPole, FluidTank, MassSpringDamper, Pendulum
from .gym_wrapper import GymWrapper
from .pmsm import MotorVariant
from .registration import EnvironmentRegistry
from .solvers import Euler, Tsit5
from .utils import (MinMaxNormalization, dump_sim_properties_to_json, load_sim_properties_from_json, tree_flatten,
                    tree_structure, tree_unflatten)

__all__ = [
    "CoreEnvironment", "PMSM", "CartPole", "MassSpringDamper", "Pendulum", "FluidTank", "Acrobot",
    "EnvironmentRegistry", "GymWrapper", "MotorVariant", "MinMaxNormalization", "Euler", "Tsit5", "config",
    "ExenvError", "build", "launch_count", "tree_flatten", "tree_unflatten", "tree_structure",
    "load_sim_properties_from_json", "dump_sim_properties_to_json",
]
