"""``EnvironmentRegistry`` with the reference's member names and ``.make(**kwargs)`` factory
(reference registration.py:12-32)."""
from enum import Enum

from .envs import PMSM, Acrobot, CartPole, FluidTank, MassSpringDamper, Pendulum


class EnvironmentRegistry(Enum):
    CART_POLE = "CartPole-v0"
    MASS_SPRING_DAMPER = "MassSpringDamper-v0"
    PENDULUM = "Pendulum-v0"
    FLUID_TANK = "FluidTank-v0"
    PMSM = "PMSM-v0"
    ACROBOT = "Acrobot-v0"

    def make(self, **env_kwargs):
        table = {"CART_POLE": CartPole, "MASS_SPRING_DAMPER": MassSpringDamper, "PENDULUM": Pendulum,
                 "FLUID_TANK": FluidTank, "PMSM": PMSM, "ACROBOT": Acrobot}
        cls = table.get(self.name)
        if cls is None:
            raise ValueError(f"Unknown environment: {self}")
        return cls(**env_kwargs)
