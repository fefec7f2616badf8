"""Helpers for the `-m gpu` parity tests: build the product environment that corresponds to an oracle
configuration and move numpy inputs to CUDA tensors."""
import numpy as np
import torch

from oracle import oracle as O

CLASS = {"pendulum": "Pendulum", "mass_spring_damper": "MassSpringDamper", "fluid_tank": "FluidTank",
         "cart_pole": "CartPole", "acrobot": "Acrobot", "pmsm": "PMSM"}


def make_env(ex, orc, B, **kw):
    """Product env with the same configuration as the oracle object."""
    tdt = torch.float64 if orc.dtype == np.float64 else torch.float32
    solver = ex.Tsit5() if orc.solver == O.TSIT5 else ex.Euler()

    def conv(v):
        return torch.as_tensor(v, dtype=tdt, device="cuda") if isinstance(v, np.ndarray) else v

    pn = {k: ex.MinMaxNormalization(conv(v[0]), conv(v[1])) for k, v in orc.pn.items()}
    an = {k: ex.MinMaxNormalization(conv(v[0]), conv(v[1])) for k, v in orc.an.items()}
    static = {k: conv(v) for k, v in orc.static.items()}
    cls = getattr(ex, CLASS[orc.env])
    extra = {}
    if orc.env == "pmsm":
        extra = dict(saturated=orc.saturated,
                     motor_variant=ex.MotorVariant[orc.motor] if orc.motor else ex.MotorVariant.DEFAULT)
    env = cls(batch_size=B, physical_normalizations=pn, action_normalizations=an, static_params=static,
              control_state=list(orc.control_state), solver=solver, tau=orc.tau, dtype=tdt, **extra, **kw)
    env.time_rule = orc.time_rule
    return env


def to_state(env, leaves, ref=None):
    """State pytree from numpy physical-state leaves."""
    B = leaves[0].shape[0]
    dev, dt = env.device, env.dtype
    phys = [torch.as_tensor(x, dtype=dt, device=dev) for x in leaves]
    nan = torch.full((B,), float("nan"), dtype=dt, device=dev)
    refs = [torch.as_tensor(x, dtype=dt, device=dev) for x in ref] if ref is not None else [nan] * len(leaves)
    return env._new_state(phys, nan, env._solver_state_dummy(B), torch.zeros(B, dtype=torch.bool, device=dev), refs)


def np_(x):
    return x.detach().cpu().numpy()
