"""The six ODE environments of the reference as thin, table-driven front ends of the CUDA kernels.

Each class keeps the reference's constructor signature and defaults:
  Pendulum          exciting_environments/pendulum/pendulum_env.py:52-114
  MassSpringDamper  exciting_environments/mass_spring_damper/mass_spring_damper_env.py:49-112
  FluidTank         exciting_environments/fluid_tank/fluid_tank_env.py:22-67
  CartPole          exciting_environments/cart_pole/cart_pole_env.py:49-124
  Acrobot           exciting_environments/acrobot/acrobot_env.py:50-133
  PMSM              exciting_environments/pmsm/pmsm_env.py:116-267
The physics itself lives in csrc/exenv_envs.cuh; nothing numeric happens in this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import MethodType
from typing import Callable

import torch

from . import _lib
from .core_env import CoreEnvironment
from .pmsm.motor_parameters import MotorVariant
from .utils import MinMaxNormalization

PI = math.pi


def _relu(x):
    return torch.clamp(torch.as_tensor(x), min=0)


class _TableEnv(CoreEnvironment):
    """Shared constructor logic: defaults -> dataclasses -> EnvProperties (e.g. pendulum_env.py:84-114)."""

    DEFAULT_PHYS_NORM: dict = {}
    DEFAULT_ACTION_NORM: dict = {}
    DEFAULT_STATIC: dict = {}
    DEFAULT_TAU: float = 1e-4
    SOFT_CONSTRAINED: tuple = ()  # fields with relu(|x_norm| - 1) soft constraints

    def __init__(self, batch_size: int = 8, physical_normalizations: dict = None, action_normalizations: dict = None,
                 soft_constraints: Callable = None, static_params: dict = None, control_state: list = None,
                 solver=None, tau: float = None, dtype=None, device=None):
        if not physical_normalizations:
            physical_normalizations = {k: MinMaxNormalization(*v) for k, v in self.DEFAULT_PHYS_NORM.items()}
        if not action_normalizations:
            action_normalizations = {k: MinMaxNormalization(*v) for k, v in self.DEFAULT_ACTION_NORM.items()}
        if not soft_constraints:
            soft_constraints = self.default_soft_constraints
        if not static_params:
            static_params = dict(self.DEFAULT_STATIC)
        if not control_state:
            control_state = []
        if tau is None:
            tau = self.DEFAULT_TAU
        self.control_state = control_state
        self.soft_constraints = soft_constraints
        env_properties = self.EnvProperties(
            physical_normalizations=self.PhysicalState(**physical_normalizations),
            action_normalizations=self.Action(**action_normalizations),
            static_params=self.StaticParams(**static_params),
        )
        super().__init__(batch_size, env_properties=env_properties, tau=tau, solver=solver, dtype=dtype, device=device)

    def default_soft_constraints(self, state, action_norm, env_properties):
        """relu(|x_norm| - 1) on the constrained fields, NaN elsewhere; relu(|a| - 1) on the action
        (e.g. pendulum_env.py:366-379, cart_pole_env.py:427-442)."""
        norm = self.normalize_state(state, env_properties).physical_state
        vals = {}
        for n in self.PHYS_FIELDS:
            vals[n] = _relu(torch.abs(torch.as_tensor(getattr(norm, n))) - 1.0) if n in self.SOFT_CONSTRAINED else float("nan")
        return self.PhysicalState(**vals), _relu(torch.abs(torch.as_tensor(action_norm)) - 1.0)


class Pendulum(_TableEnv):
    """State ``['theta', 'omega']``, action ``['torque']``; default start theta=pi, omega=0."""

    ENV_ID = 0
    PHYS_FIELDS = ("theta", "omega")
    ACTION_FIELDS = ("torque",)
    STATIC_FIELDS = ("g", "l", "m")
    N_ODE, N_OBS = 2, 2
    DEFAULT_PHYS_NORM = {"theta": (-PI, PI), "omega": (-10, 10)}
    DEFAULT_ACTION_NORM = {"torque": (-20, 20)}
    DEFAULT_STATIC = {"g": 9.81, "l": 2, "m": 1}
    DEFAULT_TAU = 1e-4
    SOFT_CONSTRAINED = ("omega",)


class MassSpringDamper(_TableEnv):
    """State ``['deflection', 'velocity']``, action ``['force']``."""

    ENV_ID = 1
    PHYS_FIELDS = ("deflection", "velocity")
    ACTION_FIELDS = ("force",)
    STATIC_FIELDS = ("d", "k", "m")
    N_ODE, N_OBS = 2, 2
    DEFAULT_PHYS_NORM = {"deflection": (-10, 10), "velocity": (-10, 10)}
    DEFAULT_ACTION_NORM = {"force": (-20, 20)}
    DEFAULT_STATIC = {"k": 100, "d": 1, "m": 1}
    DEFAULT_TAU = 1e-4
    SOFT_CONSTRAINED = ("deflection", "velocity")


class FluidTank(_TableEnv):
    """State ``['height']``, action ``['inflow']``; Torricelli outflow, height clipped at 0."""

    ENV_ID = 2
    PHYS_FIELDS = ("height",)
    ACTION_FIELDS = ("inflow",)
    STATIC_FIELDS = ("base_area", "orifice_area", "c_d", "g")
    N_ODE, N_OBS = 1, 1
    DEFAULT_PHYS_NORM = {"height": (0, 3)}
    DEFAULT_ACTION_NORM = {"inflow": (0, 0.2)}
    DEFAULT_STATIC = {"base_area": PI, "orifice_area": PI * 0.1 ** 2, "c_d": 0.6, "g": 9.81}
    DEFAULT_TAU = 1e-3

    def _draw_norm_state(self, rng, B):  # uniform(0, 1), fluid_tank_env.py:227
        from .core_env import _as_generator

        gen = _as_generator(rng, self.device)
        return torch.rand((B, 1), generator=gen, device=self.device, dtype=self.dtype)

    @property
    def states_description(self):
        import numpy as np

        return np.array(["fluid height"])

    def _obs_names(self):
        return ["fluid height"]


class CartPole(_TableEnv):
    """State ``['deflection', 'velocity', 'theta', 'omega']``, action ``['force']``."""

    ENV_ID = 3
    PHYS_FIELDS = ("deflection", "velocity", "theta", "omega")
    ACTION_FIELDS = ("force",)
    STATIC_FIELDS = ("mu_p", "mu_c", "l", "m_p", "m_c", "g")
    N_ODE, N_OBS = 4, 4
    DEFAULT_PHYS_NORM = {"deflection": (-2.4, 2.4), "velocity": (-8, 8), "theta": (-PI, PI), "omega": (-8, 8)}
    DEFAULT_ACTION_NORM = {"force": (-20, 20)}
    DEFAULT_STATIC = {"mu_p": 0.000002, "mu_c": 0.0005, "l": 0.5, "m_p": 0.1, "m_c": 1, "g": 9.81}
    DEFAULT_TAU = 2e-2
    SOFT_CONSTRAINED = ("deflection", "velocity", "omega")


class Acrobot(_TableEnv):
    """State ``['theta_1', 'theta_2', 'omega_1', 'omega_2']``, action ``['torque']``."""

    ENV_ID = 4
    PHYS_FIELDS = ("theta_1", "theta_2", "omega_1", "omega_2")
    ACTION_FIELDS = ("torque",)
    STATIC_FIELDS = ("g", "l_1", "l_2", "m_1", "m_2", "l_c1", "l_c2", "I_1", "I_2")
    N_ODE, N_OBS = 4, 4
    DEFAULT_PHYS_NORM = {"theta_1": (-PI, PI), "theta_2": (-PI, PI), "omega_1": (-10, 10), "omega_2": (-10, 10)}
    DEFAULT_ACTION_NORM = {"torque": (-20, 20)}
    DEFAULT_STATIC = {"g": 9.81, "l_1": 2, "l_2": 2, "m_1": 1, "m_2": 1, "l_c1": 1, "l_c2": 1, "I_1": 1.3, "I_2": 1.3}
    DEFAULT_TAU = 1e-3
    SOFT_CONSTRAINED = ("omega_1", "omega_2")


@dataclass
class _PmsmEnvProperties:
    """EnvProperties with the extra ``saturated`` flag (reference pmsm_env.py:307-314)."""

    saturated: object
    physical_normalizations: object
    action_normalizations: object
    static_params: object


class PMSM(CoreEnvironment):
    """Permanent-magnet synchronous motor drive in dq coordinates with inverter voltage hexagon limit, one-step
    dead time and optional saturated flux-map LUTs (reference pmsm/pmsm_env.py)."""

    ENV_ID = 5
    PHYS_FIELDS = ("u_d_buffer", "u_q_buffer", "epsilon", "i_d", "i_q", "torque", "omega_el")
    ACTION_FIELDS = ("u_d", "u_q")
    STATIC_FIELDS = ("p", "r_s", "l_d", "l_q", "psi_p", "u_dc", "deadtime")
    N_ODE, N_OBS = 3, 8
    SIM_AHEAD_NAN_REFERENCE = True  # pmsm_env.py:688-701

    EnvProperties = _PmsmEnvProperties

    def __init__(self, batch_size: int = 8, saturated=False, motor_variant: MotorVariant = MotorVariant.DEFAULT,
                 physical_normalizations: dict = None, action_normalizations: dict = None,
                 soft_constraints: Callable = None, static_params: dict = None, control_state: list = None,
                 solver=None, tau: float = 1e-4, dtype=None, device=None):
        motor = motor_variant.get_params()
        default_pn, default_an = dict(motor.physical_normalizations), dict(motor.action_normalizations)
        default_static = dict(motor.static_params)
        self._lut = None
        if motor_variant != MotorVariant.DEFAULT:
            if saturated:
                default_static["l_d"] = float("nan")
                default_static["l_q"] = float("nan")
                default_static["psi_p"] = float("nan")
                self._lut = motor_variant.prepared_lut()
        elif saturated:
            raise ValueError(
                f"MotorVariant '{motor_variant.value}' is not allowed for saturated LUTs. "
                "Use a specific motor variant. DEFAULT is only valid for saturated=False."
            )
        self.pmsm_lut = self._lut
        if not static_params:
            static_params = default_static
        if not physical_normalizations:
            physical_normalizations = default_pn
        else:
            for key in ("i_d", "i_q"):
                lim, dlim = physical_normalizations[key], default_pn[key]
                if (lim.min < dlim.min) or (lim.max > dlim.max):
                    print(
                        f"The defined permitted range of {key} ({lim}) exceeds the limits of the LUT ({dlim}). Values outside this range are extrapolated."
                    )
        if not action_normalizations:
            action_normalizations = default_an
        if not control_state:
            control_state = []
        if not soft_constraints:
            soft_constraints = MethodType(motor.default_soft_constraints, self)
        self.control_state = control_state
        self.soft_constraints = soft_constraints
        self.saturated = bool(saturated)
        env_properties = self.EnvProperties(
            saturated=saturated,
            physical_normalizations=self.PhysicalState(**physical_normalizations),
            action_normalizations=self.Action(**action_normalizations),
            static_params=self.StaticParams(**static_params),
        )
        super().__init__(batch_size, env_properties=env_properties, tau=tau, solver=solver, dtype=dtype, device=device)
        self._action_description = ["u_d", "u_q"]
        self._obs_description = ["i_d", "i_q", "cos_eps", "sin_eps", "omega_el", "torque", "u_d_buffer", "u_q_buffer"]

    def _plan_extra_key(self, env_properties):
        return (bool(self.saturated), id(self._lut))

    def _desc_extra(self, desc, keep):
        import ctypes as C

        import numpy as np

        if self.saturated:
            desc.flags |= _lib.FLAG_SATURATED
            gx, gy, maps = self._lut
            gx, gy = np.ascontiguousarray(gx, np.float64), np.ascontiguousarray(gy, np.float64)
            maps = np.ascontiguousarray(maps, np.float64)
            keep += [gx, gy, maps]
            desc.lut_gx, desc.lut_gy, desc.lut_maps = gx.ctypes.data, gy.ctypes.data, maps.ctypes.data
            desc.lut_nx, desc.lut_ny = gx.size, gy.size

    def _draw_norm_state(self, rng, B):
        """Random initial state (pmsm_env.py:402-456): currents uniform in the disc of radius i_max reflected into
        the permitted box, epsilon / omega_el uniform in their ranges, zero buffers; torque is filled by the kernel
        path (values are physical already)."""
        from .core_env import _as_generator

        gen = _as_generator(rng, self.device)
        pn = self.env_properties.physical_normalizations
        u = torch.rand((B, 4), generator=gen, device=self.device, dtype=torch.float64)
        r, phi = torch.sqrt(u[:, 0]), 2 * PI * u[:, 1]  # uniform in the unit disc (jax.random.ball(key, 2))
        i_max = max(abs(float(pn.i_d.min)), abs(float(pn.i_d.max)), abs(float(pn.i_q.min)), abs(float(pn.i_q.max)))
        i_d, i_q = r * torch.cos(phi) * i_max, r * torch.sin(phi) * i_max

        def reflect(x, lo, hi):
            return x - 2 * torch.clamp(x - hi, min=0) + 2 * torch.clamp(-x + lo, min=0)

        i_d = reflect(i_d, float(pn.i_d.min), float(pn.i_d.max))
        i_q = reflect(i_q, float(pn.i_q.min), float(pn.i_q.max))
        eps = u[:, 2] * (float(pn.epsilon.max) - float(pn.epsilon.min)) + float(pn.epsilon.min)
        om = u[:, 3] * (float(pn.omega_el.max) - float(pn.omega_el.min)) + float(pn.omega_el.min)
        zero = torch.zeros_like(eps)
        torque = self._torque_host(i_d, i_q)
        return torch.stack([zero, zero, eps, i_d, i_q, torque, om], dim=1).to(self.dtype)

    def _torque_host(self, i_d, i_q):
        """Torque of a freshly drawn random state (pmsm_env.py:428-435).  Reset is not the hot path; the linear
        formula is evaluated on the host-side tensors, the saturated one by bilinear look-up of Psi_d / Psi_q."""
        sp = self.env_properties.static_params
        if not self.saturated:
            return 1.5 * sp.p * (sp.psi_p + (sp.l_d - sp.l_q) * i_d) * i_q
        from .pmsm.lut import bilinear

        gx, gy, maps = self._lut
        psi_d = bilinear(gx, gy, maps[4], i_d.cpu().numpy(), i_q.cpu().numpy())
        psi_q = bilinear(gx, gy, maps[5], i_d.cpu().numpy(), i_q.cpu().numpy())
        t = 1.5 * sp.p * (psi_d * i_q.cpu().numpy() - psi_q * i_d.cpu().numpy())
        return torch.as_tensor(t, device=i_d.device)

    @property
    def action_description(self):
        return self._action_description

    def _obs_names(self):
        return list(self._obs_description)

    def default_soft_constraints(self, state, action_norm, env_properties):
        norm = self.normalize_state(state, env_properties).physical_state
        vals = {n: _relu(torch.abs(torch.as_tensor(getattr(norm, n))) - 1.0) for n in self.PHYS_FIELDS}
        return self.PhysicalState(**vals), None
