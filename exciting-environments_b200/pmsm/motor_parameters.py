"""Motor variants of the PMSM environment (reference pmsm/motor_parameters.py:70-163): normalisation ranges,
static parameters and, for BRUSA / SEW, the flux-map LUT files shipped next to this module (verbatim data copies
of the reference's MIT-licensed ``LUT_*_jax_grad.mat``)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from enum import Enum
from typing import Callable

import torch

from ..utils import MinMaxNormalization

PI = math.pi


def default_soft_constraints(self, state, action_norm, env_properties):
    """relu(|x_norm| - 1) for every physical state (reference motor_parameters.py:56-67)."""
    norm = self.normalize_state(state, env_properties).physical_state
    vals = {n: torch.clamp(torch.abs(torch.as_tensor(getattr(norm, n))) - 1.0, min=0) for n in self.PHYS_FIELDS}
    return self.PhysicalState(**vals), None


@dataclass
class MotorParams:
    physical_normalizations: dict
    action_normalizations: dict
    static_params: dict
    default_soft_constraints: Callable
    lut_name: str | None


def _motor(u_dc, i_lim, omega_max, torque_lim, static, lut_name):
    u = 2 * u_dc / 3
    return MotorParams(
        physical_normalizations=dict(
            u_d_buffer=MinMaxNormalization(-u, u), u_q_buffer=MinMaxNormalization(-u, u),
            epsilon=MinMaxNormalization(-PI, PI), i_d=MinMaxNormalization(-i_lim, 0),
            i_q=MinMaxNormalization(-i_lim, i_lim), omega_el=MinMaxNormalization(0, omega_max),
            torque=MinMaxNormalization(-torque_lim, torque_lim)),
        action_normalizations=dict(u_d=MinMaxNormalization(-u, u), u_q=MinMaxNormalization(-u, u)),
        static_params=static,
        default_soft_constraints=default_soft_constraints,
        lut_name=lut_name,
    )


_TABLE = {
    "BRUSA": lambda: _motor(400, 250, 3 * 11000 * 2 * PI / 60, 200,
                            dict(p=3, r_s=17.932e-3, l_d=0.37e-3, l_q=1.2e-3, psi_p=65.65e-3, u_dc=400, deadtime=1),
                            "BRUSA"),
    "SEW": lambda: _motor(550, 16, 4 * 2000 / 60 * 2 * PI, 15,
                          dict(p=4, r_s=208e-3, l_d=1.44e-3, l_q=1.44e-3, psi_p=122e-3, u_dc=550, deadtime=1), "SEW"),
    "DEFAULT": lambda: _motor(400, 250, 3 * 11000 * 2 * PI / 60, 200,
                              dict(p=3, r_s=15e-3, l_d=0.37e-3, l_q=1.2e-3, psi_p=65.6e-3, u_dc=400, deadtime=1),
                              None),
}


class MotorVariant(Enum):
    DEFAULT = "DEFAULT"
    BRUSA = "BRUSA"
    SEW = "SEW"

    def get_params(self) -> MotorParams:
        return _TABLE[self.value]()

    def prepared_lut(self):
        from . import lut

        name = _TABLE[self.value]().lut_name
        if name is None:
            raise ValueError("the DEFAULT motor has no flux-map LUT")
        return lut.load(name)
