"""ORACLE (test infrastructure, NOT product code): ctypes driver for oracle/exenv_oracle.c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  The product package never does (and fails loudly when its CUDA library is missing).

The C file restates the reference algorithm (see the file:line citations there); this module only marshals
numpy arrays.  Environment tables below restate the reference defaults:
pendulum_env.py:84-97, mass_spring_damper_env.py:82-95, fluid_tank_env.py:34-50, cart_pole_env.py:86-107,
acrobot_env.py:91-116, pmsm/motor_parameters.py:70-149.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libexenv_oracle.so"

MAX_PARAMS = 32
MAX_CONTROL = 8

EULER, TSIT5 = 0, 1


class ExoDesc(C.Structure):
    _fields_ = [
        ("env", C.c_int32),
        ("solver", C.c_int32),
        ("n_phys", C.c_int32),
        ("n_ode", C.c_int32),
        ("n_act", C.c_int32),
        ("n_obs", C.c_int32),
        ("n_static", C.c_int32),
        ("n_control", C.c_int32),
        ("control", C.c_int32 * MAX_CONTROL),
        ("saturated", C.c_int32),
        ("pmsm_deadtime", C.c_int32),
        ("time_rule", C.c_int32),
        ("pad_", C.c_int32),
        ("tau", C.c_double),
        ("pmsm_adv_coeff", C.c_double),
        ("p", C.c_double * MAX_PARAMS),
        ("pp", C.c_void_p * MAX_PARAMS),
        ("lut_maps", C.c_void_p),
        ("lut_gx", C.c_void_p),
        ("lut_gy", C.c_void_p),
        ("lut_nx", C.c_int32),
        ("lut_ny", C.c_int32),
    ]


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, a few seconds)."""
    srcs = [_HERE / "exenv_oracle.c", _HERE / "exenv_oracle_impl.h", _HERE / "exenv_oracle.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        assert _lib.exo_desc_size() == C.sizeof(ExoDesc), "ExoDesc layout mismatch"
        for sfx in ("f32", "f64"):
            getattr(_lib, f"exo_sim_ahead_{sfx}").restype = C.c_long
    return _lib


PI = math.pi


@dataclass
class EnvSpec:
    id: int
    phys: list
    n_ode: int
    act: list
    n_obs: int
    static: list
    static_defaults: dict
    phys_norm: dict
    act_norm: dict
    tau: float


SPECS = {
    "pendulum": EnvSpec(0, ["theta", "omega"], 2, ["torque"], 2, ["g", "l", "m"],
                        dict(g=9.81, l=2, m=1), dict(theta=(-PI, PI), omega=(-10, 10)), dict(torque=(-20, 20)), 1e-4),
    "mass_spring_damper": EnvSpec(1, ["deflection", "velocity"], 2, ["force"], 2, ["d", "k", "m"],
                                  dict(k=100, d=1, m=1), dict(deflection=(-10, 10), velocity=(-10, 10)),
                                  dict(force=(-20, 20)), 1e-4),
    "fluid_tank": EnvSpec(2, ["height"], 1, ["inflow"], 1, ["base_area", "orifice_area", "c_d", "g"],
                          dict(base_area=PI, orifice_area=PI * 0.1 ** 2, c_d=0.6, g=9.81), dict(height=(0, 3)),
                          dict(inflow=(0, 0.2)), 1e-3),
    "cart_pole": EnvSpec(3, ["deflection", "velocity", "theta", "omega"], 4, ["force"], 4,
                         ["mu_p", "mu_c", "l", "m_p", "m_c", "g"],
                         dict(mu_p=0.000002, mu_c=0.0005, l=0.5, m_p=0.1, m_c=1, g=9.81),
                         dict(deflection=(-2.4, 2.4), velocity=(-8, 8), theta=(-PI, PI), omega=(-8, 8)),
                         dict(force=(-20, 20)), 2e-2),
    "acrobot": EnvSpec(4, ["theta_1", "theta_2", "omega_1", "omega_2"], 4, ["torque"], 4,
                       ["g", "l_1", "l_2", "m_1", "m_2", "l_c1", "l_c2", "I_1", "I_2"],
                       dict(g=9.81, l_1=2, l_2=2, m_1=1, m_2=1, l_c1=1, l_c2=1, I_1=1.3, I_2=1.3),
                       dict(theta_1=(-PI, PI), theta_2=(-PI, PI), omega_1=(-10, 10), omega_2=(-10, 10)),
                       dict(torque=(-20, 20)), 1e-3),
    "pmsm": EnvSpec(5, ["u_d_buffer", "u_q_buffer", "epsilon", "i_d", "i_q", "torque", "omega_el"], 3,
                    ["u_d", "u_q"], 8, ["p", "r_s", "l_d", "l_q", "psi_p", "u_dc", "deadtime"],
                    dict(p=3, r_s=15e-3, l_d=0.37e-3, l_q=1.2e-3, psi_p=65.6e-3, u_dc=400, deadtime=1),
                    dict(u_d_buffer=(-2 * 400 / 3, 2 * 400 / 3), u_q_buffer=(-2 * 400 / 3, 2 * 400 / 3),
                         epsilon=(-PI, PI), i_d=(-250, 0), i_q=(-250, 250), torque=(-200, 200),
                         omega_el=(0, 3 * 11000 * 2 * PI / 60)),
                    dict(u_d=(-2 * 400 / 3, 2 * 400 / 3), u_q=(-2 * 400 / 3, 2 * 400 / 3)), 1e-4),
}

# pmsm/motor_parameters.py:70-122
PMSM_MOTORS = {
    "BRUSA": dict(static=dict(p=3, r_s=17.932e-3, l_d=0.37e-3, l_q=1.2e-3, psi_p=65.65e-3, u_dc=400, deadtime=1),
                  phys_norm=dict(u_d_buffer=(-2 * 400 / 3, 2 * 400 / 3), u_q_buffer=(-2 * 400 / 3, 2 * 400 / 3),
                                 epsilon=(-PI, PI), i_d=(-250, 0), i_q=(-250, 250), torque=(-200, 200),
                                 omega_el=(0, 3 * 11000 * 2 * PI / 60)),
                  act_norm=dict(u_d=(-2 * 400 / 3, 2 * 400 / 3), u_q=(-2 * 400 / 3, 2 * 400 / 3))),
    "SEW": dict(static=dict(p=4, r_s=208e-3, l_d=1.44e-3, l_q=1.44e-3, psi_p=122e-3, u_dc=550, deadtime=1),
                phys_norm=dict(u_d_buffer=(-2 * 550 / 3, 2 * 550 / 3), u_q_buffer=(-2 * 550 / 3, 2 * 550 / 3),
                               epsilon=(-PI, PI), i_d=(-16, 0), i_q=(-16, 16), torque=(-15, 15),
                               omega_el=(0, 4 * 2000 / 60 * 2 * PI)),
                act_norm=dict(u_d=(-2 * 550 / 3, 2 * 550 / 3), u_q=(-2 * 550 / 3, 2 * 550 / 3))),
}


def prepare_lut(mat: dict):
    """Restates PMSM.generate_interpolators_and_lut (pmsm/pmsm_env.py:316-363): NaN fill by nearest
    neighbour on index coordinates, one-cell edge replication, padded regular grids.

    Returns (gx, gy, maps) with maps[q] of shape (nx, ny) = values[x_idx][y_idx] in float64."""
    from scipy.interpolate import griddata

    quants = ["L_dd", "L_dq", "L_qd", "L_qq", "Psi_d", "Psi_q"]
    i_d_max, i_q_max = np.max(mat["i_d_vec"]), np.max(mat["i_q_vec"])
    i_d_min, i_q_min = np.min(mat["i_d_vec"]), np.min(mat["i_q_vec"])
    dd = (i_d_max - i_d_min) / (mat["i_d_vec"].shape[1] - 1)
    dq = (i_q_max - i_q_min) / (mat["i_q_vec"].shape[1] - 1)
    out = []
    for q in quants:
        qmap = np.array(mat[q], copy=True)
        x, y = np.indices(qmap.shape)
        nan_mask = np.isnan(qmap)
        if nan_mask.any():
            qmap[nan_mask] = griddata((x[~nan_mask], y[~nan_mask]), qmap[~nan_mask], (x[nan_mask], y[nan_mask]),
                                      method="nearest")
        a = np.vstack([qmap[0, :], qmap, qmap[-1, :]])
        b = np.hstack([a[:, :1], a, a[:, -1:]])
        out.append(b)
    ny, nx = out[0].shape
    gx = np.linspace(i_d_min - dd, i_d_max + dd, nx)
    gy = np.linspace(i_q_min - dq, i_q_max + dq, ny)
    # float32 maps stay float32-valued (jnp keeps the .mat dtype); arithmetic is done in the working dtype
    maps = np.stack([np.asarray(m.T, dtype=np.float64) for m in out])  # (6, nx, ny)
    return gx, gy, maps


def load_motor_lut(name: str):
    from scipy.io import loadmat

    path = _HERE.parent / "exciting-environments_b200" / "pmsm" / "data" / f"LUT_{name}_jax_grad.mat"
    return prepare_lut(loadmat(path))


@dataclass
class Oracle:
    """One configured environment; mirrors the constructor kwargs of the reference env classes."""

    env: str
    solver: int = EULER
    dtype: type = np.float64
    tau: float | None = None
    static_params: dict | None = None
    physical_normalizations: dict | None = None
    action_normalizations: dict | None = None
    control_state: list = field(default_factory=list)
    saturated: bool = False
    motor: str | None = None  # PMSM: "BRUSA" / "SEW" / None (DEFAULT)
    time_rule: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        self.spec = SPECS[self.env]
        sp = self.spec
        static, pn, an = dict(sp.static_defaults), dict(sp.phys_norm), dict(sp.act_norm)
        if self.env == "pmsm" and self.motor:
            m = PMSM_MOTORS[self.motor]
            static, pn, an = dict(m["static"]), dict(m["phys_norm"]), dict(m["act_norm"])
        if self.static_params:
            static = dict(self.static_params)
        if self.physical_normalizations:
            pn = dict(self.physical_normalizations)
        if self.action_normalizations:
            an = dict(self.action_normalizations)
        if self.tau is None:
            self.tau = sp.tau
        self.static, self.pn, self.an = static, pn, an
        d = ExoDesc()
        d.env, d.solver = sp.id, self.solver
        d.n_phys, d.n_ode, d.n_act, d.n_obs, d.n_static = len(sp.phys), sp.n_ode, len(sp.act), sp.n_obs, len(sp.static)
        d.n_control = len(self.control_state)
        for j, name in enumerate(self.control_state):
            d.control[j] = sp.phys.index(name)
        d.saturated = int(self.saturated)
        d.time_rule = self.time_rule
        d.tau = self.tau
        vals = [static[k] for k in sp.static]
        for k in sp.phys:
            vals += list(pn[k])
        for k in sp.act:
            vals += list(an[k])
        for i, v in enumerate(vals):
            if isinstance(v, np.ndarray) and v.ndim == 1:
                arr = np.ascontiguousarray(v, dtype=self.dtype)
                self._keep.append(arr)
                d.pp[i] = arr.ctypes.data
                d.p[i] = float("nan")
            else:
                d.p[i] = float(v) if v is not None else float("nan")
        if self.env == "pmsm":
            dead = static["deadtime"]
            d.pmsm_deadtime = int(dead)
            d.pmsm_adv_coeff = (dead + 0.5) * self.tau
            if self.saturated:
                gx, gy, maps = load_motor_lut(self.motor)
                gx, gy = np.ascontiguousarray(gx, self.dtype), np.ascontiguousarray(gy, self.dtype)
                maps = np.ascontiguousarray(maps, self.dtype)
                self._keep += [gx, gy, maps]
                d.lut_maps, d.lut_gx, d.lut_gy = maps.ctypes.data, gx.ctypes.data, gy.ctypes.data
                d.lut_nx, d.lut_ny = gx.size, gy.size
        self.desc = d
        self.D, self.A, self.node = d.n_phys, d.n_act, d.n_ode
        self.O = d.n_obs + d.n_control
        self.n_trunc = 1 if self.env in ("fluid_tank", "pmsm") else self.O
        self._sfx = "f64" if self.dtype == np.float64 else "f32"

    # -- helpers ---------------------------------------------------------------------------------
    def _ptrs(self, arrays):
        arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        return arr

    def _leaves(self, x, n):
        x = [np.ascontiguousarray(a, dtype=self.dtype) for a in x]
        assert len(x) == n
        return x

    # -- API -------------------------------------------------------------------------------------
    def step(self, state, action, ref=None, epilogue=False):
        """state: list of D (B,) arrays (physical SI values); action (B, A) normalised.
        Returns dict(obs, state, k (Tsit5), reward, truncated, terminated)."""
        state = self._leaves(state, self.D)
        B = state[0].shape[0]
        action = np.ascontiguousarray(action, dtype=self.dtype).reshape(B, self.A)
        out = [np.empty(B, self.dtype) for _ in range(self.D)]
        obs = np.empty((B, self.O), self.dtype)
        k = [np.empty(B, self.dtype) for _ in range(self.node)]
        refl = self._leaves(ref, self.D) if ref is not None else None
        rew = np.empty(B, self.dtype) if epilogue else None
        tr = np.empty((B, self.n_trunc), np.uint8) if epilogue else None
        te = np.empty(B, np.uint8) if epilogue else None
        fn = getattr(lib(), f"exo_step_{self._sfx}")
        rc = fn(C.byref(self.desc), C.c_long(B), self._ptrs(state), C.c_void_p(action.ctypes.data), self._ptrs(out),
                C.c_void_p(obs.ctypes.data), self._ptrs(k), self._ptrs(refl) if refl else None,
                C.c_void_p(rew.ctypes.data) if epilogue else None, C.c_void_p(tr.ctypes.data) if epilogue else None,
                C.c_void_p(te.ctypes.data) if epilogue else None)
        assert rc == 0
        return dict(obs=obs, state=out, k=k, reward=rew, truncated=tr, terminated=te)

    def n_save(self, n_actions, obs_stepsize, action_stepsize):
        t1 = action_stepsize * n_actions
        return 1 + int(t1 / obs_stepsize)

    def sim_ahead(self, state, actions, obs_stepsize, action_stepsize, ref=None, epilogue=False, max_steps=4096):
        """state: list of D (B,) arrays; actions (B, T, A) normalised.
        Returns dict(obs (B,N+1,O), traj [D x (B,N+1)], last [D x (B,)], k_last, n_steps, reward, truncated,
        terminated)."""
        state = self._leaves(state, self.D)
        B = state[0].shape[0]
        actions = np.ascontiguousarray(actions, dtype=self.dtype)
        T = actions.shape[1]
        assert actions.shape == (B, T, self.A)
        ns = self.n_save(T, obs_stepsize, action_stepsize)
        obs = np.empty((B, ns, self.O), self.dtype)
        traj = [np.empty((B, ns), self.dtype) for _ in range(self.D)]
        last = [np.empty(B, self.dtype) for _ in range(self.D)]
        kl = [np.empty(B, self.dtype) for _ in range(self.node)]
        refl = self._leaves(ref, self.D) if ref is not None else None
        rew = np.empty((B, ns - 1), self.dtype) if epilogue else None
        tr = np.empty((B, ns, self.n_trunc), np.uint8) if epilogue else None
        te = np.empty((B, ns - 1), np.uint8) if epilogue else None
        fn = getattr(lib(), f"exo_sim_ahead_{self._sfx}")
        rc = fn(C.byref(self.desc), C.c_long(B), C.c_long(T), C.c_double(obs_stepsize), C.c_double(action_stepsize),
                self._ptrs(state), C.c_void_p(actions.ctypes.data), C.c_void_p(obs.ctypes.data), self._ptrs(traj),
                self._ptrs(last), self._ptrs(kl), self._ptrs(refl) if refl else None, C.c_long(ns),
                C.c_long(max_steps), C.c_void_p(rew.ctypes.data) if epilogue else None,
                C.c_void_p(tr.ctypes.data) if epilogue else None, C.c_void_p(te.ctypes.data) if epilogue else None)
        if rc == -3:
            raise RuntimeError("The maximum number of solver steps was reached. Try increasing `max_steps`.")
        assert rc >= 0, rc
        return dict(obs=obs, traj=traj, last=last, k_last=kl, n_steps=int(rc), reward=rew, truncated=tr,
                    terminated=te)

    def reset(self, B, mode=0, values=None):
        """mode 0 default reset; 1 values = (B, D) normalised (PMSM: physical) state; 2 values = obs (B, O)."""
        out = [np.empty(B, self.dtype) for _ in range(self.D)]
        ref = [np.empty(B, self.dtype) for _ in range(self.D)]
        obs = np.empty((B, self.O), self.dtype)
        vin = None
        if mode:
            vin = np.ascontiguousarray(values, dtype=self.dtype)
        fn = getattr(lib(), f"exo_reset_{self._sfx}")
        rc = fn(C.byref(self.desc), C.c_long(B), C.c_int(mode), C.c_void_p(vin.ctypes.data) if mode else None,
                self._ptrs(out), self._ptrs(ref), C.c_void_p(obs.ctypes.data))
        assert rc == 0
        return dict(obs=obs, state=out, ref=ref)
