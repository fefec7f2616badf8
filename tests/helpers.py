"""Shared test helpers: golden-vector loading, seeded synthetic inputs, ExenvDesc construction from an oracle
configuration (so that CPU-only tests can drive the host-side schedule builder and the hostcheck harness)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np

from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "reference"
GOLDEN_DIRS = {"pendulum": "pendulum", "mass_spring_damper": "mass_spring_damper", "fluid_tank": "fluid_tank",
               "cart_pole": "cartpole", "acrobot": "acrobot", "pmsm": "pmsm"}
ENVS = list(GOLDEN_DIRS)


def load_golden(env):
    d = GOLDEN / GOLDEN_DIRS[env]
    props = json.load(open(d / "sim_properties.json"))
    pn = {k: (v["min"], v["max"]) for k, v in props["physical_normalizations"].items()}
    an = {k: (v["min"], v["max"]) for k, v in props["action_normalizations"].items()}
    return dict(actions=np.load(d / "actions.npy"), observations=np.load(d / "observations.npy"),
                static=props["params"], pn=pn, an=an, tau=props["tau"])


def golden_oracle(env, solver=O.EULER, dtype=np.float64):
    g = load_golden(env)
    return O.Oracle(env, solver=solver, dtype=dtype, tau=g["tau"], static_params=g["static"],
                    physical_normalizations=g["pn"], action_normalizations=g["an"]), g


def random_state(orc, B, rng, spread=0.9, pmsm_gentle=False):
    """Physical state leaves drawn uniformly inside the normalisation box (SURVEY §8d synthetic inputs)."""
    sp = orc.spec
    leaves = []
    for name in sp.phys:
        lo, hi = orc.pn[name]
        lo, hi = (np.asarray(lo, np.float64), np.asarray(hi, np.float64))
        u = rng.uniform(-spread, spread, B)
        if orc.env == "fluid_tank":
            u = rng.uniform(0.0, 1.0, B) * 2 - 1
        leaves.append(((u + 1) / 2 * (hi - lo) + lo).astype(orc.dtype))
    if orc.env == "pmsm":  # currents inside the unit disc, buffers small, torque consistent is not required
        r, phi = np.sqrt(rng.uniform(0, 0.8, B)), rng.uniform(0, 2 * np.pi, B)
        i_max = 250.0 if orc.motor != "SEW" else 16.0
        leaves[3] = (-np.abs(r * np.cos(phi)) * i_max).astype(orc.dtype)
        leaves[4] = (r * np.sin(phi) * i_max).astype(orc.dtype)
        if pmsm_gentle:  # low speed, small buffered voltages: currents stay inside the flux-map range for a while
            leaves[6] = (leaves[6] * 0.02).astype(orc.dtype)
            leaves[0] = (leaves[0] * 0.02).astype(orc.dtype)
            leaves[1] = (leaves[1] * 0.02).astype(orc.dtype)
    return leaves


def make_desc(orc, device_params=None):
    """ExenvDesc for the product library from an oracle configuration.  Per-env arrays: `device_params` maps slot ->
    device pointer (GPU tests) else host numpy pointers are used (hostcheck)."""
    import exciting_environments_b200 as ex

    L = ex._lib
    d = L.ExenvDesc()
    od = orc.desc
    d.env, d.solver = od.env, od.solver
    d.dtype = L.F64 if orc.dtype == np.float64 else L.F32
    d.flags = L.FLAG_SATURATED if orc.saturated else 0
    d.tau = orc.tau
    d.n_control = od.n_control
    for j in range(od.n_control):
        d.control[j] = od.control[j]
    d.time_rule = orc.time_rule
    for i in range(32):
        d.p[i] = od.p[i]
        if od.pp[i]:
            d.pp[i] = device_params[i] if device_params else od.pp[i]
    return d


def hc_lut(orc):
    """HcLut struct (interleaved cells in the plan dtype) for the hostcheck harness."""
    class HcLut(C.Structure):
        _fields_ = [("gx", C.c_void_p), ("gy", C.c_void_p), ("cells", C.c_void_p), ("nx", C.c_int), ("ny", C.c_int),
                    ("inv_dx", C.c_double), ("inv_dy", C.c_double)]

    if not orc.saturated:
        return None, []
    gx, gy, maps = O.load_motor_lut(orc.motor)
    cells = np.ascontiguousarray(np.transpose(maps, (1, 2, 0)), dtype=orc.dtype)
    gxt, gyt = np.ascontiguousarray(gx, orc.dtype), np.ascontiguousarray(gy, orc.dtype)
    h = HcLut(gxt.ctypes.data, gyt.ctypes.data, cells.ctypes.data, gx.size, gy.size, 1.0 / (gx[1] - gx[0]),
              1.0 / (gy[1] - gy[0]))
    return h, [gxt, gyt, cells]


def ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def rel_err(a, b, floor=1e-30):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)) if a.size else 0.0


def field_scales(orc):
    """Half-span of every PhysicalState field's normalisation box (the natural unit of a state error)."""
    out = []
    for name in orc.spec.phys:
        lo, hi = orc.pn[name]
        out.append(float(np.max(np.abs(np.asarray(hi, np.float64) - np.asarray(lo, np.float64)))) / 2)
    return out


def scaled_err(a, b, scale):
    """max |a - b| / max(|b|, scale): relative error with the field's normalisation half-span as the floor, so that
    zero crossings of a state do not turn rounding noise into huge 'relative' errors."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale))) if a.size else 0.0


def traj_err(a, b, scale):
    """Trajectory error per environment, normalised by max(scale, max_t |b|) of that environment."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    denom = np.maximum(np.max(np.abs(b.reshape(b.shape[0], -1)), axis=1), scale)
    err = np.max(np.abs(a - b).reshape(b.shape[0], -1), axis=1)
    return float(np.max(err / denom))


def flags_from_states(orc, leaves, refs=None):
    """The reference's truncated flags, evaluated with IEEE NumPy arithmetic in the oracle's dtype on GIVEN physical
    states (any leading shape): `abs(generate_observation(state)) > 1` (pendulum_env.py:311-329, 381-385 and the same
    lines of the other environments; normalisation utils.py:13-14), the fluid tank's constant `[0]`
    (fluid_tank_env.py:323-326), PMSM `sqrt(i_d_norm**2 + i_q_norm**2)[None] > 1` (pmsm_env.py:972-979).
    Returns uint8 of shape leaves[0].shape + (n_trunc,).  This is what the kernels' flags are compared with, bit for
    bit, on the states the kernels themselves return."""
    T = orc.dtype
    names = list(orc.spec.phys)

    def norm(x, name):
        lo, hi = orc.pn[name]
        lo, hi = np.asarray(lo, T), np.asarray(hi, T)
        if lo.ndim:  # per-environment ranges: broadcast over trailing (time) axes
            lo = lo.reshape(lo.shape + (1,) * (np.ndim(x) - 1))
            hi = hi.reshape(hi.shape + (1,) * (np.ndim(x) - 1))
        x = np.asarray(x, T)
        out = T(2) * (x - lo) / (hi - lo) - T(1)
        assert out.dtype == T
        return out

    lv = dict(zip(names, leaves))
    if orc.env == "fluid_tank":
        return np.zeros(np.shape(leaves[0]) + (1,), np.uint8)
    if orc.env == "pmsm":
        idn, iqn = norm(lv["i_d"], "i_d"), norm(lv["i_q"], "i_q")
        i_s = np.sqrt(idn * idn + iqn * iqn)
        return (i_s > T(1)).astype(np.uint8)[..., None]
    cols = [np.abs(norm(lv[n], n)) > T(1) for n in names]
    if refs is not None:
        rv = dict(zip(names, refs))
        for n in orc.control_state:
            r = np.asarray(rv[n], T)
            if r.ndim < np.ndim(leaves[0]):
                r = r.reshape(r.shape + (1,) * (np.ndim(leaves[0]) - r.ndim))
            cols.append(np.broadcast_to(np.abs(norm(r, n)) > T(1), np.shape(leaves[0])))
    return np.stack(cols, axis=-1).astype(np.uint8)


def ulp_err(a, b, dtype):
    """max |a - b| in units of the last place of b (floor: the smallest normal scale of interest is b itself)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    eps = np.finfo(dtype).eps
    denom = np.maximum(np.abs(b), np.finfo(dtype).tiny) * eps
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
