"""CPU: host-side mirror of the reference interface -- defaults tables, constructor behaviour, in_axes derivation,
pytree helpers, error behaviour (reference tests/envs/<env>/test_<env>.py initialisation tests,
tests/envs/test_core_functions.py::test_tau)."""
import math

import numpy as np
import pytest
import torch


@pytest.mark.parametrize("tau", [1e-4, 1e-5])
def test_tau(excenvs, tau):
    for reg in excenvs.EnvironmentRegistry:
        assert reg.make(tau=tau).tau == tau


def test_registry_members_and_defaults(excenvs):
    ex = excenvs
    assert [m.name for m in ex.EnvironmentRegistry] == ["CART_POLE", "MASS_SPRING_DAMPER", "PENDULUM", "FLUID_TANK",
                                                         "PMSM", "ACROBOT"]
    p = ex.EnvironmentRegistry.PENDULUM.make()
    assert p.batch_size == 8 and p.tau == 1e-4 and p.action_dim == 1 and p.physical_state_dim == 2
    sp = p.env_properties.static_params
    assert (sp.g, sp.l, sp.m) == (9.81, 2, 1)
    pn = p.env_properties.physical_normalizations
    assert (pn.theta.min, pn.theta.max, pn.omega.min, pn.omega.max) == (-math.pi, math.pi, -10, 10)
    assert p.env_properties.action_normalizations.torque.max == 20
    assert list(p.obs_description) == ["theta", "omega"] and list(p.action_description) == ["torque"]
    assert ex.CartPole().tau == 2e-2 and ex.Acrobot().tau == 1e-3 and ex.FluidTank().tau == 1e-3
    m = ex.MassSpringDamper().env_properties.static_params
    assert (m.d, m.k, m.m) == (1, 100, 1)
    t = ex.FluidTank().env_properties.static_params
    assert t.base_area == math.pi and abs(t.orifice_area - math.pi * 0.01) < 1e-18 and t.c_d == 0.6
    a = ex.Acrobot().env_properties.static_params
    assert (a.l_1, a.l_c1, a.I_1) == (2, 1, 1.3)
    pm = ex.PMSM()
    s = pm.env_properties.static_params
    assert (s.p, s.r_s, s.l_d, s.l_q, s.psi_p, s.u_dc, s.deadtime) == (3, 15e-3, 0.37e-3, 1.2e-3, 65.6e-3, 400, 1)
    assert len(pm.obs_description) == 8 and pm.physical_state_dim == 7 and pm.action_dim == 2
    assert pm.env_properties.physical_normalizations.omega_el.max == 3 * 11000 * 2 * math.pi / 60
    assert list(pm.PhysicalState.__dataclass_fields__) == ["u_d_buffer", "u_q_buffer", "epsilon", "i_d", "i_q",
                                                           "torque", "omega_el"]


def test_pmsm_motor_variants_and_lut_preparation(excenvs):
    ex = excenvs
    b = ex.PMSM(saturated=True, motor_variant=ex.MotorVariant.BRUSA)
    gx, gy, maps = b.pmsm_lut
    assert gx.shape == (28,) and gy.shape == (53,) and maps.shape == (6, 28, 53)
    assert gx[0] == -260 and gx[-1] == 10 and gy[0] == -260 and gy[-1] == 260
    assert math.isnan(b.env_properties.static_params.l_d) and b.env_properties.saturated is True
    s = ex.PMSM(saturated=True, motor_variant=ex.MotorVariant.SEW)
    gx, gy, maps = s.pmsm_lut
    assert gx.shape == (20,) and gy.shape == (35,) and not np.isnan(maps).any()
    # replicated border makes extrapolation constant
    assert np.array_equal(maps[:, 0, :], maps[:, 1, :]) and np.array_equal(maps[:, :, -1], maps[:, :, -2])
    with pytest.raises(ValueError, match="not allowed for saturated"):
        ex.PMSM(saturated=True)
    # independent re-derivation (oracle side) agrees
    from oracle import oracle as O

    ogx, ogy, omaps = O.load_motor_lut("SEW")
    assert np.array_equal(ogx, gx) and np.array_equal(ogy, gy) and np.array_equal(omaps, maps)


def test_in_axes_and_parameter_kinds(excenvs):
    ex = excenvs
    B = 4
    pn = {"theta": ex.MinMaxNormalization(-math.pi, math.pi),
          "omega": ex.MinMaxNormalization(torch.full((B,), -10.0), torch.full((B,), 10.0))}
    e = ex.Pendulum(batch_size=B, physical_normalizations=pn, static_params={"g": 9.81, "l": torch.ones(B) * 2, "m": 1})
    ax = e.in_axes_env_properties
    assert ax.physical_normalizations.omega.min == 0 and ax.physical_normalizations.theta.min is None
    assert ax.static_params.l == 0 and ax.static_params.g is None
    with pytest.raises(ValueError, match="but list is given"):
        ex.Pendulum(batch_size=B, static_params={"g": [9.81] * B, "l": 2, "m": 1})
    with pytest.raises(ValueError, match="unsupported solver"):
        ex.Pendulum(solver=object())
    assert ex.Pendulum(solver=ex.Tsit5())._solver_id == 1


def test_pytree_helpers_and_normalisation(excenvs):
    ex = excenvs
    e = ex.CartPole(batch_size=2)
    st = e._new_state([1.0, 2.0, 3.0, 4.0], float("nan"), None, False, [5.0, 6.0, 7.0, 8.0])
    leaves, struct = ex.tree_flatten(st)
    assert leaves[:4] == [1.0, 2.0, 3.0, 4.0] and leaves[5] is False and len(leaves) == 10  # solver_state None: no leaf
    assert ex.tree_unflatten(struct, leaves) == st
    st2 = e._new_state([1.0] * 4, 0.0, (1.0, 2.0, 3.0, 4.0), True, [0.0] * 4)
    assert ex.tree_structure(st) != ex.tree_structure(st2)
    mm = ex.MinMaxNormalization(-2.4, 2.4)
    assert mm.normalize(mm.denormalize(0.3)) == pytest.approx(0.3) and mm.denormalize(1.0) == 2.4
    n = e.normalize_state(e._new_state([2.4, 8.0, math.pi, -8.0], 0.0, None, False, [0.0] * 4), e.env_properties)
    assert n.physical_state.deflection == 1.0 and n.physical_state.omega == -1.0
    assert e.denormalize_state(n, e.env_properties).physical_state.theta == pytest.approx(math.pi)


def test_json_fixture_loader_round_trip(excenvs, tmp_path):
    ex = excenvs
    from helpers import GOLDEN

    params, an, pn, tau = ex.load_sim_properties_from_json(GOLDEN / "pmsm" / "sim_properties.json")
    assert params["deadtime"] == 1 and tau == 1e-4 and an["u_d"].max == pytest.approx(2 * 400 / 3)
    ex.dump_sim_properties_to_json(params, an, pn, tau, tmp_path / "p.json")
    p2, an2, pn2, tau2 = ex.load_sim_properties_from_json(tmp_path / "p.json")
    assert p2 == params and an2 == an and pn2 == pn and tau2 == tau
    env = ex.EnvironmentRegistry.PMSM.make(tau=tau, static_params=params, physical_normalizations=pn,
                                           action_normalizations=an)
    assert env.env_properties.static_params.u_dc == 400


def test_bench_clock_sampler_keeps_the_samples_of_the_timed_region(tmp_path):
    """bench.py: nvidia-smi is polled from before the warm-up; only samples whose own timestamps fall inside the timed
    region count for the reported clocks / throttle reasons; a region shorter than the polling period falls back to the
    nearest samples."""
    import datetime
    import importlib.util
    import sys
    import time
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("bench_under_test", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

    def sampler(rows):
        cs = bench.ClockSampler.__new__(bench.ClockSampler)
        f = tmp_path / f"smi{len(rows)}.csv"
        f.write_text("".join(rows))
        cs.f = open(f, "r+")
        cs.p = Done()
        return cs

    now = time.time()

    def row(dt, mhz, cap):
        ts = datetime.datetime.fromtimestamp(now + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return f"{ts}, {mhz}, 1965, 400.5, Not Active, Not Active, Not Active, {cap}\n"

    rows = [row(0.00, 1965, "Not Active"), row(0.05, 1935, "Not Active"), row(0.10, 1905, "Active"),
            row(0.15, 1965, "Not Active"), "garbage line\n"]
    out = sampler(rows).stop(now + 0.04, now + 0.11)
    assert out["samples"] == 2 and out["sm_mhz_min"] == 1905.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region" and out["power_w_max"] == 400.5
    out = sampler(rows[:2]).stop(now + 0.20, now + 0.21)  # nothing inside: nearest samples, said so
    assert out["samples"] == 2 and out["window"].startswith("nearest") and out["reasons"] == []


def test_plan_cache_is_keyed_by_parameter_values(excenvs, monkeypatch):
    """The plan cache must not confuse two EnvProperties objects that happen to share an id() after garbage collection,
    must notice reassigned / mutated properties and tau, and stays bounded (LRU)."""
    ex = excenvs
    env = ex.MassSpringDamper(batch_size=4, dtype=torch.float32, device="cpu")
    made = []

    class FakePlan:
        def __init__(self, props):
            self.k = props.static_params.k

    def fake_make(self, props, batch):
        made.append((props.static_params.k, batch, self.tau))
        return FakePlan(props)

    monkeypatch.setattr(type(env), "_make_plan", fake_make)
    p0 = env._plan()
    assert env._plan() is p0 and len(made) == 1
    # a parameter sweep building fresh EnvProperties per call: every distinct value gets its own plan, equal values
    # share one, whatever id() the temporaries have
    for k in (50.0, 60.0, 50.0, 70.0):
        props = env.EnvProperties(physical_normalizations=env.env_properties.physical_normalizations,
                                  action_normalizations=env.env_properties.action_normalizations,
                                  static_params=env.StaticParams(d=1, k=k, m=1))
        assert env._plan(props, 1).k == k
        del props
    assert [m[0] for m in made] == [100, 50.0, 60.0, 70.0]
    # in-place mutation of the environment's own properties and a new tau are noticed
    env.env_properties.static_params.k = 123.0
    assert env._plan().k == 123.0
    env.tau = 2e-4
    env._plan()
    assert made[-1] == (123.0, 4, 2e-4)
    # per-environment arrays: keyed by the array object
    arr = np.full(4, 9.0, np.float32)
    env.env_properties.static_params.k = arr
    a1 = env._plan()
    assert env._plan() is a1
    env.env_properties.static_params.k = arr.copy()
    assert env._plan() is not a1
    # bounded
    for k in range(40):
        env.env_properties.static_params.k = float(k)
        env._plan()
    assert len(env._plans) <= env._PLAN_CACHE_SIZE


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_truncation_thresholds_reproduce_the_reference_expression(excenvs, dtype):
    """The fused trajectory epilogue flags `abs(2 (x - min) / (max - min) - 1) > 1` by two threshold compares prepared on
    the host (exenv_trunc_thresholds).  They must reproduce the reference expression (IEEE NumPy, utils.py:13-14 +
    pendulum_env.py:381-385) for every float: checked on the neighbours of both thresholds, the range ends, zeros,
    infinities, NaN and random states, for default and awkward ranges."""
    import ctypes as C

    L = excenvs._lib
    T = dtype
    rng = np.random.default_rng(3)
    ranges = [(-math.pi, math.pi), (-10, 10), (0, 3), (-2.4, 2.4), (-8, 8), (-250, 0), (-0.37, 1.3), (1e-3, 7.0),
              (-41.0, -3.0), (5.0, 5.0000005), (-1e30, 1e30), (0.1, 0.3)]
    ranges += [tuple(sorted(rng.normal(0, 10 ** rng.uniform(-3, 3), 2))) for _ in range(40)]
    for lo, hi in ranges:
        lo_t, hi_t = T(lo), T(hi)
        if not hi_t > lo_t:
            continue
        a, b = C.c_double(), C.c_double()
        L.check(L.lib().exenv_trunc_thresholds(L.F64 if T == np.float64 else L.F32, float(lo_t), float(hi_t),
                                               C.byref(a), C.byref(b)))
        lo_last, hi_first = T(a.value), T(b.value)
        assert float(lo_last) == a.value and float(hi_first) == b.value
        pts = [lo_t, hi_t, T(0), -T(0), T(np.inf), -T(np.inf), T(np.nan)]
        for c in (lo_last, hi_first, lo_t, hi_t):
            x = c
            for _ in range(4):
                x = np.nextafter(x, T(np.inf))
                pts.append(x)
            x = c
            for _ in range(4):
                x = np.nextafter(x, -T(np.inf))
                pts.append(x)
            pts.append(c)
        x = np.concatenate([np.array(pts, T), rng.uniform(2 * lo - hi, 2 * hi - lo, 2000).astype(T)])
        with np.errstate(invalid="ignore", over="ignore"):
            ref = np.abs(T(2) * (x - lo_t) / (hi_t - lo_t) - T(1)) > T(1)
        got = (x >= hi_first) | (x <= lo_last)
        assert np.array_equal(ref, got), (lo, hi, x[ref != got][:4])
    a, b = C.c_double(), C.c_double()
    assert L.lib().exenv_trunc_thresholds(L.F32, 1.0, 1.0, C.byref(a), C.byref(b)) == -1
    assert L.lib().exenv_trunc_thresholds(L.F32, 2.0, 1.0, C.byref(a), C.byref(b)) == -1
