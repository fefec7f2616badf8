"""CPU-only: the kernels' __host__ __device__ math (tests/hostcheck, host instantiation of csrc/*.cuh) and the
host-side diffeqsolve schedule builder of libexenv_b200.so against the oracle.  The GPU parity tests proper are
test_step_gpu.py / test_sim_gpu.py; this file makes sure the arithmetic and the time bookkeeping they rely on are
right before GPU time is spent."""
import ctypes as C

import numpy as np
import pytest

from helpers import ENVS, golden_oracle, hc_lut, make_desc, ptr_array, random_state, rel_err
from oracle import oracle as O

TOL = {np.float64: 1e-12, np.float32: 2e-6}


def hc_step(lib, orc, state, action):
    d = make_desc(orc)
    lut, keep = hc_lut(orc)
    B = state[0].shape[0]
    out = [np.empty(B, orc.dtype) for _ in range(orc.D)]
    obs = np.empty((B, orc.desc.n_obs), orc.dtype)
    k = [np.empty(B, orc.dtype) for _ in range(orc.node)]
    action = np.ascontiguousarray(action, orc.dtype)
    lib.hc_step(C.byref(d), C.byref(lut) if lut else None, C.c_long(B), ptr_array(state),
                C.c_void_p(action.ctypes.data), ptr_array(out), C.c_void_p(obs.ctypes.data), ptr_array(k))
    return out, obs, k


@pytest.mark.parametrize("env", ENVS)
def test_device_math_replays_reference_golden_vectors(hostcheck_lib, env):
    """Euler step of the kernel math, fp64, replayed over the reference's own golden trajectory
    (reference tests/envs/<env>/test_<env>.py: test_step_results; tolerance rtol 1e-16 + atol 1e-8, PMSM rtol 1e-8)."""
    orc, g = golden_oracle(env)
    st = orc.reset(1, mode=2, values=g["observations"][0:1])["state"]
    n = min(2000, g["actions"].shape[0])
    gen = [g["observations"][0]]
    for i in range(n):
        st, obs, _ = hc_step(hostcheck_lib, orc, st, g["actions"][i:i + 1])
        gen.append(obs[0])
    gen = np.array(gen)
    ref = g["observations"][: n + 1]
    rtol = 1e-8 if env == "pmsm" else 1e-16
    assert np.allclose(gen, ref, rtol=rtol, atol=1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("solver", [O.EULER, O.TSIT5])
@pytest.mark.parametrize("env", ENVS + ["pmsm_brusa_sat", "pmsm_sew_sat"])
def test_device_math_step_matches_oracle(hostcheck_lib, env, solver, dtype):
    rng = np.random.default_rng(1234)
    if env.startswith("pmsm_"):
        orc = O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=True, motor=env.split("_")[1].upper())
    else:
        orc = O.Oracle(env, solver=solver, dtype=dtype)
    B = 257
    st = random_state(orc, B, rng)
    act = rng.uniform(-1, 1, (B, orc.A)).astype(dtype)
    ref = orc.step(st, act)
    out, obs, k = hc_step(hostcheck_lib, orc, st, act)
    if env.startswith("pmsm_") and solver == O.TSIT5 and dtype == np.float32:
        # Saturated PMSM, Tsit5 kernels: the right-hand side solves L x = b directly instead of forming inv(L) and three
        # products (exenv_envs.cuh).  At speed b = u - r i - omega J Psi cancels heavily, and the fp32 reference formula
        # itself sits up to 1e-4 (of the field's range) from the fp64 result after ONE step for some states; the kernel
        # form must be at least as close to the fp64 result as the reference formula is (within 3x: same order of rounding noise).
        ref64 = O.Oracle("pmsm", solver=solver, dtype=np.float64, saturated=True, motor=orc.motor).step(
            [x.astype(np.float64) for x in st], act.astype(np.float64))
        from helpers import field_scales, scaled_err
        for a, b, b64, sc in zip(out, ref["state"], ref64["state"], field_scales(orc)):
            noise = scaled_err(b, b64, sc)
            assert scaled_err(a, b64, sc) <= max(2e-6, 3 * noise), (scaled_err(a, b64, sc), noise)
        return
    for a, b in zip(out, ref["state"]):
        assert rel_err(a, b, floor=1e-3) < TOL[dtype] * 50
    assert rel_err(obs, ref["obs"], floor=1e-3) < TOL[dtype] * 50
    if solver == O.TSIT5:
        for a, b in zip(k, ref["k"]):
            assert rel_err(a, b, floor=1.0) < TOL[dtype] * 1e3


def hc_sim(lib, orc, state, actions, obs_dt, act_dt, tile_steps=32):
    import exciting_environments_b200 as ex

    d = make_desc(orc)
    lut, keep = hc_lut(orc)
    B, T = actions.shape[0], actions.shape[1]
    info, words, dt, theta, tiles = ex._lib.schedule(d, T, obs_dt, act_dt, tile_steps)
    nst = 8 if orc.env == "pmsm" else orc.node
    traj = [np.zeros((B, info.n_save), orc.dtype) for _ in range(nst)]
    actions = np.ascontiguousarray(actions, orc.dtype)
    lib.hc_sim(C.byref(d), C.byref(lut) if lut else None, C.c_long(B), C.c_long(T), C.c_long(info.n_save),
               C.c_long(info.n_steps), C.c_void_p(words.ctypes.data), C.c_void_p(dt.ctypes.data),
               C.c_void_p(theta.ctypes.data), ptr_array(state), C.c_void_p(actions.ctypes.data), ptr_array(traj))
    return info, traj, (words, dt, theta, tiles)


CASES = [
    ("pendulum", 1e-4, 1e-4, 200), ("pendulum", 2e-2, 2e-2, 200), ("mass_spring_damper", 1e-4, 1e-4, 300),
    ("fluid_tank", 1e-3, 1e-3, 150), ("cart_pole", 2e-2, 2e-2, 100), ("acrobot", 1e-3, 1e-3, 100),
    ("pmsm", 1e-4, 1e-4, 120), ("pmsm_brusa_sat", 1e-4, 1e-4, 80), ("mass_spring_damper", 5e-5, 1e-4, 64),
    ("pendulum", 1e-4, 3e-4, 40),
]


@pytest.mark.parametrize("time_rule", [0, 1])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("solver", [O.EULER, O.TSIT5])
@pytest.mark.parametrize("env,obs_dt,act_dt,T", CASES)
def test_schedule_and_sim_core_match_oracle(hostcheck_lib, env, obs_dt, act_dt, T, solver, dtype, time_rule):
    """The product's host-built schedule (exenv_schedule_build) driving the kernel's per-step function reproduces the
    oracle's independent, inline emulation of the diffeqsolve loop, including ZOH lag and fp32 time drift."""
    rng = np.random.default_rng(99)
    if env == "pmsm_brusa_sat":
        orc = O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=True, motor="BRUSA", time_rule=time_rule)
    else:
        orc = O.Oracle(env, solver=solver, dtype=dtype, time_rule=time_rule, tau=min(obs_dt, act_dt) if env != "pmsm" else obs_dt)
    B = 33
    st = random_state(orc, B, rng, spread=0.5)
    acts = rng.uniform(-1, 1, (B, T, orc.A)).astype(dtype)
    ref = orc.sim_ahead(st, acts, obs_dt, act_dt)
    info, traj, _ = hc_sim(hostcheck_lib, orc, st, acts, obs_dt, act_dt)
    assert info.n_steps == ref["n_steps"]
    assert info.n_save == ref["obs"].shape[1]
    if orc.env == "pmsm":
        pairs = [(traj[0], ref["traj"][2]), (traj[1], ref["traj"][3]), (traj[2], ref["traj"][4]),
                 (traj[3], ref["traj"][5]), (traj[4], ref["traj"][0]), (traj[5], ref["traj"][1])]
    else:
        pairs = list(zip(traj, ref["traj"]))
    # chaotic envs amplify rounding differences along the horizon; tolerances are per-trajectory here, the per-step
    # bound is checked in test_device_math_step_matches_oracle
    tol = (1e-9 if dtype == np.float64 else 2e-3)
    if env == "pmsm_brusa_sat":
        # full-range random voltages drive the currents thousands of amperes outside the flux maps, where the dynamics
        # amplify any rounding difference (the kernels' derivative-form Tsit5 stage sums round differently from the
        # oracle's increment form): bound = 20 x the oracle's own sensitivity to a one-ulp change of the initial state
        eps = np.finfo(dtype).eps
        pert = orc.sim_ahead([(x * (1 + eps)).astype(dtype) for x in st], acts, obs_dt, act_dt)
        sens = max(np.max(np.abs(a - b) / (1.0 + np.abs(b))) for a, b in zip(pert["traj"], ref["traj"]))
        tol = max(tol, 20 * float(sens))
    for a, b in pairs:
        assert np.max(np.abs(a - b) / (1.0 + np.abs(b))) < tol


def test_schedule_exposes_zoh_lag_and_fp32_drift():
    """with accumulated time the truncating ZOH index lags for many steps."""
    import exciting_environments_b200 as ex

    orc = O.Oracle("pendulum", solver=O.EULER, dtype=np.float32)
    info, words, dt, theta, tiles = ex._lib.schedule(make_desc(orc), 2000, 1e-4, 1e-4, 32)
    base = (words & np.uint64(0xFFFFFFFF)).astype(np.int64)
    lag = np.sum(base != np.arange(info.n_steps).clip(max=1999))
    assert lag > 100          # the index is not simply the step number in fp32
    assert info.n_save == 2001
    assert np.all((theta >= 0) & (theta <= 1))
    assert tiles[:, 1].sum() == info.n_steps and tiles[:, 5].sum() == info.n_save


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fast_fmod_is_exact(hostcheck_lib, dtype):
    """exenv_math.cuh fmod_pos / wrap_pi: bit-identical to C fmod (numpy.fmod) and to the reference's
    ((theta + pi) % (2 pi)) - pi, including multiples of 2 pi, signed zeros, huge arguments, inf and nan."""
    rng = np.random.default_rng(0)
    two_pi = dtype(2 * np.pi)
    k = np.arange(-40, 41).astype(dtype)
    edge = np.concatenate([k * two_pi, np.nextafter(k * two_pi, dtype(np.inf)), np.nextafter(k * two_pi, dtype(-np.inf)),
                           np.array([0.0, -0.0, 1e-30, -1e-30, 3e7, -3e7, 1e20, -1e20, np.inf, -np.inf, np.nan], dtype)])
    x = np.concatenate([edge, rng.uniform(-50, 50, 200000).astype(dtype), rng.uniform(-7, 7, 200000).astype(dtype),
                        (rng.standard_normal(50000) * 1e5).astype(dtype)])
    f, w = np.empty_like(x), np.empty_like(x)
    hostcheck_lib.hc_fmod_two_pi(C.c_long(x.size), C.c_int(dtype == np.float64), C.c_void_p(x.ctypes.data),
                                 C.c_void_p(f.ctypes.data), C.c_void_p(w.ctypes.data))
    with np.errstate(invalid="ignore"):
        ref = np.fmod(x, two_pi)
        pi = dtype(np.pi)
        wref = np.mod(x + pi, two_pi) - pi
    u = np.uint32 if dtype == np.float32 else np.uint64
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(f), nan) and np.array_equal(f[~nan].view(u), ref[~nan].view(u))
    nanw = np.isnan(wref)
    assert np.array_equal(np.isnan(w), nanw) and np.array_equal(w[~nanw], wref[~nanw])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_normalisation_forms(hostcheck_lib, dtype):
    """exenv_math.cuh: norm_inv<true> (prepared reciprocal + one correction; the form the truncation flags are evaluated
    on) is bit-identical to the reference's 2 (x - min) / (max - min) - 1 with an IEEE division, flags included; norm_fma
    (one FMA, what the time-major kernels return) stays within 3 ulp of 1 of it, its offset is exactly 0 / -1 for
    symmetric / zero-based ranges, and it is exact at 0 and at the lower edge of a zero-based range."""
    rng = np.random.default_rng(5)
    eps = np.finfo(dtype).eps
    for mn, mx in [(-np.pi, np.pi), (-10, 10), (-15, 15), (0, 3), (0, 0.2), (-400, 0), (-0.37, 0.37), (-41, 41),
                   (-3.3, 7.9), (1.5, 2.5)]:
        lo, hi = dtype(mn), dtype(mx)
        edge = np.array([lo, hi, np.nextafter(lo, dtype(-np.inf)), np.nextafter(hi, dtype(np.inf)), 0.0, -0.0,
                         lo * dtype(1.001) if lo != 0 else dtype(-1e-3), hi * dtype(1.001)], dtype)
        x = np.concatenate([edge, rng.uniform(float(lo) - 0.1 * float(hi - lo), float(hi) + 0.1 * float(hi - lo),
                                              100000).astype(dtype)])
        o = [np.empty_like(x) for _ in range(3)]
        coef = np.empty(2, dtype)
        hostcheck_lib.hc_norm_forms(C.c_long(x.size), C.c_int(dtype == np.float64), C.c_void_p(x.ctypes.data),
                                    C.c_double(float(lo)), C.c_double(float(hi)), *[C.c_void_p(a.ctypes.data) for a in o],
                                    C.c_void_p(coef.ctypes.data))
        ref = dtype(2) * (x - lo) / (hi - lo) - dtype(1)
        assert ref.dtype == dtype
        assert np.array_equal(o[0], ref)
        assert np.array_equal(o[1], ref), (mn, mx)                      # correctly rounded: same value, same flags
        assert np.array_equal(np.abs(o[1]) > 1, np.abs(ref) > 1)
        assert np.max(np.abs(o[2] - ref)) <= 3 * eps * max(1.0, float(np.max(np.abs(ref)))), (mn, mx)
        if mn == -mx:
            assert coef[1] == 0 and o[2][4] == 0
        if mn == 0:
            assert coef[1] == -1 and o[2][0] == -1


def test_packed_sincos_accuracy(hostcheck_lib):
    """exenv_math.cuh: sincos2 -- the packed (two angles per FFMA2 chain) sine / cosine of the trajectory kernels of the
    pendulum, cart-pole and acrobot.  The host build performs the same correctly rounded FMAs as the device's FFMA2,
    so this is the device's result: at most 2 ulp from the exact value for |x| < 1e5 (CUDA's sinf / cosf: 1.5 ulp),
    including arguments within 1e-4 of the zeros (multiples of pi / 2); huge arguments, inf and nan go to libm."""
    rng = np.random.default_rng(0)
    parts = [rng.uniform(-np.pi, np.pi, 400000), rng.uniform(-100, 100, 400000), rng.uniform(-1e5, 1e5, 400000),
             (np.arange(-2000, 2001)[:, None] * np.pi / 2 + rng.uniform(-1e-3, 1e-3, (4001, 20))).ravel(),
             (np.arange(-40, 41)[:, None] * np.pi / 2 + rng.uniform(-1e-4, 1e-4, (81, 500))).ravel(),
             np.array([0.0, -0.0, 1e-30, -1e-30, 1e-8, 0.5, -0.5])]
    x = np.concatenate(parts).astype(np.float32)
    if x.size % 2:
        x = np.concatenate([x, x[:1]])
    s, c = np.empty_like(x), np.empty_like(x)
    hostcheck_lib.hc_sincos2(C.c_long(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(s.ctypes.data),
                             C.c_void_p(c.ctypes.data))
    xs = x.astype(np.float64)
    for got, want in ((s, np.sin(xs)), (c, np.cos(xs))):
        ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
        err = np.abs(got.astype(np.float64) - want) / ulp
        assert err.max() <= 2.0, (err.max(), x[err.argmax()])
    big = np.array([2e5, -3e6, 1e9, np.inf, -np.inf, np.nan, 7e7, 1.5e5], np.float32)
    sb, cb = np.empty_like(big), np.empty_like(big)
    hostcheck_lib.hc_sincos2(C.c_long(big.size), C.c_void_p(big.ctypes.data), C.c_void_p(sb.ctypes.data),
                             C.c_void_p(cb.ctypes.data))
    with np.errstate(invalid="ignore"):
        assert np.allclose(sb, np.sin(big.astype(np.float64)), rtol=0, atol=2e-7, equal_nan=True)
        assert np.allclose(cb, np.cos(big.astype(np.float64)), rtol=0, atol=2e-7, equal_nan=True)


@pytest.mark.parametrize("motor", ["BRUSA", "SEW"])
def test_flux_map_coefficient_form(hostcheck_lib, motor):
    """exenv_envs.cuh: lut_eval_coef / Pmsm<float, 2> -- the coefficient form of the saturated-PMSM flux maps that the
    fp32 time-major trajectory kernel reads (one 16-byte read per map and cell).  Against the bilinear interpolant in
    float64 (RegularGridInterpolator semantics incl. linear extrapolation outside the grid, pmsm_env.py:352-361) it is
    as close as the four-corner form of the reference evaluated in float32, and so is the right-hand side built on it
    (pmsm_env.py:487-507)."""
    from scipy.interpolate import RegularGridInterpolator

    gx, gy, maps = O.load_motor_lut(motor)
    nx, ny = gx.size, gy.size
    rng = np.random.default_rng(5)
    n = 200000
    lo = np.array([gx[0], gy[0]]) - 2 * (gx[1] - gx[0])
    hi = np.array([gx[-1], gy[-1]]) + 2 * (gx[1] - gx[0])
    pts = rng.uniform(lo, hi, (n, 2))
    # grid nodes, points a rounding error away from grid lines and far outside
    k = 2000
    pts[:k, 0] = rng.choice(gx, k) + rng.choice([0.0, 1e-5, -1e-5], k)
    pts[k:2 * k, 1] = rng.choice(gy, k) + rng.choice([0.0, 1e-5, -1e-5], k)
    pts[2 * k:2 * k + 100] *= 3.0
    idq = pts.astype(np.float32)
    i_d, i_q = np.ascontiguousarray(idq[:, 0]), np.ascontiguousarray(idq[:, 1])
    orc = O.Oracle("pmsm", solver=O.TSIT5, dtype=np.float32, saturated=True, motor=motor)
    params = np.ascontiguousarray(np.array([orc.desc.p[i] for i in range(32)], np.float32))
    u = rng.uniform(-200, 200, (n, 2)).astype(np.float32)
    om = rng.uniform(-3000, 3000, n).astype(np.float32)
    node, coef = np.empty((n, 6), np.float32), np.empty((n, 6), np.float32)
    dyn, dyc = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    mp = np.ascontiguousarray(maps, np.float64)
    P = lambda a: C.c_void_p(a.ctypes.data)
    hostcheck_lib.hc_lut_forms.restype = C.c_int
    exact = hostcheck_lib.hc_lut_forms(C.c_long(n), P(np.ascontiguousarray(gx)), P(np.ascontiguousarray(gy)), P(mp),
                                       C.c_int(nx), C.c_int(ny), P(i_d), P(i_q), P(u), P(om), P(params), P(node), P(coef),
                                       P(dyn), P(dyc))
    assert exact == 1  # the shipped motors' grids are integer-valued and equidistant
    x64 = idq.astype(np.float64)
    truth = np.stack([RegularGridInterpolator((gx, gy), maps[q], method="linear", bounds_error=False, fill_value=None)(x64)
                      for q in range(6)], axis=1)
    scale = np.abs(maps).reshape(6, -1).max(axis=1)
    inside = (np.abs(x64[:, 0] - (gx[0] + gx[-1]) / 2) <= (gx[-1] - gx[0]) / 2) & \
             (np.abs(x64[:, 1] - (gy[0] + gy[-1]) / 2) <= (gy[-1] - gy[0]) / 2)
    e_node = np.abs(node - truth) / scale
    e_coef = np.abs(coef - truth) / scale
    # inside the grid: a few float32 ulp of the map's range (random points: 1.4e-7 against 2.1e-7 for the four-corner
    # form; worst at the far edge of a cell of the small cross-inductance maps, where the increments cancel: 6e-7);
    # outside the extrapolated value grows with the distance
    assert e_coef[inside].max() <= 1e-6, e_coef[inside].max()
    assert e_coef[inside].max() <= 4 * e_node[inside].max() + 1e-7
    assert e_coef[inside].mean() <= e_node[inside].mean()
    grow = 1.0 + np.abs(truth) / scale
    assert (e_coef / grow).max() <= 1e-6  # (the four-corner weights cancel far outside: up to 1e-4 there)
    # right-hand side in float64 from the float64 look-up: inv(L) (u - r_s i - omega J Psi)
    r_s = float(params[1])
    L = truth[:, :4].reshape(n, 2, 2)
    psi = truth[:, 4:6]
    b = u.astype(np.float64) - r_s * x64 - om[:, None].astype(np.float64) * np.stack([-psi[:, 1], psi[:, 0]], axis=1)
    want = np.linalg.solve(L, b[:, :, None])[:, :, 0]
    # error scale: the terms of the right-hand side before they cancel, through the inverse
    terms = np.abs(u.astype(np.float64)) + r_s * np.abs(x64) + np.abs(om[:, None].astype(np.float64)) * np.abs(psi[:, ::-1])
    scl = (np.linalg.norm(np.linalg.inv(L), 2, axis=(1, 2)) * terms.max(axis=1))[:, None]
    en = (np.abs(dyn[:, :2] - want) / scl)[inside]
    ec = (np.abs(dyc[:, :2] - want) / scl)[inside]
    print("rhs error / scale: node", en.max(), "coef", ec.max())
    assert ec.max() <= 1e-6, ec.max()
    assert ec.max() <= 4 * en.max() + 1e-7, (ec.max(), en.max())
    assert np.array_equal(dyc[:, 2], om) and np.array_equal(dyn[:, 2], om)


def test_coefficient_form_is_only_built_for_exactly_representable_grids(hostcheck_lib):
    """exenv_envs.cuh: lut_grid_exact_f32 -- the coefficient-form kernel rebuilds cell origins as g0 + i * d in float32,
    so plans get that table only for grids where this is exact; everything else keeps the node form with its per-cell
    origin / width reads."""
    f = hostcheck_lib.hc_lut_grid_exact
    f.restype = C.c_int

    def ok(g):
        g = np.ascontiguousarray(g, np.float64)
        return f(C.c_void_p(g.ctypes.data), C.c_int(g.size)) == 1

    assert ok(np.arange(-260.0, 261.0, 10.0)) and ok(np.arange(-17.0, 18.0, 1.0)) and ok(np.arange(0, 8) * 0.25 - 3.5)
    assert not ok(np.arange(0, 50) * 0.1)                      # 0.1 is not a float32 number
    assert not ok(np.array([0.0, 1.0, 2.0, 3.5, 4.0]))         # not equidistant
    assert not ok(np.array([0.0, 1.0, 2.0, 3.0 + 1e-9]))       # equidistant only up to rounding
    assert not ok(np.arange(0, 40) * 1.0 + 2.0 ** 25)          # nodes beyond the float32 integer range
    assert not ok(np.array([3.0])) and not ok(np.array([2.0, 1.0, 0.0]))
