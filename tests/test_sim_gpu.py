"""`-m gpu`: vmap_sim_ahead (persistent multi-step kernel) through the C ABI against the oracle."""
import numpy as np
import pytest
import torch

from gpu_helpers import make_env, np_, to_state
from helpers import ENVS, field_scales, flags_from_states, random_state, traj_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _orc(env, solver, dtype, **kw):
    if env.startswith("pmsm_"):
        return O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=True, motor=env.split("_")[1].upper(), **kw)
    return O.Oracle(env, solver=solver, dtype=dtype, **kw)


def _cmp(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / (1.0 + np.abs(b))) < tol


# Trajectory tolerances: 1e-4 relative (fp32) over the horizon as the north star states for the non-chaotic envs,
# far tighter in fp64; errors are normalised per environment by max(normalisation half-span, max_t |x|)
# (helpers.traj_err).  Rounding differences (libm, FMA contraction) are amplified by chaotic dynamics (cart-pole,
# acrobot, pendulum at coarse tau) and by ill-conditioned ones (Euler at |1 + h*lambda| > 1 for the PMSM at speed;
# saturated PMSM far outside its flux maps), so the bound is widened to 20x the trajectory's own sensitivity to a
# one-ulp perturbation of the initial state, measured with the oracle.
HORIZON_TOL = {np.float64: 1e-9, np.float32: 1e-4}


def horizon_tol(orc, st, acts, dtype, ref, scales):
    eps = np.finfo(dtype).eps
    st2 = [(x * (1 + eps)).astype(dtype) for x in st]
    pert = orc.sim_ahead(st2, acts, orc.tau, orc.tau)
    sens = max(traj_err(p, r, s) for p, r, s in zip(pert["traj"], ref["traj"], scales))
    return max(HORIZON_TOL[dtype], 20 * sens), sens


_ORACLE_CACHE = {}


def _oracle_case(env, solver, dtype):
    """Seeded inputs + oracle trajectories + tolerance of one (env, solver, dtype) case (shared by both layouts)."""
    key = (env, solver, dtype)
    if key not in _ORACLE_CACHE:
        rng = np.random.default_rng(42)
        orc = _orc(env, solver, dtype)
        B, T = 300, 200  # 3 CTAs, the last one ragged
        gentle = orc.env == "pmsm"
        st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=gentle)
        acts = rng.uniform(-1, 1, (B, T, orc.A)).astype(dtype)
        if gentle:
            acts *= 0.02
        ref = orc.sim_ahead(st, acts, orc.tau, orc.tau, epilogue=True)
        scales = field_scales(orc)
        tol, sens = horizon_tol(orc, st, acts, dtype, ref, scales)
        _ORACLE_CACHE.clear()  # keep one case alive: the two layouts of a case run back to back
        _ORACLE_CACHE[key] = (orc, B, T, gentle, st, acts, ref, scales, tol, sens)
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("solver", [O.EULER, O.TSIT5])
@pytest.mark.parametrize("env", ENVS + ["pmsm_brusa_sat", "pmsm_sew_sat"])
@pytest.mark.parametrize("layout,epilogue", [("batch_major", True), ("time_major", True), ("time_major", False)])
def test_sim_ahead_matches_oracle(excenvs, env, solver, dtype, layout, epilogue):
    """Both memory layouts (and, time-major, both the plain and the fused-epilogue instantiation) against the oracle."""
    orc, B, T, gentle, st, acts, ref, scales, tol, sens = _oracle_case(env, solver, dtype)
    e = make_env(excenvs, orc, B)
    a_dev = torch.as_tensor(acts, device="cuda")
    if epilogue:
        obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(
            to_state(e, st), a_dev, orc.tau, orc.tau, layout=layout)
    else:
        obs, states, last = e.vmap_sim_ahead(to_state(e, st), a_dev, orc.tau, orc.tau, layout=layout)
    if layout == "time_major":  # same logical shapes, batch axis innermost in memory
        assert obs.permute(1, 0, 2).is_contiguous()
        assert getattr(states.physical_state, e.PHYS_FIELDS[0]).t().is_contiguous()
    if gentle or orc.env in ("mass_spring_damper", "fluid_tank"):
        assert tol == HORIZON_TOL[dtype], ("scenario is unexpectedly ill-conditioned", sens)
    assert obs.shape == ref["obs"].shape
    for name, b, s_ in zip(e.PHYS_FIELDS, ref["traj"], scales):
        err = traj_err(np_(getattr(states.physical_state, name)), b, s_)
        assert err <= tol, (name, err, tol)
    ob, ob_ref = np_(obs), ref["obs"]
    for o in range(orc.O):
        err = traj_err(ob[..., o], ob_ref[..., o], 1.0)
        assert err <= tol, ("obs", o, err, tol)
    for name, b, s_ in zip(e.PHYS_FIELDS, ref["last"], scales):
        assert np.array_equal(np_(getattr(last.physical_state, name)), np_(getattr(states.physical_state, name))[:, -1])
    if solver == O.TSIT5:
        ode = ("i_d", "i_q", "epsilon") if orc.env == "pmsm" else e.PHYS_FIELDS
        for a, b, n in zip(last.additions.solver_state, ref["k_last"], ode):
            s_ = scales[e.PHYS_FIELDS.index(n)]
            err = traj_err((np_(a) * orc.tau)[:, None], (np.asarray(b, np.float64) * orc.tau)[:, None], s_)
            assert err <= max(tol, 1e-5 if dtype == np.float32 else 1e-10) * 10, ("k_last", n, err)
    if not epilogue:
        return
    # flags: bit exact -- the reference expression (IEEE NumPy, helpers.flags_from_states) applied to the states this very
    # call returned, every environment, every saved sample; terminated as the reference defines it from them
    tr = np_(trunc).astype(np.uint8)
    own = [np_(getattr(states.physical_state, n)) for n in e.PHYS_FIELDS]
    tr_def = flags_from_states(orc, own)
    assert tr.shape == tr_def.shape and np.array_equal(tr, tr_def), (int(np.sum(tr != tr_def)), tr.size)
    te = np_(term)[..., 0].astype(np.uint8)
    if orc.env == "pmsm":
        assert np.array_equal(te, tr[:, 1:, 0])  # pmsm_env.py:981-983
    elif orc.env == "fluid_tank":
        assert not te.any()  # fluid_tank_env.py:328-331
    else:
        assert np.array_equal(te, (np_(rew)[..., 0] == 0).astype(np.uint8))  # reward == 0 (pendulum_env.py:387-390)
    # and against the oracle's own flags: they can only differ where the oracle's state sits within the trajectory
    # error of the threshold (the two state trajectories differ by rounding)
    tr_ref = ref["truncated"]
    if orc.n_trunc == orc.O:
        diff = tr != tr_ref
        assert not np.any(diff & (np.abs(np.abs(ref["obs"]) - 1.0) > 100 * tol))
    assert np.array_equal(np_(rew)[..., 0], ref["reward"])
    assert rew.shape == (B, T, 1) and term.shape == (B, T, 1) and trunc.shape == (B, T + 1, orc.n_trunc)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("env", ["pmsm", "pmsm_brusa_sat"])
def test_pmsm_open_loop_full_range_actions(excenvs, env, dtype):
    """Full-range random voltages drive the currents far outside the permitted box (u / r_s ~ 15 kA); Euler at speed is
    outside its stability region (|1 + h*lambda| ~ 1.03).  Compared with the sensitivity-scaled bound only."""
    rng = np.random.default_rng(42)
    for solver in (O.EULER, O.TSIT5):
        orc = _orc(env, solver, dtype)
        B, T = 300, 200
        st = random_state(orc, B, rng, spread=0.5)
        acts = rng.uniform(-1, 1, (B, T, 2)).astype(dtype)
        ref = orc.sim_ahead(st, acts, orc.tau, orc.tau)
        e = make_env(excenvs, orc, B)
        obs, states, last = e.vmap_sim_ahead(to_state(e, st), torch.as_tensor(acts, device="cuda"), orc.tau, orc.tau)
        scales = field_scales(orc)
        tol, sens = horizon_tol(orc, st, acts, dtype, ref, scales)
        for name, b, s_ in zip(e.PHYS_FIELDS, ref["traj"], scales):
            err = traj_err(np_(getattr(states.physical_state, name)), b, s_)
            assert err <= tol, (name, err, tol)


@pytest.mark.parametrize("tile_steps,block_envs", [(1, 32), (7, 64), (16, 128), (64, 128), (32, 256)])
def test_tile_geometry_does_not_change_results(excenvs, tile_steps, block_envs):
    """Any (tile_steps, block_envs) staging must give bit-identical trajectories."""
    rng = np.random.default_rng(5)
    orc = O.Oracle("pendulum", solver=O.TSIT5, dtype=np.float32)
    B, T = 777, 333
    st = random_state(orc, B, rng)
    acts = torch.as_tensor(rng.uniform(-1, 1, (B, T, 1)).astype(np.float32), device="cuda")
    e = make_env(excenvs, orc, B)
    base = e.vmap_sim_ahead(to_state(e, st), acts, 1e-4, 1e-4)
    other = e.vmap_sim_ahead(to_state(e, st), acts, 1e-4, 1e-4, tile_steps=tile_steps, block_envs=block_envs)
    assert torch.equal(base[0], other[0])
    assert torch.equal(base[1].physical_state.theta, other[1].physical_state.theta)
    assert torch.equal(base[2].physical_state.omega, other[2].physical_state.omega)


@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("obs_dt,act_dt,T", [(5e-5, 1e-4, 64), (1e-4, 3e-4, 50), (2e-2, 2e-2, 300), (1e-3, 1e-3, 1),
                                             (1e-4, 1e-4, 3000)])
def test_zoh_and_save_grid_quirks(excenvs, obs_dt, act_dt, T, dtype, layout):
    """obs_stepsize < action_stepsize, coarse tau (ZOH lag), single action; time-varying actions."""
    rng = np.random.default_rng(11)
    orc = O.Oracle("mass_spring_damper", solver=O.TSIT5, dtype=dtype, tau=obs_dt)
    B = 130
    st = random_state(orc, B, rng)
    acts = rng.uniform(-1, 1, (B, T, 1)).astype(dtype)
    ref = orc.sim_ahead(st, acts, obs_dt, act_dt)
    e = make_env(excenvs, orc, B)
    obs, states, last = e.vmap_sim_ahead(to_state(e, st), torch.as_tensor(acts, device="cuda"), obs_dt, act_dt,
                                         layout=layout)
    assert obs.shape == ref["obs"].shape
    assert e.sim_ahead_shape(T, obs_dt, act_dt) == (ref["obs"].shape[1], ref["n_steps"])
    assert _cmp(np_(obs), ref["obs"], HORIZON_TOL[dtype])


@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
def test_reference_tracking_per_env_params_and_epilogue(excenvs, layout):
    rng = np.random.default_rng(8)
    B, T = 200, 60
    base = O.Oracle("cart_pole", dtype=np.float32)
    static = {k: (np.float64(v) * rng.uniform(0.9, 1.1, B)).astype(np.float32) for k, v in base.static.items()}
    orc = O.Oracle("cart_pole", solver=O.TSIT5, dtype=np.float32, static_params=static, tau=1e-3,
                   control_state=["theta", "velocity"])
    st, refl = random_state(base, B, rng, 0.4), random_state(base, B, rng, 0.4)
    acts = rng.uniform(-1, 1, (B, T, 1)).astype(np.float32)
    ref = orc.sim_ahead(st, acts, 1e-3, 1e-3, ref=refl, epilogue=True)
    e = make_env(excenvs, orc, B)
    obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(
        to_state(e, st, refl), torch.as_tensor(acts, device="cuda"), 1e-3, 1e-3, layout=layout)
    assert obs.shape == (B, T + 1, 6)
    assert _cmp(np_(obs), ref["obs"], 1e-4)
    assert _cmp(np_(rew)[..., 0], ref["reward"], 1e-3)
    assert np.mean(np_(term)[..., 0].astype(np.uint8) == ref["terminated"]) == 1.0
    # truncated incl. the reference entries: the reference expression on the returned states, bit for bit
    own = [np_(getattr(states.physical_state, n)) for n in e.PHYS_FIELDS]
    assert np.array_equal(np_(trunc).astype(np.uint8), flags_from_states(orc, own, refl))
    assert np.array_equal(np_(term)[..., 0], np_(rew)[..., 0] == 0)


def test_max_steps_mirrors_diffrax(excenvs):
    """diffeqsolve's default max_steps=4096: > 4096 solver steps raise unless lifted."""
    e = excenvs.MassSpringDamper(batch_size=32, dtype=torch.float32)
    _, s = e.vmap_reset()
    acts = torch.zeros((32, 5000, 1), device="cuda")
    with pytest.raises(excenvs.ExenvError, match="maximum number of solver steps"):
        e.vmap_sim_ahead(s, acts, e.tau, e.tau)
    e.max_steps = None
    obs, _, _ = e.vmap_sim_ahead(s, acts, e.tau, e.tau)
    assert obs.shape == (32, 5001, 2)


def test_euler_sim_ahead_equals_repeated_step(excenvs):
    """reference tests/envs/test_core_functions.py:134-155 for every env (10 steps, all-ones actions, fp64)."""
    for reg in excenvs.EnvironmentRegistry:
        e = reg.make(batch_size=4, solver=excenvs.Euler(), dtype=torch.float64)
        obs, state = e.vmap_reset()
        acts = torch.ones((4, 10, e.action_dim), dtype=torch.float64, device="cuda")
        obs_a, states_a, last_a = e.vmap_sim_ahead(state, acts, e.tau, e.tau)
        last_obs = e.vmap_generate_observation(last_a)
        for i in range(10):
            obs_s, state = e.vmap_step(state, acts[:, i])
        assert np.allclose(np_(last_obs), np_(obs_s), rtol=1e-16, atol=1e-8), reg
        assert np.allclose(np_(obs_a[:, -1]), np_(obs_s), rtol=1e-16, atol=1e-8), reg


@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
def test_full_size_linearity_property(excenvs, layout):
    """BASELINE config 2 scale check with a size-independent property: the mass-spring-damper is linear, so the
    trajectory of (c*x0, c*u) is c times the trajectory of (x0, u) -- checked on 2^17 envs x 2000 Tsit5 steps."""
    B, T = 1 << 17, 2000
    e = excenvs.MassSpringDamper(batch_size=B, solver=excenvs.Tsit5(), dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = [torch.rand(B, device="cuda", generator=g) - 0.5 for _ in range(2)]
    acts = torch.rand((B, T, 1), device="cuda", generator=g) * 0.5 - 0.25
    s1 = e._new_state(x0, None, None, None, [None, None])
    s2 = e._new_state([2 * x for x in x0], None, None, None, [None, None])
    nan = torch.full((B,), float("nan"), device="cuda")
    for s in (s1, s2):
        s.PRNGKey = nan
        s.reference = e.PhysicalState(nan, nan)
    kw = dict(layout=layout, block_envs=4) if layout == "time_major" else {}  # 4 envs / thread: the packed FFMA2 kernel
    _, st1, _ = e.vmap_sim_ahead(s1, acts, 1e-4, 1e-4, **kw)
    # action normalisation is affine with zero offset (min = -max), so scaling the normalised action scales the force
    _, st2, _ = e.vmap_sim_ahead(s2, 2 * acts, 1e-4, 1e-4, **kw)
    d = (st2.physical_state.deflection - 2 * st1.physical_state.deflection).abs().max().item()
    v = (st2.physical_state.velocity - 2 * st1.physical_state.velocity).abs().max().item()
    assert d < 1e-5 and v < 1e-4


def test_host_buffer_entry_point_equals_device_entry_point(excenvs):
    """exenv_sim_ahead_host (H2D / kernel / D2H pipeline over chunks) == exenv_sim_ahead."""
    import ctypes as C

    L = excenvs._lib
    B, T = 1000, 120
    e = excenvs.Pendulum(batch_size=B, solver=excenvs.Tsit5(), dtype=torch.float32)
    _, s = e.vmap_reset(rng=1)
    acts = torch.rand((B, T, 1), device="cuda") * 2 - 1
    obs, states, last = e.vmap_sim_ahead(s, acts, e.tau, e.tau)
    h_state = [getattr(s.physical_state, n).cpu().numpy().copy() for n in e.PHYS_FIELDS]
    h_acts = acts.cpu().numpy().copy()
    h_obs = np.zeros((B, T + 1, 2), np.float32)
    h_traj = [np.zeros((B, T + 1), np.float32) for _ in range(2)]
    h_last = [np.zeros(B, np.float32) for _ in range(2)]
    a = L.ExenvSimArgs()
    a.batch, a.n_actions, a.obs_stepsize, a.action_stepsize, a.max_steps = B, T, e.tau, e.tau, 4096
    for i in range(2):
        a.state_in[i] = h_state[i].ctypes.data
        a.traj[i] = h_traj[i].ctypes.data
        a.last[i] = h_last[i].ctypes.data
    a.actions, a.obs = h_acts.ctypes.data, h_obs.ctypes.data
    L.check(L.lib().exenv_sim_ahead_host(e._plan().handle, C.byref(a), 256))
    assert np.array_equal(h_obs, np_(obs))
    assert np.array_equal(h_traj[0], np_(states.physical_state.theta))
    assert np.array_equal(h_last[1], np_(last.physical_state.omega))


@pytest.mark.parametrize("env,ctrl", [("pendulum", ["theta"]), ("cart_pole", ["theta", "deflection"]),
                                      ("mass_spring_damper", []), ("fluid_tank", ["height"]), ("acrobot", ["omega_2"]),
                                      ("pmsm", ["i_d", "i_q"]), ("pmsm", ["torque"])])
def test_standalone_rew_trunc_term_ahead_equals_fused(excenvs, env, ctrl):
    """vmap_generate_rew_trunc_term_ahead (reference core_env.py:618-647) on the materialised trajectory == the
    outputs fused into the trajectory kernel == the oracle."""
    rng = np.random.default_rng(3)
    B, T = 257, 40
    orc = O.Oracle(env, solver=O.TSIT5, dtype=np.float64, control_state=ctrl, tau=1e-4)
    base = O.Oracle(env, dtype=np.float64)
    st = random_state(base, B, rng, 0.6, pmsm_gentle=(env == "pmsm"))
    refl = random_state(base, B, rng, 0.6)
    acts = rng.uniform(-1, 1, (B, T, orc.A)) * (0.02 if env == "pmsm" else 1.0)
    # PMSM: the reference's sim_ahead drops init_state.reference and returns NaN references (pmsm_env.py:688-701), so
    # the *_ref observation columns and the rewards of a PMSM trajectory are NaN upstream -- reproduced and pinned here
    oracle_ref = [np.full(B, np.nan)] * len(refl) if env == "pmsm" else refl
    ref = orc.sim_ahead(st, acts, 1e-4, 1e-4, ref=oracle_ref if ctrl else None, epilogue=True)
    e = make_env(excenvs, orc, B)
    s0 = to_state(e, st, refl if ctrl else None)
    a = torch.as_tensor(acts, device="cuda")
    obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(s0, a, 1e-4, 1e-4)
    if env == "pmsm":
        assert bool(torch.isnan(obs[..., 8:]).all())  # (the rewards follow from NaN references: compared below)
        assert bool(torch.isnan(states.reference.i_d).all()) and bool(torch.isnan(last.reference.torque).all())
    rew2, trunc2, term2 = e.vmap_generate_rew_trunc_term_ahead(states, a)
    assert rew2.shape == (B, T, 1) and trunc2.shape == (B, T + 1, orc.n_trunc) and term2.shape == (B, T, 1)
    assert torch.allclose(rew, rew2, rtol=1e-12, atol=1e-14, equal_nan=True)
    assert torch.equal(trunc, trunc2) and torch.equal(term, term2)
    assert np.allclose(np_(rew)[..., 0], ref["reward"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.array_equal(np_(term)[..., 0].astype(np.uint8), ref["terminated"])
    # GymWrapper-style epilogue of a single batched state
    r1, tr1, te1 = e.vmap_generate_reward_truncated_terminated(last)
    assert torch.allclose(r1[:, 0], rew[:, -1, 0], rtol=1e-12, atol=1e-14, equal_nan=True) and torch.equal(tr1, trunc[:, -1])


@pytest.mark.parametrize("env", ["mass_spring_damper", "pendulum", "cart_pole", "pmsm"])
@pytest.mark.parametrize("solver", [O.EULER, O.TSIT5])
def test_time_major_equals_batch_major(excenvs, env, solver):
    """Both layouts run the same per-environment arithmetic (exenv_sim_core.cuh): the time-major views hold the same
    values as the reference-contiguous arrays (to rounding: the two kernels are separate compilations), ragged batch,
    zero-copy consumption of a time-major action view."""
    rng = np.random.default_rng(17)
    orc = O.Oracle(env, solver=solver, dtype=np.float32)
    B, T = 1000 + 37, 96
    gentle = env == "pmsm"
    st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=gentle)
    acts = rng.uniform(-1, 1, (B, T, orc.A)).astype(np.float32) * (0.02 if gentle else 1.0)
    e = make_env(excenvs, orc, B)
    a_bm = torch.as_tensor(acts, device="cuda")
    a_tm = e.time_major_actions(T, fill=a_bm)
    assert a_tm.shape == a_bm.shape and torch.equal(a_tm, a_bm) and a_tm.permute(1, 0, 2).is_contiguous()
    o1, s1, l1 = e.vmap_sim_ahead(to_state(e, st), a_bm, orc.tau, orc.tau)
    o2, s2, l2 = e.vmap_sim_ahead(to_state(e, st), a_tm, orc.tau, orc.tau, layout="time_major")
    assert o1.shape == o2.shape and o1.is_contiguous() and not o2.is_contiguous()
    scales = field_scales(orc)
    for name, sc in zip(e.PHYS_FIELDS, scales):
        x1, x2 = np_(getattr(s1.physical_state, name)), np_(getattr(s2.physical_state, name))
        assert x1.shape == x2.shape == (B, T + 1)
        assert traj_err(x2, x1, sc) <= 2e-6, name
        assert np.array_equal(np_(getattr(l2.physical_state, name)), x2[:, -1])
    assert np.max(np.abs(np_(o1) - np_(o2))) <= 1e-5


@pytest.mark.parametrize("env,solver", [("mass_spring_damper", O.TSIT5), ("pendulum", O.EULER), ("cart_pole", O.TSIT5),
                                        ("fluid_tank", O.TSIT5)])
def test_time_major_envs_per_thread_and_fallbacks(excenvs, env, solver):
    """Time-major kernel: 1, 2 or 4 environments per thread (vector accesses) give the same trajectories; a batch
    that does not split evenly falls back to narrower vectors; the fused-epilogue flavour matches the plain one.
    The packed-arithmetic instantiations (fp32, 2 or 4 environments per thread: FFMA2) perform the same algorithm as
    the scalar one-environment-per-thread kernel; compared to rounding."""
    rng = np.random.default_rng(23)
    orc = O.Oracle(env, solver=solver, dtype=np.float32)
    # fp32 kernels with 2 or 4 environments per thread use packed arithmetic (FFMA2; sine / cosine through the packed
    # polynomial pair): same algorithm as the one-environment-per-thread scalar kernel, values equal to rounding
    # (amplified along the 150 coarse steps of the chaotic cart-pole)
    tol = {"mass_spring_damper": 2e-6, "pendulum": 1e-5, "fluid_tank": 5e-6, "cart_pole": 2e-3}[env]
    scales = field_scales(orc)

    def same(x, y, scale=1.0):
        return ((x - y).abs().max() / scale).item() <= tol

    for B in (4 * 97, 4 * 97 + 2, 4 * 97 + 1):  # V = 4 / 2 / 1 at most
        T = 150
        st = random_state(orc, B, rng, spread=0.5)
        acts = torch.as_tensor(rng.uniform(-1, 1, (B, T, orc.A)).astype(np.float32), device="cuda")
        e = make_env(excenvs, orc, B)
        base = e.vmap_sim_ahead(to_state(e, st), acts, orc.tau, orc.tau, layout="time_major", block_envs=1)
        for v in (0, 2, 4):
            other = e.vmap_sim_ahead(to_state(e, st), acts, orc.tau, orc.tau, layout="time_major", block_envs=v)
            assert same(base[0], other[0]), (B, v)
            for name, sc in zip(e.PHYS_FIELDS, scales):
                assert same(getattr(base[1].physical_state, name), getattr(other[1].physical_state, name), sc)
                assert same(getattr(base[2].physical_state, name), getattr(other[2].physical_state, name), sc)
        bm = e.vmap_sim_ahead_with_epilogue(to_state(e, st), acts, orc.tau, orc.tau)
        for v in (0, 1, 2, 4):  # fused-epilogue flavour of every width (v = 0: the automatic choice)
            full = e.vmap_sim_ahead_with_epilogue(to_state(e, st), acts, orc.tau, orc.tau, layout="time_major",
                                                  block_envs=v)
            assert same(base[0], full[0]), (B, v)
            # flags: the reference expression on the states this flavour returned, bit for bit
            own = [np_(getattr(full[1].physical_state, n)) for n in e.PHYS_FIELDS]
            assert np.array_equal(np_(full[4]).astype(np.uint8), flags_from_states(orc, own)), (B, v)
            assert torch.equal(full[5], bm[5]), (B, v)
            assert torch.equal(full[3], bm[3]), (B, v)
            assert full[3].shape == (B, T, 1) and full[4].shape == (B, T + 1, orc.n_trunc) and full[5].shape == (B, T, 1)


def test_time_major_c_abi_misaligned_and_partial_outputs(excenvs):
    """Raw C-ABI call in the time-major layout with buffers that rule out vector accesses (4-byte aligned only) and with
    a skipped trajectory leaf: the generic instantiation must give the same values as the vectorised one."""
    import ctypes as C

    L = excenvs._lib
    B, T = 256, 64
    e = excenvs.MassSpringDamper(batch_size=B, solver=excenvs.Tsit5(), dtype=torch.float32)
    _, s = e.vmap_reset(rng=3)
    acts = torch.rand((T, B, 1), device="cuda") * 2 - 1
    obs_ref, states_ref, _ = e.vmap_sim_ahead(s, acts.permute(1, 0, 2), e.tau, e.tau, layout="time_major")
    pad = 1  # one float off a 16-byte boundary
    a_buf = torch.empty(T * B + pad, device="cuda")
    a_buf[pad:] = acts.reshape(-1)
    o_buf = torch.empty((T + 1) * B * 2 + pad, device="cuda")
    t_buf = torch.empty((T + 1) * B + pad, device="cuda")
    a = L.ExenvSimArgs()
    a.batch, a.n_actions, a.obs_stepsize, a.action_stepsize, a.max_steps = B, T, e.tau, e.tau, 4096
    a.layout = L.LAYOUT_TIME_MAJOR
    leaves = [getattr(s.physical_state, n).contiguous() for n in e.PHYS_FIELDS]
    for i in range(2):
        a.state_in[i] = leaves[i].data_ptr()
    a.traj[1] = t_buf.data_ptr() + 4 * pad  # leaf 0 skipped (NULL)
    a.actions, a.obs = a_buf.data_ptr() + 4 * pad, o_buf.data_ptr() + 4 * pad
    L.check(L.lib().exenv_sim_ahead(e._plan().handle, C.byref(a), e._stream()))
    torch.cuda.synchronize()
    assert torch.equal(o_buf[pad:].reshape(T + 1, B, 2).permute(1, 0, 2), obs_ref)
    assert torch.equal(t_buf[pad:].reshape(T + 1, B).t(), states_ref.physical_state.velocity)
    a.layout = 7
    assert L.lib().exenv_sim_ahead(e._plan().handle, C.byref(a), e._stream()) == -1  # EXENV_ERR_INVALID


def test_config4_pmsm_saturated_full_horizon_properties(excenvs):
    """BASELINE config 4 scale check (PMSM, saturated BRUSA flux maps, Tsit5, fused epilogue, 5000 steps; the reference
    itself stops at max_steps = 4096, lifted here; per GPU the batch is a shard).  Size-independent properties:
    * a 2500-step solve reproduces the first half of the 5000-step solve up to its last sample.  (The last sample
      differs by construction: the reference's zero-order hold is evaluated at the END of a Tsit5 step for stages 6
      and 7 (c = 1), i.e. with the NEXT action, which a solve's final step does not have -- the index is clamped,
      pendulum_env.py:215-216.  Later chunks do not line up sample by sample either: the fp32 time accumulation of
      diffeqsolve makes the ZOH index lag differently at t = 0.25 s than at t = 0, SURVEY C2/C3 -- both reproduced.)
    * the fused truncated / terminated / reward outputs equal their definition recomputed from the returned
      observations (pmsm_env.py:972-983);
    * both memory layouts hold the same trajectories."""
    B, T = 1 << 14, 5000
    orc = _orc("pmsm_brusa_sat", O.TSIT5, np.float32)  # configuration + state sampler only (no oracle compute here)
    rng = np.random.default_rng(5)
    st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=True)
    e = make_env(excenvs, orc, B)
    e.max_steps = None
    s0 = to_state(e, st)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = (torch.rand((T, B, 2), device="cuda", generator=g) * 2 - 1).permute(1, 0, 2) * 0.02
    o_full, st_full, last_full, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(s0, acts, e.tau, e.tau,
                                                                                  layout="time_major")
    assert o_full.shape == (B, T + 1, 8) and trunc.shape == (B, T + 1, 1) and rew.shape == (B, T, 1)
    assert bool(torch.isfinite(o_full).all())
    o1, st1, last1 = e.vmap_sim_ahead(s0, acts[:, : T // 2], e.tau, e.tau, layout="time_major")
    scales = dict(zip(e.PHYS_FIELDS, field_scales(orc)))
    for name in ("i_d", "i_q", "torque", "epsilon"):
        full, c1 = getattr(st_full.physical_state, name), getattr(st1.physical_state, name)
        assert ((full[:, : T // 2] - c1[:, :-1]).abs().max() / scales[name]).item() < 1e-3, name
    # epilogue definition on the returned observations (entries 0, 1 = normalised i_d, i_q)
    own = [np_(getattr(st_full.physical_state, n)) if n in ("i_d", "i_q") else None for n in e.PHYS_FIELDS]
    assert np.array_equal(np_(trunc).astype(np.uint8), flags_from_states(orc, own))  # bit exact, every sample
    assert torch.equal(term, trunc[:, 1:]) and bool((rew == 0).all())
    # reference-contiguous layout: same values (separate compilation: to rounding)
    o_bm, st_bm, _ = e.vmap_sim_ahead(s0, acts, e.tau, e.tau)
    for name in ("i_d", "i_q", "torque"):
        a, b = getattr(st_full.physical_state, name), getattr(st_bm.physical_state, name)
        assert ((a - b).abs().max() / scales[name]).item() < 1e-3, name


@pytest.mark.parametrize("motor", ["brusa", "sew"])
def test_persistent_pmsm_kernel_job_mapping(excenvs, motor):
    """The coefficient-form saturated-PMSM kernel (Pmsm<float, 2>) runs one persistent 512-thread CTA per SM whose warps
    walk 32-environment jobs (exenv_simtm_common.cuh).  A ragged batch of more than one round (148 SMs x 512 threads):
    some warps run two jobs, the last job is partial.  Every environment must get the trajectory the same inputs get
    in a small batch (a different job, warp and CTA: bit-identical), a sample of them is compared with the oracle,
    and the fused flags equal their definition on the returned states."""
    import torch.cuda

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B, T = sms * 512 + 5 * 32 + 13, 24
    orc = _orc(f"pmsm_{motor}_sat", O.TSIT5, np.float32)
    rng = np.random.default_rng(17)
    st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=True)
    acts = (rng.uniform(-1, 1, (B, T, 2)) * 0.02).astype(np.float32)
    e = make_env(excenvs, orc, B)
    a_dev = torch.as_tensor(acts, device="cuda")
    obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(to_state(e, st), a_dev, orc.tau, orc.tau,
                                                                         layout="time_major")
    # the same environments in a small batch, in another order of jobs
    pick = np.concatenate([np.arange(0, 64), np.arange(B - 77, B), rng.choice(B, 500, replace=False)])
    sub = [x[pick] for x in st]
    e2 = make_env(excenvs, orc, pick.size)
    obs2, states2, last2, rew2, trunc2, term2 = e2.vmap_sim_ahead_with_epilogue(
        to_state(e2, sub), torch.as_tensor(acts[pick], device="cuda"), orc.tau, orc.tau, layout="time_major")
    idx = torch.as_tensor(pick, device="cuda")
    assert torch.equal(obs[idx], obs2) and torch.equal(trunc[idx], trunc2) and torch.equal(term[idx], term2)
    for n in e.PHYS_FIELDS:
        assert torch.equal(getattr(states.physical_state, n)[idx], getattr(states2.physical_state, n)), n
        assert torch.equal(getattr(last.physical_state, n)[idx], getattr(last2.physical_state, n)), n
    for a, b in zip(last.additions.solver_state, last2.additions.solver_state):
        assert torch.equal(a[idx], b)
    # against the oracle on the sample
    ref = orc.sim_ahead(sub, acts[pick], orc.tau, orc.tau)
    scales = field_scales(orc)
    for name, b, s_ in zip(e.PHYS_FIELDS, ref["traj"], scales):
        err = traj_err(np_(getattr(states2.physical_state, name)), b, s_)
        assert err <= 1e-4, (name, err)
    own = [np_(getattr(states.physical_state, n)) if n in ("i_d", "i_q") else None for n in e.PHYS_FIELDS]
    assert np.array_equal(np_(trunc).astype(np.uint8), flags_from_states(orc, own))
    assert bool(torch.isfinite(obs).all()) and bool((rew == 0).all())


@pytest.mark.parametrize("layout", ["batch_major", "time_major"])
@pytest.mark.parametrize("env,solver", [("mass_spring_damper", O.TSIT5), ("pendulum", O.EULER), ("pmsm", O.TSIT5),
                                        ("pmsm_brusa_sat", O.TSIT5)])
def test_trajectory_kernels_write_only_their_outputs(excenvs, env, solver, layout):
    """Guard zones around every output buffer stay untouched (ragged batches, every envs-per-thread width, the first
    step's double save, the L2 prefetch running ahead of the action stream near its end; the saturated PMSM runs the
    persistent coefficient-form kernel in the time-major layout: jobs per warp, surplus lanes of the last job)."""
    rng = np.random.default_rng(31)
    orc = _orc(env, solver, np.float32)
    G = 256  # guard floats on either side
    for B, T in ((388, 37), (390, 5), (129, 1), (4 * 1024, 9)):
        st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=(orc.env == "pmsm"))
        acts = torch.as_tensor(rng.uniform(-1, 1, (B, T, orc.A)).astype(np.float32), device="cuda") * 0.05
        e = make_env(excenvs, orc, B)
        NS, OW, D = T + 1, e.obs_dim, e.physical_state_dim
        sizes = [NS * B * OW] + [NS * B] * D
        buf = torch.full((sum(sizes) + G * (len(sizes) + 1),), 12345.0, device="cuda")
        views, off = [], G
        for n in sizes:
            views.append(buf[off:off + n])
            off += n + G
        tm = layout == "time_major"
        out = dict(obs=views[0].view((NS, B, OW) if tm else (B, NS, OW)),
                   traj=[v.view((NS, B) if tm else (B, NS)) for v in views[1:]])
        for v in ((0, 1, 2, 4) if tm else (0,)):
            for x in views:
                x.fill_(777.0)
            e._sim_impl(e.env_properties, B, to_state(e, st), acts, orc.tau, orc.tau, out=out, layout=layout,
                        block_envs=v)
            torch.cuda.synchronize()
            off = 0
            for n in sizes:
                assert bool((buf[off:off + G] == 12345.0).all()), (B, T, v, "guard before")
                off += G + n
            assert bool((buf[off:off + G] == 12345.0).all()), (B, T, v, "guard after")
            for x in views:
                assert not bool((x == 777.0).any()), (B, T, v, "unwritten output")


@pytest.mark.parametrize("layout", ["time_major", "batch_major"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_state_on_the_edge_of_its_range_is_not_truncated(excenvs, layout, dtype):
    """truncated = abs(obs) > 1 is a statement about the reference's observation 2 (x - min) / (max - min) - 1, which is
    exactly +-1 on the edge of the range.  The time-major kernels return a one-FMA normalisation (within a few ulp of 1 of it)
    but must evaluate the flag on the reference form: a state exactly on the edge is not flagged, one beyond it
    is, for ranges whose reciprocal is not exact."""
    tau, T = 1e-4, 4
    for lim in (0.37, 1.3, 7.0, 10.0, 41.0):
        pn = {"deflection": excenvs.MinMaxNormalization(-lim, lim), "velocity": excenvs.MinMaxNormalization(0.0, lim)}
        e = excenvs.MassSpringDamper(batch_size=8, physical_normalizations=pn, solver=excenvs.Tsit5(), dtype=dtype,
                                     tau=tau)
        _, s = e.vmap_reset()
        edge = torch.tensor(lim, dtype=dtype, device="cuda")
        beyond = torch.nextafter(edge, torch.tensor(float("inf"), dtype=dtype, device="cuda"))
        far = edge * 1.001
        defl = torch.stack([edge, -edge, beyond, -beyond, edge, -edge, far, -far])
        vel = torch.stack([edge, edge * 0, edge, edge * 0, beyond, edge * 0 - 1e-30, far, edge * 0])
        leaves = [defl.to(dtype), vel.to(dtype)]
        st = to_state(e, [np_(x) for x in leaves])
        acts = torch.zeros((8, T, 1), dtype=dtype, device="cuda")
        out = e.vmap_sim_ahead_with_epilogue(st, acts, tau, tau, layout=layout)
        obs0, trunc0 = out[0][:, 0], out[4][:, 0]
        # the reference expression with IEEE divisions, on the CPU (torch's CUDA `tensor / scalar` multiplies by the
        # reciprocal, which is exactly the difference this test is about)
        ndt = np.float32 if dtype == torch.float32 else np.float64
        x0, x1, lo, hi = np_(leaves[0]), np_(leaves[1]), ndt(-lim), ndt(lim)
        ref_np = np.stack([ndt(2) * (x0 - lo) / (hi - lo) - ndt(1), ndt(2) * (x1 - ndt(0)) / (hi - ndt(0)) - ndt(1)], axis=1)
        assert ref_np.dtype == ndt
        ref = torch.as_tensor(ref_np, device="cuda")
        assert np.array_equal(np_(trunc0).astype(bool), np.abs(ref_np) > 1), (lim, dtype, layout)
        # (one ulp beyond the edge still normalises to exactly 1 in the reference's own arithmetic: rows 2-5 are checked
        # against the reference expression above only)
        assert not trunc0[0].any() and not trunc0[1].any() and trunc0[6].all() and trunc0[7, 0] and not trunc0[7, 1]
        eps = torch.finfo(dtype).eps
        assert (obs0 - ref).abs().max().item() <= 6 * eps
