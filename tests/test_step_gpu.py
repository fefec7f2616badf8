"""`-m gpu`: vmap_step through the C ABI (CUDA kernels) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from gpu_helpers import make_env, np_, to_state
from helpers import ENVS, field_scales, flags_from_states, golden_oracle, random_state, scaled_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu
# north_star tolerances: fp32 1e-6 relative per step, fp64 1e-12 per step.  "Relative" is measured against
# max(|x|, half-span of the field's normalisation box) -- see helpers.scaled_err.
TOL = {np.float64: 1e-12, np.float32: 1e-6}


def _orc(env, solver, dtype, **kw):
    if env.startswith("pmsm_"):
        return O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=True, motor=env.split("_")[1].upper(), **kw)
    return O.Oracle(env, solver=solver, dtype=dtype, **kw)


@pytest.mark.parametrize("env", ENVS)
def test_golden_vectors_replayed_on_gpu(excenvs, env):
    """reference tests/envs/<env>/test_<env>.py::test_step_results, kernel instead of JAX: fp64 Euler, closed loop,
    same tolerance as the reference's own assertion (rtol 1e-16 + atol 1e-8; PMSM rtol 1e-8)."""
    orc, g = golden_oracle(env)
    e = make_env(excenvs, orc, 1)
    obs0 = torch.as_tensor(g["observations"][0:1], dtype=torch.float64, device="cuda")
    state = e.vmap_generate_state_from_observation(obs0)
    acts = torch.as_tensor(g["actions"], dtype=torch.float64, device="cuda")
    n = acts.shape[0]
    out = torch.empty((n + 1, e.obs_dim), dtype=torch.float64, device="cuda")
    out[0] = obs0[0]
    for i in range(n):
        obs, state = e.vmap_step(state, acts[i:i + 1])
        out[i + 1] = obs[0]
    gen = np_(out)
    rtol = 1e-8 if env == "pmsm" else 1e-16
    assert np.allclose(gen, g["observations"], rtol=rtol, atol=1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("solver", [O.EULER, O.TSIT5])
@pytest.mark.parametrize("env", ENVS + ["pmsm_brusa_sat", "pmsm_sew_sat"])
def test_step_matches_oracle(excenvs, env, solver, dtype):
    rng = np.random.default_rng(1234)
    orc = _orc(env, solver, dtype)
    B = 4099  # 128-bit vector body + ragged scalar tail
    st = random_state(orc, B, rng)
    act = rng.uniform(-1, 1, (B, orc.A)).astype(dtype)
    ref = orc.step(st, act, epilogue=True)
    e = make_env(excenvs, orc, B)
    obs, new, rew, trunc, term = e.vmap_step_with_epilogue(to_state(e, st), torch.as_tensor(act, device="cuda"))
    sc = field_scales(orc)
    tol = TOL[dtype]
    if orc.saturated and solver == O.TSIT5 and dtype == np.float32:
        # the Tsit5 kernels solve L x = b directly (exenv_envs.cuh); at speed b cancels heavily and the fp32 reference
        # formula itself is up to 1e-4 from the fp64 result after one step: the bound is the reference formula's own
        # distance from fp64, the kernel is compared with the fp64 result
        ref64 = _orc(env, solver, np.float64).step([x.astype(np.float64) for x in st], act.astype(np.float64),
                                                   epilogue=True)
        noise = max(scaled_err(b, b64, s_) for b, b64, s_ in zip(ref["state"], ref64["state"], sc))
        tol = max(tol, 3 * noise)
        for key in ("state", "obs", "k"):
            ref[key] = ref64[key]
    for name, b, s_ in zip(e.PHYS_FIELDS, ref["state"], sc):
        err = scaled_err(np_(getattr(new.physical_state, name)), b, s_)
        assert err <= tol, (name, err)
    err = scaled_err(np_(obs), ref["obs"], 1.0)
    assert err <= tol, ("obs", err)
    if solver == O.TSIT5:  # solver_state leaf: derivative, compared as the increment tau * k
        ode_scale = [sc[e.PHYS_FIELDS.index(n)] for n in (("i_d", "i_q", "epsilon") if orc.env == "pmsm" else e.PHYS_FIELDS)]
        for a, b, s_ in zip(new.additions.solver_state, ref["k"], ode_scale):
            err = scaled_err(np_(a) * orc.tau, np.asarray(b, np.float64) * orc.tau, s_)
            assert err <= tol * 10, ("k", err)
    # flags are bit-exact: against the oracle (except where the ill-conditioned case above moves a state across the
    # threshold) and against the reference expression (IEEE NumPy) on the returned state
    if tol == TOL[dtype]:
        assert np.array_equal(np_(trunc).astype(np.uint8), ref["truncated"])
        assert np.array_equal(np_(term)[:, 0].astype(np.uint8), ref["terminated"])
    assert np.array_equal(np_(rew)[:, 0], ref["reward"])
    own = [np_(getattr(new.physical_state, n)) for n in e.PHYS_FIELDS]
    assert np.array_equal(np_(trunc).astype(np.uint8), flags_from_states(orc, own))
    assert bool(new.additions.active_solver_state.all())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("env", ["pendulum", "cart_pole", "acrobot", "mass_spring_damper", "fluid_tank"])
def test_step_with_reference_tracking_and_per_env_parameters(excenvs, env, dtype):
    """control_state + references (reward / terminated values) and (B,) parameter arrays (core_env.py:253-277)."""
    rng = np.random.default_rng(7)
    B = 1031
    base = O.Oracle(env, dtype=dtype)
    static = {k: (np.float64(v) * rng.uniform(0.9, 1.1, B)).astype(dtype) for k, v in base.static.items()}
    pn = dict(base.pn)
    first = base.spec.phys[0]
    lo, hi = pn[first]
    pn[first] = ((np.full(B, lo) * (1 + 0.05 * rng.uniform(0, 1, B)) - 0.01).astype(dtype),
                 (np.full(B, hi) * (1 + 0.05 * rng.uniform(0, 1, B)) + 0.01).astype(dtype))
    ctrl = list(base.spec.phys[:2])
    orc = O.Oracle(env, solver=O.TSIT5, dtype=dtype, static_params=static, physical_normalizations=pn,
                   control_state=ctrl)
    st = random_state(base, B, rng)
    ref_leaves = random_state(base, B, rng)
    act = rng.uniform(-1, 1, (B, orc.A)).astype(dtype)
    ref = orc.step(st, act, ref=ref_leaves, epilogue=True)
    e = make_env(excenvs, orc, B)
    assert e.in_axes_env_properties.static_params.__dict__[base.spec.static[0]] == 0
    obs, new, rew, trunc, term = e.vmap_step_with_epilogue(to_state(e, st, ref_leaves),
                                                           torch.as_tensor(act, device="cuda"))
    assert obs.shape == (B, orc.O)
    err = scaled_err(np_(obs), ref["obs"], 1.0)
    assert err <= TOL[dtype], ("obs", err)
    err = scaled_err(np_(rew)[:, 0], ref["reward"], 1.0)
    assert err <= TOL[dtype] * 10, ("reward", err)
    assert np.array_equal(np_(trunc).astype(np.uint8), ref["truncated"])
    assert np.array_equal(np_(term)[:, 0].astype(np.uint8), ref["terminated"])
    own = [np_(getattr(new.physical_state, n)) for n in e.PHYS_FIELDS]
    assert np.array_equal(np_(trunc).astype(np.uint8), flags_from_states(orc, own, ref_leaves))


def test_inplace_step_equals_out_of_place(excenvs):
    e = excenvs.CartPole(batch_size=5000, solver=excenvs.Tsit5(), dtype=torch.float32)
    _, s = e.vmap_reset(rng=3)
    a = torch.rand((5000, 1), device="cuda") * 2 - 1
    obs1, s1 = e.vmap_step(s, a)
    keep = [getattr(s1.physical_state, n).clone() for n in e.PHYS_FIELDS]
    obs2, s2 = e.vmap_step(s, a, inplace=True)
    assert torch.equal(obs1, obs2)
    for n, k in zip(e.PHYS_FIELDS, keep):
        assert torch.equal(getattr(s2.physical_state, n), k)
        assert getattr(s2.physical_state, n).data_ptr() == getattr(s.physical_state, n).data_ptr()


def test_error_paths(excenvs):
    e = excenvs.Pendulum(batch_size=8)
    _, s = e.vmap_reset()
    with pytest.raises(AssertionError):
        e.vmap_step(s, torch.zeros((7, 1), device="cuda"))
    with pytest.raises(ValueError):
        excenvs.PMSM(batch_size=2, saturated=True)  # DEFAULT motor has no LUT (pmsm_env.py:189-193)
    with pytest.raises(ValueError):
        excenvs.Pendulum(batch_size=2, static_params={"g": [9.81, 9.81], "l": 2, "m": 1})  # list rejected


def test_readme_example_config1(excenvs):
    """BASELINE.json configs[0] = the reference's README example (README.md:15-33): PENDULUM batch_size=5, explicit
    Euler, tau=2e-2, torque normalisation +-15, 1000 sequential vmap_step calls with actions linspace(-1, 1, 1000)
    repeated over the batch, default reset -- replayed by the oracle; the five trajectories are identical."""
    B, N = 5, 1000
    env = excenvs.EnvironmentRegistry.PENDULUM.make(
        batch_size=B, action_normalizations={"torque": excenvs.MinMaxNormalization(min=-15, max=15)}, tau=2e-2)
    orc = O.Oracle("pendulum", solver=O.EULER, dtype=np.float32, tau=2e-2, action_normalizations={"torque": (-15.0, 15.0)})
    obs, state = env.vmap_reset()
    actions = torch.linspace(-1, 1, N, device="cuda")[None, :, None].repeat(B, 1, 1)
    observations = [obs]
    for idx in range(N):
        obs, state = env.vmap_step(state, actions[:, idx, :])
        observations.append(obs)
    observations = torch.stack(observations, dim=1)
    assert observations.shape == (B, N + 1, 2)
    assert all(torch.equal(observations[0], observations[b]) for b in range(1, B))
    st = [np.full(B, np.pi, np.float32), np.zeros(B, np.float32)]  # default reset: theta = pi, omega = 0
    acts = np_(actions)
    cur = st
    out = []
    for idx in range(N):
        r = orc.step(cur, acts[:, idx, :])
        cur = r["state"]
        out.append(r["obs"])
    out = np.stack(out, axis=1)
    got = np_(observations)[:, 1:]
    # 1000 closed-loop fp32 Euler steps of a pendulum swinging through several revolutions: libm-level differences
    # (sinf) accumulate; the per-step bound (1e-6) is asserted by test_step_matches_oracle
    assert np.max(np.abs(got - out)) < 5e-3
    assert np.max(np.abs(got[:, :50] - out[:, :50])) < 1e-5


def test_step_is_cuda_graph_capturable(excenvs):
    """include/exenv_b200.h: launches are asynchronous on the caller's stream, no host synchronisation, no allocation
    -> a vmap_step call can be captured into a CUDA graph and replayed (launch-latency regime of configs[0])."""
    B = 5
    env = excenvs.Pendulum(batch_size=B, tau=2e-2, dtype=torch.float32)
    _, s_eager = env.vmap_reset()
    _, s_graph = env.vmap_reset()
    action = torch.zeros((B, 1), device="cuda")
    env.vmap_step(s_graph, action)  # plan creation happens outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            obs_g, s_out = env.vmap_step(s_graph, action, inplace=True)
    for i in range(20):
        action.fill_(0.1 * (i % 5) - 0.2)
        obs_e, s_eager = env.vmap_step(s_eager, action)
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(obs_e, obs_g)
    assert torch.equal(s_eager.physical_state.theta, s_out.physical_state.theta)


def test_first_sim_ahead_of_a_shape_is_cuda_graph_capturable(excenvs):
    """exenv_sim_ahead_prepare uploads the time schedule of a call shape ahead of time, so that exenv_sim_ahead itself
    never allocates or synchronises: the FIRST launch of a shape can be issued inside a stream capture (an XLA custom
    call).  Without the prepare call a capture gets an error, not a cudaMalloc."""
    import ctypes as C

    L = excenvs._lib
    B, T = 512, 37
    env = excenvs.MassSpringDamper(batch_size=B, solver=excenvs.Tsit5(), dtype=torch.float32)
    _, s = env.vmap_reset(rng=2)
    acts = torch.rand((T, B, 1), device="cuda") * 2 - 1
    leaves = [getattr(s.physical_state, n).contiguous() for n in env.PHYS_FIELDS]
    obs = torch.zeros((T + 1, B, 2), device="cuda")
    traj = [torch.zeros((T + 1, B), device="cuda") for _ in range(2)]
    a = L.ExenvSimArgs()
    a.batch, a.n_actions, a.obs_stepsize, a.action_stepsize, a.max_steps = B, T, env.tau, env.tau, 4096
    a.layout = L.LAYOUT_TIME_MAJOR
    for i in range(2):
        a.state_in[i], a.traj[i] = leaves[i].data_ptr(), traj[i].data_ptr()
    a.actions, a.obs = acts.data_ptr(), obs.data_ptr()
    plan = env._plan().handle
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    # (1) unprepared shape inside a capture: refused
    g0 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g0, stream=side):
            rc = L.lib().exenv_sim_ahead(plan, C.byref(a), C.c_void_p(side.cuda_stream))
    assert rc == -1 and "exenv_sim_ahead_prepare" in L.last_error()
    # (2) prepared: the very first launch of the shape is captured and replays
    ns, nst = C.c_int64(), C.c_int64()
    L.check(L.lib().exenv_sim_ahead_prepare(plan, T, env.tau, env.tau, 0, C.byref(ns), C.byref(nst)))
    assert ns.value == T + 1
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            L.check(L.lib().exenv_sim_ahead(plan, C.byref(a), C.c_void_p(side.cuda_stream)))
    g.replay()
    torch.cuda.synchronize()
    ref = env.vmap_sim_ahead(s, acts.permute(1, 0, 2), env.tau, env.tau, layout="time_major")[0]
    assert torch.equal(obs.permute(1, 0, 2), ref)


def test_entry_points_run_on_the_plans_device(excenvs):
    """Every entry point selects the plan's device and restores the calling thread's current device: plan
    creation, launches, the host-buffer pipeline and plan destruction leave torch's current device alone; with two GPUs
    an environment on cuda:1 works while cuda:0 is current and gives the same numbers."""
    import gc

    cur = torch.cuda.current_device()
    env = excenvs.Pendulum(batch_size=64, solver=excenvs.Tsit5(), dtype=torch.float32, device=f"cuda:{cur}")
    _, s = env.vmap_reset(rng=1)
    a = torch.rand((64, 1), device=f"cuda:{cur}") * 2 - 1
    obs0, s1 = env.vmap_step(s, a)
    acts = torch.rand((64, 20, 1), device=f"cuda:{cur}") * 2 - 1
    tr0 = env.vmap_sim_ahead(s, acts, env.tau, env.tau)[0]
    assert torch.cuda.current_device() == cur
    if torch.cuda.device_count() >= 2:
        other = (cur + 1) % torch.cuda.device_count()
        env2 = excenvs.Pendulum(batch_size=64, solver=excenvs.Tsit5(), dtype=torch.float32, device=f"cuda:{other}")
        st2 = env2._new_state([getattr(s.physical_state, n).to(f"cuda:{other}") for n in env2.PHYS_FIELDS],
                              s.PRNGKey.to(f"cuda:{other}"), env2._solver_state_dummy(64),
                              torch.zeros(64, dtype=torch.bool, device=f"cuda:{other}"),
                              [getattr(s.reference, n).to(f"cuda:{other}") for n in env2.PHYS_FIELDS])
        obs2, _ = env2.vmap_step(st2, a.to(f"cuda:{other}"))
        tr2 = env2.vmap_sim_ahead(st2, acts.to(f"cuda:{other}"), env2.tau, env2.tau)[0]
        assert torch.cuda.current_device() == cur
        assert obs2.device.index == other and torch.equal(obs2.cpu(), obs0.cpu()) and torch.equal(tr2.cpu(), tr0.cpu())
        del env2, st2
    del env
    gc.collect()  # _Plan.__del__ -> exenv_plan_destroy
    assert torch.cuda.current_device() == cur
