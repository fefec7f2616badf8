"""Developer micro-benchmarks (not the driver's bench.py): device-timed throughput of the step and sim_ahead kernels
over a few configurations, plus a tile-geometry sweep.  Prints one JSON object per line."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import exciting_environments_b200 as ex  # noqa: E402

HBM = 6559.4e9


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def bench_step(name, cls, solver, dtype, B):
    env = cls(batch_size=B, solver=solver, dtype=dtype)
    _, state = env.vmap_reset(rng=0)
    S = 4 if dtype == torch.float32 else 8
    D, A, O = env.physical_state_dim, env.action_dim, env.obs_dim
    action = torch.rand((B, A), device="cuda", dtype=dtype) * 2 - 1
    t = timed(lambda: env.vmap_step(state, action, inplace=True))
    bytes_ = S * (2 * D + A + O + (env.N_ODE if isinstance(solver, ex.Tsit5) else 0)) * B
    print(json.dumps(dict(kind="step", env=name, solver=repr(solver), dtype=str(dtype), B=B, ms=t * 1e3,
                          env_steps_per_s=B / t, GBps=bytes_ / t / 1e9, frac_hbm=bytes_ / t / HBM)), flush=True)


def bench_sim(name, cls, solver, dtype, B, T, tile_steps=0, block_envs=0, iters=3, layout="batch_major", **kw):
    env = cls(batch_size=B, solver=solver, dtype=dtype, **kw)
    env.max_steps = None
    _, state = env.vmap_reset(rng=0)
    S = 4 if dtype == torch.float32 else 8
    D, A, O = env.physical_state_dim, env.action_dim, env.obs_dim
    tm = layout == "time_major"
    if tm:
        actions = (torch.rand((T, B, A), device="cuda", dtype=dtype) * 2 - 1).permute(1, 0, 2)
        obs = torch.empty((T + 1, B, O), device="cuda", dtype=dtype)
        traj = [torch.empty((T + 1, B), device="cuda", dtype=dtype) for _ in range(D)]
    else:
        actions = torch.rand((B, T, A), device="cuda", dtype=dtype) * 2 - 1
        obs = torch.empty((B, T + 1, O), device="cuda", dtype=dtype)
        traj = [torch.empty((B, T + 1), device="cuda", dtype=dtype) for _ in range(D)]
    out = dict(obs=obs, traj=traj)
    t = timed(lambda: env._sim_impl(env.env_properties, B, state, actions, env.tau, env.tau, tile_steps=tile_steps,
                                    block_envs=block_envs, out=out, layout=layout), warm=2, iters=iters)
    bytes_ = S * (A + O + D) * B * T
    print(json.dumps(dict(kind="sim", layout=layout, env=name, solver=repr(solver), dtype=str(dtype), B=B, T=T, tile=tile_steps,
                          block=block_envs, ms=t * 1e3, env_steps_per_s=B * T / t, GBps=bytes_ / t / 1e9,
                          frac_hbm=bytes_ / t / HBM)), flush=True)
    del obs, traj, actions


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    f32, f64 = torch.float32, torch.float64
    if what in ("all", "step"):
        for name, cls in (("pendulum", ex.Pendulum), ("msd", ex.MassSpringDamper), ("tank", ex.FluidTank),
                          ("cartpole", ex.CartPole), ("acrobot", ex.Acrobot), ("pmsm", ex.PMSM)):
            for solver in (ex.Euler(), ex.Tsit5()):
                for dt in (f32, f64):
                    bench_step(name, cls, solver, dt, 1 << 24)
        bench_step("pendulum", ex.Pendulum, ex.Euler(), f32, 1 << 26)
    if what in ("all", "sweep"):
        for tile in (8, 16, 32, 64):
            for blk in (64, 128, 256):
                try:
                    bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f32, 1 << 18, 500, tile, blk)
                except Exception as e:  # smem limit
                    print(json.dumps(dict(kind="sim", tile=tile, block=blk, error=str(e)[:100])), flush=True)
    if what in ("all", "sim"):
        bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f32, 1 << 20, 2000, iters=2)
        bench_sim("msd", ex.MassSpringDamper, ex.Euler(), f32, 1 << 20, 2000, iters=2)
        bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f64, 1 << 19, 2000, iters=2)
        bench_sim("pendulum", ex.Pendulum, ex.Tsit5(), f32, 1 << 20, 1000, iters=2)
        bench_sim("cartpole", ex.CartPole, ex.Tsit5(), f32, 1 << 22, 200, iters=2)
        bench_sim("acrobot", ex.Acrobot, ex.Tsit5(), f32, 1 << 22, 200, iters=2)
        bench_sim("pmsm", ex.PMSM, ex.Tsit5(), f32, 1 << 18, 1000, iters=2)
        bench_sim("pmsm_sat", ex.PMSM, ex.Tsit5(), f32, 1 << 17, 1000, iters=2, saturated=True,
                  motor_variant=ex.MotorVariant.BRUSA)
    if what in ("all", "simtm"):
        tm = dict(layout="time_major", iters=2)
        bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f32, 1 << 20, 2000, **tm)
        bench_sim("msd", ex.MassSpringDamper, ex.Euler(), f32, 1 << 20, 2000, **tm)
        bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f64, 1 << 19, 2000, **tm)
        bench_sim("msd", ex.MassSpringDamper, ex.Euler(), f64, 1 << 19, 2000, **tm)
        bench_sim("tank", ex.FluidTank, ex.Tsit5(), f32, 1 << 20, 1000, **tm)
        bench_sim("pendulum", ex.Pendulum, ex.Tsit5(), f32, 1 << 20, 1000, **tm)
        bench_sim("pendulum", ex.Pendulum, ex.Euler(), f32, 1 << 20, 1000, **tm)
        bench_sim("cartpole", ex.CartPole, ex.Tsit5(), f32, 1 << 22, 200, **tm)
        bench_sim("cartpole", ex.CartPole, ex.Euler(), f32, 1 << 22, 200, **tm)
        bench_sim("acrobot", ex.Acrobot, ex.Tsit5(), f32, 1 << 22, 200, **tm)
        bench_sim("acrobot", ex.Acrobot, ex.Euler(), f32, 1 << 22, 200, **tm)
        bench_sim("pmsm", ex.PMSM, ex.Tsit5(), f32, 1 << 18, 2500, **tm)
        bench_sim("pmsm", ex.PMSM, ex.Euler(), f32, 1 << 18, 2500, **tm)
        bench_sim("pmsm_sat", ex.PMSM, ex.Tsit5(), f32, 1 << 17, 2500, saturated=True,
                  motor_variant=ex.MotorVariant.BRUSA, **tm)
    if what in ("all", "simtm_v"):  # environments per thread of the time-major kernel
        cases = [("msd", ex.MassSpringDamper, 1 << 20, 2000, {}), ("tank", ex.FluidTank, 1 << 20, 1000, {}),
                 ("pendulum", ex.Pendulum, 1 << 20, 1000, {}), ("cartpole", ex.CartPole, 1 << 22, 200, {}),
                 ("acrobot", ex.Acrobot, 1 << 22, 200, {}), ("pmsm", ex.PMSM, 1 << 18, 2500, {}),
                 ("pmsm_sat", ex.PMSM, 1 << 17, 2500, dict(saturated=True, motor_variant=ex.MotorVariant.BRUSA))]
        for name, cls, B, T, kw in cases:
            for solver in (ex.Tsit5(), ex.Euler()):
                for dt in (f32, f64):
                    for v in (1, 2, 4):
                        if v == 4 and name.startswith("pmsm"):
                            continue
                        bench_sim(name, cls, solver, dt, B if dt == f32 else B // 2, T, block_envs=v,
                                  layout="time_major", iters=2, **kw)
    if what in ("all", "config1"):  # BASELINE configs[0]: the launch-latency regime (pendulum, B = 5, Euler, tau = 2e-2)
        B, N = 5, 1000
        env = ex.EnvironmentRegistry.PENDULUM.make(batch_size=B, tau=2e-2,
                                                   action_normalizations={"torque": ex.MinMaxNormalization(min=-15, max=15)})
        _, state = env.vmap_reset()
        actions = torch.linspace(-1, 1, N, device="cuda")[None, :, None].repeat(B, 1, 1)
        for i in range(50):
            _, state = env.vmap_step(state, actions[:, i, :])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            _, state = env.vmap_step(state, actions[:, i, :])
        torch.cuda.synchronize()
        t_eager = (time.perf_counter() - t0) / N
        # the same 1000 launches captured once into a CUDA graph (in-place state, actions read from the (B, T, A) tensor)
        _, sg = env.vmap_reset()
        acts = [actions[:, i, :].contiguous() for i in range(N)]
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                s = sg
                for i in range(N):
                    _, s = env.vmap_step(s, acts[i], inplace=True)
        g.replay()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            g.replay()
        torch.cuda.synchronize()
        t_graph = (time.perf_counter() - t0) / (10 * N)
        # one persistent sim_ahead launch over the same 1000 actions
        env2 = ex.EnvironmentRegistry.PENDULUM.make(batch_size=B, tau=2e-2,
                                                    action_normalizations={"torque": ex.MinMaxNormalization(min=-15, max=15)})
        _, s2 = env2.vmap_reset()
        env2.vmap_sim_ahead(s2, actions, 2e-2, 2e-2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            env2.vmap_sim_ahead(s2, actions, 2e-2, 2e-2)
        torch.cuda.synchronize()
        t_sim = (time.perf_counter() - t0) / (10 * N)
        print(json.dumps(dict(kind="config1", B=B, N=N, us_per_step_eager=t_eager * 1e6, us_per_step_graph=t_graph * 1e6,
                              us_per_step_sim_ahead=t_sim * 1e6, env_steps_per_s_eager=B / t_eager,
                              env_steps_per_s_graph=B / t_graph, env_steps_per_s_sim_ahead=B / t_sim)), flush=True)
    if what in ("all", "sweep_step"):  # BASELINE configs[4]: batch-size sweep of vmap_step, all six envs
        for name, cls in (("pendulum", ex.Pendulum), ("msd", ex.MassSpringDamper), ("tank", ex.FluidTank),
                          ("cartpole", ex.CartPole), ("acrobot", ex.Acrobot), ("pmsm", ex.PMSM)):
            for solver in (ex.Euler(), ex.Tsit5()):
                for lb in range(10, 29, 2):
                    bench_step(name, cls, solver, f32, 1 << lb)
