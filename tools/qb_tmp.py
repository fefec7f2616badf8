import sys, torch
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import quick_bench as q
ex = q.ex
f32, f64 = torch.float32, torch.float64
for v in (1, 2, 4):
    q.bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f32, 1 << 20, 2000, block_envs=v, layout="time_major", iters=2)
for v in (2, 4):
    q.bench_sim("msd", ex.MassSpringDamper, ex.Euler(), f32, 1 << 20, 2000, block_envs=v, layout="time_major", iters=2)
q.bench_sim("msd", ex.MassSpringDamper, ex.Tsit5(), f64, 1 << 19, 2000, block_envs=2, layout="time_major", iters=2)
q.bench_sim("pendulum", ex.Pendulum, ex.Tsit5(), f32, 1 << 20, 1000, block_envs=2, layout="time_major", iters=2)
q.bench_sim("pmsm", ex.PMSM, ex.Tsit5(), f32, 1 << 18, 2500, block_envs=1, layout="time_major", iters=2)
