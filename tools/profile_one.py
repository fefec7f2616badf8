"""Run ONE configuration of the step or sim_ahead kernel a few times (for `ncu` captures and quick timing).
usage: profile_one.py sim|step <env> <euler|tsit5> <f32|f64> <B> [T] [tile_steps] [block_envs] [bm|tm]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import exciting_environments_b200 as ex  # noqa: E402

ENVS = dict(pendulum=ex.Pendulum, msd=ex.MassSpringDamper, tank=ex.FluidTank, cartpole=ex.CartPole, acrobot=ex.Acrobot,
            pmsm=ex.PMSM, pmsm_sat=ex.PMSM)
kind, name, solver, dt, B = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
T = int(sys.argv[6]) if len(sys.argv) > 6 else 500
tile = int(sys.argv[7]) if len(sys.argv) > 7 else 0
blk = int(sys.argv[8]) if len(sys.argv) > 8 else 0
tm = len(sys.argv) > 9 and sys.argv[9] == "tm"
epi = len(sys.argv) > 10 and sys.argv[10] == "epi"  # fused reward / truncated / terminated outputs
dtype = torch.float32 if dt == "f32" else torch.float64
import os
SCALE = float(os.environ.get("EXENV_PROFILE_ACT_SCALE", "1.0"))  # e.g. 0.02: PMSM voltages that keep the currents inside the flux maps
kw = dict(saturated=True, motor_variant=ex.MotorVariant.BRUSA) if name == "pmsm_sat" else {}
env = ENVS[name](batch_size=B, solver=ex.Tsit5() if solver == "tsit5" else ex.Euler(), dtype=dtype, **kw)
env.max_steps = None
_, state = env.vmap_reset(rng=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if kind == "sim":
    if tm:
        actions = ((torch.rand((T, B, env.action_dim), device="cuda", dtype=dtype) * 2 - 1) * SCALE).permute(1, 0, 2)
        out = dict(obs=torch.empty((T + 1, B, env.obs_dim), device="cuda", dtype=dtype),
                   traj=[torch.empty((T + 1, B), device="cuda", dtype=dtype) for _ in range(env.physical_state_dim)])
    else:
        actions = (torch.rand((B, T, env.action_dim), device="cuda", dtype=dtype) * 2 - 1) * SCALE
        out = dict(obs=torch.empty((B, T + 1, env.obs_dim), device="cuda", dtype=dtype),
                   traj=[torch.empty((B, T + 1), device="cuda", dtype=dtype) for _ in range(env.physical_state_dim)])
    fn = lambda: env._sim_impl(env.env_properties, B, state, actions, env.tau, env.tau, tile_steps=tile,
                               block_envs=blk, out=out, layout="time_major" if tm else "batch_major",
                               epilogue=epi)
    units = B * T
else:
    action = torch.rand((B, env.action_dim), device="cuda", dtype=dtype) * 2 - 1
    fn = lambda: env.vmap_step(state, action, inplace=True)
    units = B
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    fn()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3 * 1e-3
print(f"{kind} {name} {solver} {dt} {'tm' if tm else 'bm'}{' epi' if kind == 'sim' and epi else ''} B={B} T={T} tile={tile} blk={blk}: {t*1e3:.3f} ms  {units/t:.4e} env-steps/s")
