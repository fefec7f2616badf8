"""Generates tests/golden/self_generated/fixtures.npz from the CPU oracle (oracle/).

SELF-GENERATED, REFERENCE-UNPINNED: the reference's own tests only pin fp64 Euler `step` (tests/golden/reference/).
Everything else on the path -- Tsit5 step / sim_ahead, saturated flux-map PMSM, fp32, the zero-order-hold and save-grid
quirks of the diffeqsolve loop, per-environment parameters, reference tracking, reward / truncated / terminated -- has
no reference fixture, so these vectors freeze what the validated oracle computes today.  They
guard against silent changes of the oracle and give the GPU tests a committed target that does not depend on the
oracle being rebuilt on the GPU box.

    python tests/golden/self_generated/make_fixtures.py        # rewrites fixtures.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[3]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import random_state  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "fixtures.npz"
ENVS = ["pendulum", "mass_spring_damper", "fluid_tank", "cart_pole", "acrobot", "pmsm", "pmsm_brusa_sat", "pmsm_sew_sat"]
B, T = 4, 24


def make_oracle(env, solver, dtype, **kw):
    if env.startswith("pmsm_"):
        return O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=True, motor=env.split("_")[1].upper(), **kw)
    return O.Oracle(env, solver=solver, dtype=dtype, **kw)


def case_inputs(orc, seed):
    rng = np.random.default_rng(seed)
    gentle = orc.env == "pmsm"
    st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=gentle)
    acts = rng.uniform(-1, 1, (B, T, orc.A)).astype(orc.dtype) * (0.02 if gentle else 1.0)
    return st, acts


def main():
    out = {}
    for ei, env in enumerate(ENVS):
        for solver, sname in ((O.EULER, "euler"), (O.TSIT5, "tsit5")):
            for dtype, dname in ((np.float32, "f32"), (np.float64, "f64")):
                orc = make_oracle(env, solver, dtype)
                st, acts = case_inputs(orc, 1000 + ei)
                key = f"{env}/{sname}/{dname}"
                for i, x in enumerate(st):
                    out[f"{key}/state{i}"] = x
                out[f"{key}/actions"] = acts
                r = orc.step(st, acts[:, 0], epilogue=True)
                for i, x in enumerate(r["state"]):
                    out[f"{key}/step/state{i}"] = x
                out[f"{key}/step/obs"] = r["obs"]
                for i, x in enumerate(r["k"]):
                    out[f"{key}/step/k{i}"] = x
                out[f"{key}/step/truncated"] = r["truncated"]
                s = orc.sim_ahead(st, acts, orc.tau, orc.tau, epilogue=True)
                out[f"{key}/sim/obs"] = s["obs"]
                for i, x in enumerate(s["traj"]):
                    out[f"{key}/sim/traj{i}"] = x
                for i, x in enumerate(s["k_last"]):
                    out[f"{key}/sim/k_last{i}"] = x
                out[f"{key}/sim/truncated"] = s["truncated"]
    # zero-order-hold / save-grid quirks: obs_stepsize < action_stepsize, coarse tau (lagging ZOH index in fp32)
    for name, obs_dt, act_dt, n in (("half", 5e-5, 1e-4, 16), ("third", 1e-4, 3e-4, 12), ("coarse", 2e-2, 2e-2, 40)):
        for dtype, dname in ((np.float32, "f32"), (np.float64, "f64")):
            orc = O.Oracle("mass_spring_damper", solver=O.TSIT5, dtype=dtype, tau=obs_dt)
            rng = np.random.default_rng(77)
            st = random_state(orc, B, rng)
            acts = rng.uniform(-1, 1, (B, n, 1)).astype(dtype)
            s = orc.sim_ahead(st, acts, obs_dt, act_dt)
            key = f"quirk_{name}/{dname}"
            out[f"{key}/state0"], out[f"{key}/state1"], out[f"{key}/actions"] = st[0], st[1], acts
            out[f"{key}/steps"] = np.array([obs_dt, act_dt, s["n_steps"]])
            out[f"{key}/obs"] = s["obs"]
    # reference tracking + per-environment parameters + reward / terminated values
    base = O.Oracle("cart_pole", dtype=np.float32)
    rng = np.random.default_rng(8)
    static = {k: (np.float64(v) * rng.uniform(0.9, 1.1, B)).astype(np.float32) for k, v in base.static.items()}
    orc = O.Oracle("cart_pole", solver=O.TSIT5, dtype=np.float32, static_params=static, tau=1e-3,
                   control_state=["theta", "velocity"])
    st, refl = random_state(base, B, rng, 0.4), random_state(base, B, rng, 0.4)
    acts = rng.uniform(-1, 1, (B, T, 1)).astype(np.float32)
    s = orc.sim_ahead(st, acts, 1e-3, 1e-3, ref=refl, epilogue=True)
    key = "tracking_cart_pole/f32"
    for k, v in static.items():
        out[f"{key}/static/{k}"] = v
    for i in range(4):
        out[f"{key}/state{i}"], out[f"{key}/ref{i}"] = st[i], refl[i]
    out[f"{key}/actions"], out[f"{key}/obs"] = acts, s["obs"]
    out[f"{key}/reward"], out[f"{key}/terminated"], out[f"{key}/truncated"] = s["reward"], s["terminated"], s["truncated"]
    np.savez_compressed(OUT, **out)
    print(f"{OUT}: {len(out)} arrays, {OUT.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
