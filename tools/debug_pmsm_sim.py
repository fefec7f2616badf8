import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import exciting_environments_b200 as ex
from oracle import oracle as O
from helpers import random_state
from gpu_helpers import make_env, to_state, np_

for (sat, motor, solver, dtype) in [(False, None, O.EULER, np.float32), (True, "BRUSA", O.EULER, np.float64), (True, "BRUSA", O.TSIT5, np.float64), (True, "SEW", O.EULER, np.float64)]:
    rng = np.random.default_rng(42)
    orc = O.Oracle("pmsm", solver=solver, dtype=dtype, saturated=sat, motor=motor)
    B, T = 300, 200
    st = random_state(orc, B, rng, spread=0.5)
    acts = rng.uniform(-1, 1, (B, T, 2)).astype(dtype)
    ref = orc.sim_ahead(st, acts, orc.tau, orc.tau)
    e = make_env(ex, orc, B)
    obs, states, last = e.vmap_sim_ahead(to_state(e, st), torch.as_tensor(acts, device="cuda"), orc.tau, orc.tau)
    print("case", sat, motor, solver, dtype.__name__)
    for i, n in enumerate(e.PHYS_FIELDS):
        a = np_(getattr(states.physical_state, n)).astype(np.float64); b = ref["traj"][i].astype(np.float64)
        err = np.abs(a - b)
        t_first = np.argmax(err.max(axis=0) > 1e-6 * (1 + np.abs(b).max()))
        print(f"  {n:12s} max abs err {err.max():.3e}  scale {np.abs(b).max():.3e}  first bad t {t_first}  worst env {np.argmax(err.max(axis=1))} nan {np.isnan(a).sum()}")
    a = np_(obs).astype(np.float64); b = ref["obs"].astype(np.float64)
    print("  obs err per column", np.abs(a - b).max(axis=(0, 1)))
