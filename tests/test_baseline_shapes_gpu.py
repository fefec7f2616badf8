"""`-m gpu`: parity at the BASELINE.json shapes.  Each case runs the FULL batch of the configuration on the GPU, in both
memory layouts, and compares a random 2048-environment subset with the oracle over the FULL horizon (the oracle
finishes 2048 environments in seconds); truncation flags are checked bit for bit against the reference expression on
the returned states.  The measured errors (normalised, true relative, ulp) are appended to
``gpurun_out/r02_parity_errors.jsonl`` (copied to ``profiles/`` after a GPU run) so that the slack of the tolerances is
visible.

  configs[1]  MASS_SPRING_DAMPER  B = 2^20, Tsit5 vmap_sim_ahead, 2000 steps, fp32 and fp64
  configs[2]  CART_POLE, ACROBOT  B = 2^22, Euler vmap_step x 200 and Tsit5 vmap_sim_ahead, 200 steps
  configs[3]  PMSM (saturated BRUSA flux maps)  B = 2^17 (the per-GPU shard of 2^20 on 8 GPUs), Tsit5 vmap_sim_ahead
              with the fused epilogue, 5000 steps as two 2500-step chunks (diffrax stops at max_steps = 4096)

Tolerances (north star): non-chaotic environments 1e-4 relative over the horizon in fp32 (1e-9 in fp64), errors
normalised per environment by max(half-span of the field, max_t |x|); chaotic cart-pole / acrobot over 200 steps are
bounded by 20x the trajectory's own sensitivity to a one-ulp perturbation of the initial state (measured with the
oracle), never tighter than the non-chaotic bound.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from gpu_helpers import make_env, np_, to_state
from helpers import field_scales, flags_from_states, random_state, traj_err, ulp_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
N_SUB = 2048
HORIZON_TOL = {np.float64: 1e-9, np.float32: 1e-4}


def _report(**kw):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "r02_parity_errors.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def _true_rel(a, b, scale):
    """max |a - b| / |b| over the samples whose magnitude is at least 1e-3 of the field's half-span (a state crossing zero
    has no meaningful relative error)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.abs(b) >= 1e-3 * scale
    return float(np.max(np.abs(a - b)[m] / np.abs(b)[m])) if m.any() else 0.0


ANGLES = {"pendulum": ("theta",), "cart_pole": ("theta",), "acrobot": ("theta_1", "theta_2"), "pmsm": ("epsilon",)}


def _unwrap_to(got, ref, period):
    """Angles are stored wrapped to [-pi, pi): two trajectories that differ by rounding can sit on opposite sides of the
    wrap (+pi - eps vs -pi + eps are the same angle).  Compare `got` moved by whole periods next to `ref`."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return got - period * np.round((got - ref) / period)


def _errors(name, got, ref, scale, dtype, period=None):
    if period is not None:
        got = _unwrap_to(got, ref, period)
    return dict(field=name, normalised=traj_err(got, ref, scale), true_relative=_true_rel(got, ref, scale),
                max_abs=float(np.max(np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)))),
                ulp_at_final_sample=ulp_err(got[:, -1], ref[:, -1], dtype))


def _sens_tol(orc, st, acts, dtype, ref, scales, steps=None):
    eps = np.finfo(dtype).eps
    st2 = [(x * (1 + eps)).astype(dtype) for x in st]
    pert = orc.sim_ahead(st2, acts, orc.tau, orc.tau)
    sens = max(traj_err(p, r, s) for p, r, s in zip(pert["traj"], ref["traj"], scales))
    return max(HORIZON_TOL[dtype], 20 * sens), sens


def _full_batch_inputs(orc, B, T, dtype, seed, act_scale=1.0, gentle=False):
    rng = np.random.default_rng(seed)
    st = random_state(orc, B, rng, spread=0.5, pmsm_gentle=gentle)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    g = torch.Generator(device="cuda").manual_seed(seed)
    # physical (T, B, A): the time-major layout consumes it as it is, the batch-major one gets a transposed copy
    acts_tm = (torch.rand((T, B, orc.A), device="cuda", dtype=tdt, generator=g) * 2 - 1) * act_scale
    idx = np.sort(rng.choice(B, N_SUB, replace=False))
    return st, acts_tm, idx


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("layout", ["time_major", "batch_major"])
def test_config1_mass_spring_damper_full_batch_against_oracle(excenvs, layout, dtype):
    """BASELINE configs[1] at its full size, the headline kernel (time-major fp32: simtm_packed_kernel, four
    environments per thread) and the reference-contiguous layout."""
    B, T = 1 << 20, 2000
    orc = O.Oracle("mass_spring_damper", solver=O.TSIT5, dtype=dtype)
    st, acts_tm, idx = _full_batch_inputs(orc, B, T, dtype, seed=20)
    e = make_env(excenvs, orc, B)
    s0 = to_state(e, st)
    acts = acts_tm.permute(1, 0, 2)  # logical (B, T, A)
    if layout == "batch_major":
        acts = acts.contiguous()
        del acts_tm
    obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(s0, acts, orc.tau, orc.tau, layout=layout)
    assert obs.shape == (B, T + 1, 2) and trunc.shape == (B, T + 1, 2) and rew.shape == (B, T, 1)
    tix = torch.as_tensor(idx, device="cuda")
    sub_acts = np_(acts[tix])
    ref = orc.sim_ahead([x[idx] for x in st], sub_acts, orc.tau, orc.tau, epilogue=True)
    tol = HORIZON_TOL[dtype]
    scales = field_scales(orc)
    rows = []
    own = []
    for name, r, sc in zip(e.PHYS_FIELDS, ref["traj"], scales):
        got = np_(getattr(states.physical_state, name)[tix])
        own.append(got)
        rows.append(_errors(name, got, r, sc, dtype))
        assert rows[-1]["normalised"] <= tol, rows[-1]
        assert np.array_equal(np_(getattr(last.physical_state, name)[tix]), got[:, -1])
    got_obs = np_(obs[tix])
    for o in range(2):
        rows.append(_errors(f"obs[{o}]", got_obs[..., o], ref["obs"][..., o], 1.0, dtype))
        assert rows[-1]["normalised"] <= tol, rows[-1]
    # flags: bit exact on the returned states (subset on the host, and the WHOLE batch on the device with the same
    # IEEE expression: torch's `/` on CUDA is a true division for tensor / tensor)
    tr_sub = np_(trunc[tix]).astype(np.uint8)
    assert np.array_equal(tr_sub, flags_from_states(orc, own))
    lo = torch.tensor(-10.0, dtype=obs.dtype, device="cuda")
    span = torch.tensor(10.0, dtype=obs.dtype, device="cuda") - lo
    for i, name in enumerate(e.PHYS_FIELDS):
        t = getattr(states.physical_state, name) - lo  # 2 * (x - min) / (max - min) - 1, one temporary
        t.mul_(2).div_(span).sub_(1).abs_()
        assert torch.equal(trunc[..., i], t > 1), name
        del t
    assert bool(term.all()) and bool((rew == 0).all())  # control_state == []: reward 0, terminated = (reward == 0)
    _report(case="configs[1] mass_spring_damper Tsit5 sim_ahead", B=B, T=T, dtype=np.dtype(dtype).name, layout=layout,
            subset=N_SUB, tolerance=tol, errors=rows,
            flags="bit-exact: subset vs NumPy and full batch vs the IEEE expression on device")


@pytest.mark.parametrize("env", ["cart_pole", "acrobot"])
def test_config2_chaotic_envs_full_batch_against_oracle(excenvs, env):
    """BASELINE configs[2]: B = 2^22, 200-step horizon, Tsit5 vmap_sim_ahead in both layouts and 200 Euler vmap_step calls."""
    B, T, dtype = 1 << 22, 200, np.float32
    orc = O.Oracle(env, solver=O.TSIT5, dtype=dtype)
    st, acts_tm, idx = _full_batch_inputs(orc, B, T, dtype, seed=30)
    tix = torch.as_tensor(idx, device="cuda")
    acts = acts_tm.permute(1, 0, 2)
    sub_acts = np_(acts[tix])
    sub_st = [x[idx] for x in st]
    ref = orc.sim_ahead(sub_st, sub_acts, orc.tau, orc.tau, epilogue=True)
    scales = field_scales(orc)
    tol, sens = _sens_tol(orc, sub_st, sub_acts, dtype, ref, scales)
    e = make_env(excenvs, orc, B)
    s0 = to_state(e, st)
    for layout in ("time_major", "batch_major"):
        a = acts if layout == "time_major" else acts.contiguous()
        obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(s0, a, orc.tau, orc.tau, layout=layout)
        rows, own = [], []
        for name, r, sc in zip(e.PHYS_FIELDS, ref["traj"], scales):
            got = np_(getattr(states.physical_state, name)[tix])
            own.append(got)
            rows.append(_errors(name, got, r, sc, dtype, 2 * np.pi if name in ANGLES[env] else None))
            assert rows[-1]["normalised"] <= tol, (layout, rows[-1], tol)
        assert np.array_equal(np_(trunc[tix]).astype(np.uint8), flags_from_states(orc, own)), layout
        _report(case=f"configs[2] {env} Tsit5 sim_ahead", B=B, T=T, dtype="float32", layout=layout, subset=N_SUB,
                tolerance=tol, one_ulp_sensitivity=sens, errors=rows, flags="bit-exact vs NumPy on the returned states")
        del obs, states, last, rew, trunc, term, a
    # Euler vmap_step, closed over 200 calls on the full batch
    orc_e = O.Oracle(env, solver=O.EULER, dtype=dtype)
    ee = make_env(excenvs, orc_e, B)
    state = to_state(ee, st)
    cur = [x.copy() for x in sub_st]
    worst = 0.0
    for t in range(T):
        obs, state, rew, trunc, term = ee.vmap_step_with_epilogue(state, acts_tm[t])
        if t % 20 == 19 or t == T - 1:  # the oracle is stepped every step, compared at these
            pass
        r = orc_e.step(cur, sub_acts[:, t], epilogue=True)
        cur = r["state"]
        if t % 20 == 19 or t == T - 1:
            got = [np_(getattr(state.physical_state, n)[tix]) for n in ee.PHYS_FIELDS]
            for n, g, b, sc in zip(ee.PHYS_FIELDS, got, cur, scales):
                g = _unwrap_to(g, b, 2 * np.pi) if n in ANGLES[env] else g.astype(np.float64)
                worst = max(worst, float(np.max(np.abs(g - b) / np.maximum(np.abs(b), sc))))
            assert np.array_equal(np_(trunc[tix]).astype(np.uint8), flags_from_states(orc_e, got))
    assert worst <= tol, (worst, tol)
    _report(case=f"configs[2] {env} Euler vmap_step x200 (closed loop)", B=B, T=T, dtype="float32", subset=N_SUB,
            tolerance=tol, errors=[dict(field="all", normalised=worst)], flags="bit-exact vs NumPy on the returned states")


@pytest.mark.parametrize("layout", ["time_major", "batch_major"])
@pytest.mark.parametrize("act_scale,gentle", [(0.02, True), (1.0, False)])
def test_config4_pmsm_saturated_two_chunks_against_oracle(excenvs, layout, act_scale, gentle):
    """BASELINE configs[3] ("config 4" of SURVEY 8d): PMSM with the saturated BRUSA flux maps, Tsit5, fused epilogue,
    5000 steps as the canonical 2 x 2500-step chunks (the second chunk starts from the first one's last_state, like a
    reference user would have to do under diffrax's max_steps = 4096), on the per-GPU shard of 2^17 environments.
    Gentle inputs (low speed, 2 % voltages: currents stay inside the flux maps) are held to the non-chaotic 1e-4
    bound; full-range random voltages drive the currents far outside the maps (u / r_s ~ 15 kA) and are held to the
    sensitivity-scaled bound."""
    B, T, CH, dtype = 1 << 17, 5000, 2500, np.float32
    orc = O.Oracle("pmsm", solver=O.TSIT5, dtype=dtype, saturated=True, motor="BRUSA")
    st, acts_tm, idx = _full_batch_inputs(orc, B, T, dtype, seed=40, act_scale=act_scale, gentle=gentle)
    tix = torch.as_tensor(idx, device="cuda")
    acts = acts_tm.permute(1, 0, 2)
    sub_acts = np_(acts[tix])
    scales = field_scales(orc)
    e = make_env(excenvs, orc, B)
    state = to_state(e, st)
    cur = [x[idx] for x in st]
    tol_all, rows_all = [], []
    for c in range(T // CH):
        a = acts[:, c * CH:(c + 1) * CH]
        if layout == "batch_major":
            a = a.contiguous()
        obs, states, last, rew, trunc, term = e.vmap_sim_ahead_with_epilogue(state, a, orc.tau, orc.tau, layout=layout)
        ref = orc.sim_ahead(cur, sub_acts[:, c * CH:(c + 1) * CH], orc.tau, orc.tau, epilogue=True)
        if gentle:
            tol, sens = HORIZON_TOL[dtype], None
        else:
            tol, sens = _sens_tol(orc, cur, sub_acts[:, c * CH:(c + 1) * CH], dtype, ref, scales)
        # The electrical angle integrates the constant omega_el whatever the actions are, so its fp32 accumulation noise
        # can be measured directly: the fp32 oracle against the fp64 oracle (the other fields cannot be compared across
        # precisions: the zero-order-hold index lags differently in fp32 and fp64).  After 2500 steps that
        # noise is ~1.5e-4 of pi -- the bound for epsilon is twice the reference formula's own distance from fp64.
        ref64 = O.Oracle("pmsm", solver=O.TSIT5, dtype=np.float64, saturated=True, motor="BRUSA").sim_ahead(
            [np.asarray(x, np.float64) for x in cur], sub_acts[:, c * CH:(c + 1) * CH].astype(np.float64), orc.tau, orc.tau)
        eps_noise = traj_err(_unwrap_to(ref["traj"][2], ref64["traj"][2], 2 * np.pi), ref64["traj"][2], scales[2])
        own = {}
        for name, r, sc in zip(e.PHYS_FIELDS, ref["traj"], scales):
            got = np_(getattr(states.physical_state, name)[tix])
            own[name] = got
            row = _errors(name, got, r, sc, dtype, 2 * np.pi if name == "epsilon" else None)
            row["chunk"] = c
            bound = max(tol, 2 * eps_noise) if name == "epsilon" else tol
            if name == "epsilon":
                row["fp32_reference_formula_vs_fp64"] = eps_noise
            rows_all.append(row)
            assert row["normalised"] <= bound, (c, row, bound)
        # observations: cos(eps), sin(eps) (continuous across the wrap) carry the angle's accumulation noise in radians
        got_obs = np_(obs[tix])
        obs_bound = max(2 * tol, 2.1 * eps_noise * np.pi)
        assert traj_err(got_obs.reshape(N_SUB, -1), ref["obs"].reshape(N_SUB, -1), 1.0) <= obs_bound
        # flags: the reference expression on the returned states, subset on the host + consistency on the full batch
        leaves = [own[n] if n in ("i_d", "i_q") else None for n in e.PHYS_FIELDS]
        assert np.array_equal(np_(trunc[tix]).astype(np.uint8), flags_from_states(orc, leaves))
        assert torch.equal(term, trunc[:, 1:]) and bool((rew == 0).all())
        tol_all.append(tol)
        # next chunk: the kernel continues from ITS last state, the oracle from its own (as two reference runs would)
        state = to_state(e, [np_(getattr(last.physical_state, n)) for n in e.PHYS_FIELDS])
        cur = ref["last"]
        del obs, states, last, rew, trunc, term
    _report(case="configs[3] pmsm saturated (BRUSA) Tsit5 sim_ahead + epilogue, 2 x 2500 steps", B=B, T=T,
            dtype="float32", layout=layout, inputs="gentle" if gentle else "full-range random voltages",
            subset=N_SUB, tolerance=tol_all, errors=rows_all, flags="bit-exact vs NumPy on the returned states")
