"""jax.ffi registration of the C-ABI kernels (BASELINE.json north_star: "registered as a JAX FFI custom call").

JAX is NOT installable in this image (no wheel, no network), so this module cannot be exercised here; it is the
binding a maintainer of the reference adds (INTEGRATION.md walks through it).  It needs
  * jax >= 0.4.38 (jax.ffi), a CUDA jaxlib, and
  * libexenv_b200_ffi.so = csrc/xla_ffi_shim.cc compiled against jax.ffi.include_dir() and linked to libexenv_b200.so.
Nothing in the product's torch-facing path imports this file.

One wrapper per handler of csrc/xla_ffi_shim.cc; each keeps the reference's leaf order:

  ffi_vmap_step        <- CoreEnvironment.vmap_step            core_env.py:533-569  (+ gym epilogue, gym_wrapper.py:117-126)
  ffi_vmap_sim_ahead   <- CoreEnvironment.vmap_sim_ahead       core_env.py:571-616  (+ vmap_generate_rew_trunc_term_ahead)
  ffi_reset_pack       <- vmap_reset packing / vmap_generate_state_from_observation  core_env.py:664-705
  ffi_epilogue         <- vmap(generate_reward / truncated / terminated)             gym_wrapper.py:117-126
  ffi_gym_epilogue     <- GymWrapper.gym_step after its vmap_step (update_ref + epilogue)  gym_wrapper.py:111-126, 170-192
"""
from __future__ import annotations

import ctypes as C

from . import _lib

TARGETS = {"exenv_step": "ExenvStepFfi", "exenv_sim_ahead": "ExenvSimAheadFfi", "exenv_reset_pack": "ExenvResetPackFfi",
           "exenv_epilogue": "ExenvEpilogueFfi", "exenv_gym_epilogue": "ExenvGymEpilogueFfi"}


def available() -> bool:
    try:
        import jax  # noqa: F401
        import jax.ffi  # noqa: F401

        return (_lib.LIB_PATH.parent / "libexenv_b200_ffi.so").exists()
    except Exception:
        return False


def register():
    """Registers the custom-call targets on the CUDA platform."""
    import jax

    shim = C.CDLL(str(_lib.LIB_PATH.parent / "libexenv_b200_ffi.so"))
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(shim, symbol)), platform="CUDA")


def _dims(env):
    D = env.physical_state_dim
    node = env.N_ODE if env._solver_id == _lib.TSIT5 else 0
    return D, node


def ffi_vmap_step(env, plan_handle: int, state_leaves, action, ref_leaves=(), epilogue=False):
    """`CoreEnvironment.vmap_step` as one custom call: (obs, new_state_leaves, k_leaves[, reward, truncated, terminated])."""
    import jax
    import jax.numpy as jnp

    B, dt = action.shape[0], action.dtype
    D, node = _dims(env)
    outs = [jax.ShapeDtypeStruct((B, env.obs_dim), dt)] + [jax.ShapeDtypeStruct((B,), dt)] * (D + node)
    if epilogue:
        outs += [jax.ShapeDtypeStruct((B, 1), dt), jax.ShapeDtypeStruct((B, env.n_trunc), jnp.bool_),
                 jax.ShapeDtypeStruct((B, 1), jnp.bool_)]
    res = jax.ffi.ffi_call("exenv_step", outs, vmap_method="sequential")(
        *state_leaves, action, *ref_leaves, plan=int(plan_handle),
        n_phys=D, n_ode=node, n_control=len(env.control_state), epilogue=int(bool(epilogue)))
    out = (res[0], res[1:1 + D], res[1 + D:1 + D + node])
    return out + tuple(res[1 + D + node:]) if epilogue else out


def ffi_vmap_sim_ahead(env, plan_handle: int, state_leaves, actions, obs_stepsize, action_stepsize, ref_leaves=(),
                       layout="time_major", epilogue=False):
    """`CoreEnvironment.vmap_sim_ahead` as one custom call: (observations, trajectory leaves, last leaves, k_last
    [, reward, truncated, terminated]).

    ``layout="time_major"`` declares the physical layouts of the trajectory operands on the custom call (major-to-minor
    axis orders): the arrays keep the reference's logical shapes, XLA relayouts only where a consumer needs the default
    layout.  The time schedule of the call shape is built and uploaded HERE, at trace time
    (``exenv_sim_ahead_prepare``), so that the handler never allocates inside the executable.
    [recall of the public jax.ffi API -- unverified in this image]"""
    import jax
    import jax.numpy as jnp

    B, T, dt = actions.shape[0], actions.shape[1], actions.dtype
    D, node = _dims(env)
    ns = C.c_int64()
    _lib.check(_lib.lib().exenv_sim_ahead_prepare(C.c_void_p(int(plan_handle)), T, float(obs_stepsize),
                                                  float(action_stepsize), 0, C.byref(ns), None))
    ns = ns.value
    outs = ([jax.ShapeDtypeStruct((B, ns, env.obs_dim), dt)] + [jax.ShapeDtypeStruct((B, ns), dt)] * D
            + [jax.ShapeDtypeStruct((B,), dt)] * (D + node))
    if epilogue:
        outs += [jax.ShapeDtypeStruct((B, ns - 1, 1), dt), jax.ShapeDtypeStruct((B, ns, env.n_trunc), jnp.bool_),
                 jax.ShapeDtypeStruct((B, ns - 1, 1), jnp.bool_)]
    tm = layout == "time_major"
    kw = {}
    if tm:
        kw = dict(input_layouts=[None] * D + [(1, 0, 2)] + [None] * len(ref_leaves),
                  output_layouts=[(1, 0, 2)] + [(1, 0)] * D + [None] * (D + node) + ([(1, 0, 2)] * 3 if epilogue else []))
    res = jax.ffi.ffi_call("exenv_sim_ahead", outs, vmap_method="sequential", **kw)(
        *state_leaves, actions, *ref_leaves, plan=int(plan_handle), n_phys=D, n_ode=node,
        n_control=len(env.control_state), obs_stepsize=float(obs_stepsize), action_stepsize=float(action_stepsize),
        max_steps=int(env.max_steps or 0), layout=_lib.LAYOUT_TIME_MAJOR if tm else _lib.LAYOUT_BATCH_MAJOR,
        epilogue=int(bool(epilogue)))
    out = (res[0], res[1:1 + D], res[1 + D:1 + 2 * D], res[1 + 2 * D:1 + 2 * D + node])
    return out + tuple(res[1 + 2 * D + node:]) if epilogue else out


def ffi_reset_pack(env, plan_handle: int, mode: int, values=None, dtype=None):
    """`vmap_reset` state packing (mode RESET_DEFAULT / RESET_FROM_NORM with the (B, D) uniform draws made by
    ``jax.random`` in the caller) and `vmap_generate_state_from_observation` (RESET_FROM_OBS):
    (state leaves, reference leaves, obs)."""
    import jax
    import jax.numpy as jnp

    B = env.batch_size
    dt = dtype or (values.dtype if values is not None else jnp.float32)
    D, _ = _dims(env)
    outs = [jax.ShapeDtypeStruct((B,), dt)] * (2 * D) + [jax.ShapeDtypeStruct((B, env.obs_dim), dt)]
    args = () if values is None else (values,)
    res = jax.ffi.ffi_call("exenv_reset_pack", outs, vmap_method="sequential")(
        *args, plan=int(plan_handle), n_phys=D, mode=int(mode), batch=int(B))
    return res[:D], res[D:2 * D], res[2 * D]


def ffi_epilogue(env, plan_handle: int, state_leaves, ref_leaves=()):
    """reward / truncated / terminated of materialised states: leaves of any shape (flattened by the kernel):
    (reward (n,), truncated (n, n_trunc), terminated (n,))."""
    import jax
    import jax.numpy as jnp

    n, dt = state_leaves[0].size, state_leaves[0].dtype
    D, _ = _dims(env)
    outs = [jax.ShapeDtypeStruct((n,), dt), jax.ShapeDtypeStruct((n, env.n_trunc), jnp.bool_),
            jax.ShapeDtypeStruct((n,), jnp.bool_)]
    return jax.ffi.ffi_call("exenv_epilogue", outs, vmap_method="sequential")(
        *state_leaves, *ref_leaves, plan=int(plan_handle), n_phys=D, n_control=len(env.control_state))


def ffi_gym_epilogue(env, plan_handle: int, state_leaves, ref_leaves, hold_steps, rng, hold_steps_min, hold_steps_max,
                     force_new=False):
    """`GymWrapper.gym_step` after its `vmap_step`: reference generator (hold counters, device-side counter RNG) +
    epilogue against the updated reference.  ``hold_steps`` (B,) int32 and ``rng`` (B, 2) uint64 are updated in place
    through ``input_output_aliases``, like the tracked reference leaves:
    (new reference leaves, hold_steps, rng, reward (B, 1), truncated (B, n_trunc), terminated (B, 1))."""
    import jax
    import jax.numpy as jnp

    B, dt = state_leaves[0].shape[0], state_leaves[0].dtype
    D, _ = _dims(env)
    outs = ([jax.ShapeDtypeStruct((B,), dt)] * D
            + [jax.ShapeDtypeStruct((B,), jnp.int32), jax.ShapeDtypeStruct((B, 2), jnp.uint64),
               jax.ShapeDtypeStruct((B, 1), dt), jax.ShapeDtypeStruct((B, env.n_trunc), jnp.bool_),
               jax.ShapeDtypeStruct((B, 1), jnp.bool_)])
    aliases = {D + d: d for d in range(D)}  # reference leaf d (argument D + d) -> result d
    aliases.update({2 * D: D, 2 * D + 1: D + 1})
    res = jax.ffi.ffi_call("exenv_gym_epilogue", outs, vmap_method="sequential", input_output_aliases=aliases)(
        *state_leaves, *ref_leaves, hold_steps, rng, plan=int(plan_handle), n_phys=D,
        hold_steps_min=int(hold_steps_min), hold_steps_max=int(hold_steps_max), force_new=int(bool(force_new)))
    return res[:D], res[D], res[D + 1], res[D + 2], res[D + 3], res[D + 4]
