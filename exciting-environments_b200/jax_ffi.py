"""jax.ffi registration of the C-ABI kernels (BASELINE.json north_star: "registered as a JAX FFI custom call").

JAX is NOT installable in this image (no wheel, no network), so this module cannot be exercised here; it is the
binding a maintainer of the reference adds (INTEGRATION.md walks through it).  It needs
  * jax >= 0.4.38 (jax.ffi), a CUDA jaxlib, and
  * libexenv_b200_ffi.so = csrc/xla_ffi_shim.cc compiled against jax.ffi.include_dir() and linked to libexenv_b200.so.
Nothing in the product's torch-facing path imports this file.
"""
from __future__ import annotations

import ctypes as C

from . import _lib


def available() -> bool:
    try:
        import jax  # noqa: F401
        import jax.ffi  # noqa: F401

        return (_lib.LIB_PATH.parent / "libexenv_b200_ffi.so").exists()
    except Exception:
        return False


def register():
    """Registers the two custom-call targets on the CUDA platform."""
    import jax

    shim = C.CDLL(str(_lib.LIB_PATH.parent / "libexenv_b200_ffi.so"))
    jax.ffi.register_ffi_target("exenv_step", jax.ffi.pycapsule(shim.ExenvStepFfi), platform="CUDA")
    jax.ffi.register_ffi_target("exenv_sim_ahead", jax.ffi.pycapsule(shim.ExenvSimAheadFfi), platform="CUDA")


def ffi_vmap_step(env, plan_handle: int, state_leaves, action, ref_leaves=()):
    """`CoreEnvironment.vmap_step` as one custom call: returns (obs, new_state_leaves, k_leaves)."""
    import jax
    import jax.numpy as jnp

    B, dt = action.shape[0], action.dtype
    D, node = env.physical_state_dim, (env.N_ODE if env._solver_id == _lib.TSIT5 else 0)
    outs = [jax.ShapeDtypeStruct((B, env.obs_dim), dt)] + [jax.ShapeDtypeStruct((B,), dt)] * (D + node)
    res = jax.ffi.ffi_call("exenv_step", outs, vmap_method="sequential")(
        *state_leaves, action, *ref_leaves, plan=int(plan_handle),
        n_phys=D, n_ode=node, n_control=len(env.control_state))
    return res[0], res[1:1 + D], res[1 + D:]


def ffi_vmap_sim_ahead(env, plan_handle: int, state_leaves, actions, obs_stepsize, action_stepsize, ref_leaves=(),
                       layout="time_major"):
    """`CoreEnvironment.vmap_sim_ahead` as one custom call: (observations, trajectory leaves, last leaves, k_last).

    ``layout="time_major"`` declares the physical layouts of the trajectory operands on the custom call (major-to-minor
    axis orders): the arrays keep the reference's logical shapes, XLA relayouts only where a consumer needs the default
    layout.  [recall of the public jax.ffi API -- unverified in this image]"""
    import jax

    B, T, dt = actions.shape[0], actions.shape[1], actions.dtype
    D, node = env.physical_state_dim, (env.N_ODE if env._solver_id == _lib.TSIT5 else 0)
    ns = 1 + int(action_stepsize * T / obs_stepsize)
    outs = ([jax.ShapeDtypeStruct((B, ns, env.obs_dim), dt)] + [jax.ShapeDtypeStruct((B, ns), dt)] * D
            + [jax.ShapeDtypeStruct((B,), dt)] * (D + node))
    tm = layout == "time_major"
    kw = {}
    if tm:
        kw = dict(input_layouts=[None] * D + [(1, 0, 2)] + [None] * len(ref_leaves),
                  output_layouts=[(1, 0, 2)] + [(1, 0)] * D + [None] * (D + node))
    res = jax.ffi.ffi_call("exenv_sim_ahead", outs, vmap_method="sequential", **kw)(
        *state_leaves, actions, *ref_leaves, plan=int(plan_handle), n_phys=D, n_ode=node,
        n_control=len(env.control_state), obs_stepsize=float(obs_stepsize), action_stepsize=float(action_stepsize),
        max_steps=int(env.max_steps or 0), layout=_lib.LAYOUT_TIME_MAJOR if tm else _lib.LAYOUT_BATCH_MAJOR)
    return res[0], res[1:1 + D], res[1 + D:1 + 2 * D], res[1 + 2 * D:]
