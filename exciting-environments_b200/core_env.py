"""Host-side mirror of ``exciting_environments.core_env.CoreEnvironment`` (reference core_env.py:15-705).

Same class surface (``batch_size``, ``tau``, ``env_properties``, ``in_axes_env_properties``, ``action_dim``,
``physical_state_dim``, nested ``State`` / ``PhysicalState`` / ``Action`` / ``StaticParams`` / ``Additions`` /
``EnvProperties`` containers, ``vmap_reset`` / ``vmap_step`` / ``vmap_sim_ahead`` /
``vmap_generate_state_from_observation`` and the single-environment ``reset`` / ``step`` / ``sim_ahead``), but
every numeric operation is one call through the C ABI of ``libexenv_b200.so`` into a hand-written sm_100a kernel.
Arrays are ``torch`` CUDA tensors (torch is only the owner of device memory and streams here; JAX is not installable
in this image -- ``jax_ffi.py`` holds the ``jax.ffi`` registration for when it is).

Differences a reference user will notice (documented in DESIGN.md):
  * dtype is chosen per environment (``dtype=torch.float32`` / ``torch.float64`` or ``config.update("enable_x64")``)
    instead of the global ``jax_enable_x64`` flag;
  * ``rng`` arguments are ``torch.Generator`` objects / integer seeds (JAX's threefry keys cannot be reproduced
    without JAX): random resets have the reference's distribution, not its bit pattern;
  * there is no CPU path: without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, fields, make_dataclass
from typing import Callable

import torch

from . import _lib
from .solvers import Euler, solver_id
from .utils import MinMaxNormalization, is_array, tree_flatten, tree_structure, tree_unflatten


class _Config:
    """Stand-in for ``jax.config`` (only ``enable_x64`` / ``jax_enable_x64`` is meaningful here)."""

    enable_x64 = False

    def update(self, name, value):
        if name in ("enable_x64", "jax_enable_x64"):
            self.enable_x64 = bool(value)
        elif name in ("jax_platform_name",):
            pass
        else:
            raise AttributeError(f"unknown config option {name}")


config = _Config()


def _mk(name, field_names):
    return make_dataclass(name, [(f, object) for f in field_names])


@dataclass
class _State:
    """The state of the environment (reference core_env.py:236-243)."""

    physical_state: object
    PRNGKey: object
    additions: object
    reference: object


@dataclass
class _EnvProperties:
    """Properties that stay constant during simulation (reference core_env.py:245-251)."""

    physical_normalizations: object
    action_normalizations: object
    static_params: object


@dataclass
class _Additions:
    solver_state: object
    active_solver_state: object


class CoreEnvironment:
    # ---- per-environment metadata, filled by the subclasses in envs.py -------------------------------------------
    ENV_ID: int = -1
    PHYS_FIELDS: tuple = ()
    ACTION_FIELDS: tuple = ()
    STATIC_FIELDS: tuple = ()
    N_ODE: int = 0
    N_OBS: int = 0
    ODE_FIELDS: tuple = ()  # PhysicalState fields that form the ODE state, in solver order
    # PMSM's _ode_solver_simulate_ahead returns an all-NaN reference whatever init_state.reference holds
    # (pmsm_env.py:688-701), so observations' *_ref columns, rewards and last_state.reference of a PMSM sim_ahead are
    # NaN upstream; reproduced (SURVEY Appendix C style quirk, pinned by tests/test_sim_gpu.py)
    SIM_AHEAD_NAN_REFERENCE: bool = False

    State = _State
    EnvProperties = _EnvProperties
    Additions = _Additions

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if cls.PHYS_FIELDS:
            cls.PhysicalState = _mk("PhysicalState", cls.PHYS_FIELDS)
            cls.Action = _mk("Action", cls.ACTION_FIELDS)
            cls.StaticParams = _mk("StaticParams", cls.STATIC_FIELDS)

    def __init__(self, batch_size: int, env_properties, tau: float = 1e-4, solver=None, dtype=None, device=None):
        self.batch_size = int(batch_size)
        self.tau = tau
        self._solver = solver if solver is not None else Euler()
        self._solver_id = solver_id(self._solver)
        self.env_properties = env_properties
        self.in_axes_env_properties = self.create_in_axes_dataclass(env_properties)
        self.action_dim = len(fields(self.Action))
        self.physical_state_dim = len(fields(self.PhysicalState))
        if dtype is None:
            dtype = torch.float64 if config.enable_x64 else torch.float32
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.dtype = dtype
        self._device = torch.device(device) if device is not None else None
        self._plans = {}
        self._const = {}
        self.time_rule = 0
        self.max_steps = 4096  # diffrax.diffeqsolve default; None lifts the limit

    # ---- containers -----------------------------------------------------------------------------------------------
    def create_in_axes_dataclass(self, dataclass_obj):
        """vmap axes of every property: ``None`` = shared scalar, ``0`` = per-environment ``(B,)`` array
        (reference core_env.py:253-277, same error messages)."""
        out = {}
        for f in fields(dataclass_obj):
            name, value = f.name, getattr(dataclass_obj, f.name)
            if value is None:
                out[name] = None
            elif isinstance(value, list):
                raise ValueError(
                    f'Passed env property "{name}" needs to be a jnp.array to have different setting per batch, but list is given.'
                )
            elif hasattr(value, "__dataclass_fields__"):
                out[name] = self.create_in_axes_dataclass(value)
            elif isinstance(value, (int, float, bool)) or (isinstance(value, torch.Tensor) and value.ndim == 0):
                out[name] = None
            elif isinstance(value, torch.Tensor):
                out[name] = 0 if value.shape[0] == self.batch_size else None
            elif hasattr(value, "shape") and hasattr(value, "dtype"):  # numpy array / numpy scalar
                out[name] = 0 if (len(value.shape) and value.shape[0] == self.batch_size) else None
            else:
                raise ValueError(
                    f'Passed env property "{name}" needs to be a scalar, jnp.array or jdc.pytree_dataclass, but {type(value)} is given.'
                )
        return type(dataclass_obj)(**out)

    @property
    def device(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("exciting-environments_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            self._device = torch.device("cuda", torch.cuda.current_device())
        elif self._device.index is None:  # "cuda" -> "cuda:<current>" (tensor.device always carries the index)
            self._device = torch.device(self._device.type, torch.cuda.current_device())
        return self._device

    def _stream(self):
        """The caller's current CUDA stream as a raw ``cudaStream_t`` (the C call is a few hundred ns; the public
        ``torch.cuda.current_stream`` wrapper costs several microseconds per control step)."""
        dev = self.device
        try:
            return C.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index))
        except AttributeError:  # private API moved: fall back to the public one
            return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    @property
    def obs_dim(self):
        return self.N_OBS + len(self.control_state)

    @property
    def n_trunc(self):
        return 1 if self.ENV_ID in (2, 5) else self.obs_dim

    # ---- plan (ExenvDesc) -----------------------------------------------------------------------------------------
    def _param_slots(self, env_properties):
        sp, pn, an = env_properties.static_params, env_properties.physical_normalizations, env_properties.action_normalizations
        vals = [getattr(sp, n) for n in self.STATIC_FIELDS]
        for n in self.PHYS_FIELDS:
            m = getattr(pn, n)
            vals += [m.min, m.max]
        for n in self.ACTION_FIELDS:
            m = getattr(an, n)
            vals += [m.min, m.max]
        return vals

    def _desc_extra(self, desc, keep):
        """Hook for PMSM (saturated flag, LUT)."""

    def _make_plan(self, env_properties, batch):
        desc = _lib.ExenvDesc()
        desc.env, desc.solver = self.ENV_ID, self._solver_id
        desc.dtype = _lib.F64 if self.dtype == torch.float64 else _lib.F32
        desc.tau = float(self.tau)
        desc.n_control = len(self.control_state)
        for j, name in enumerate(self.control_state):
            desc.control[j] = self.PHYS_FIELDS.index(name)
        desc.time_rule = int(self.time_rule)
        keep = []
        for i, v in enumerate(self._param_slots(env_properties)):
            if hasattr(v, "shape") and len(v.shape) >= 1 and v.shape[0] == batch:
                t = torch.as_tensor(v).to(device=self.device, dtype=self.dtype).contiguous()
                keep.append(t)
                desc.pp[i] = t.data_ptr()
                desc.p[i] = float("nan")
            else:
                desc.p[i] = float("nan") if v is None else float(v)
        self._desc_extra(desc, keep)
        handle = C.c_void_p()
        _lib.check(_lib.lib().exenv_plan_create(C.byref(desc), self.device.index or 0, C.byref(handle)))
        return _Plan(handle, keep)

    _PLAN_CACHE_SIZE = 16

    def _props_key(self, env_properties, batch):
        """Cache key of a plan by VALUE: scalar slots by their float value, per-environment arrays by identity of the
        array object (the entry keeps a strong reference, so the id cannot be recycled while the entry lives) together
        with its data pointer and shape."""
        key, keep = [], []
        for v in self._param_slots(env_properties):
            if hasattr(v, "shape") and len(v.shape) >= 1 and v.shape[0] == batch:
                ptr = v.data_ptr() if isinstance(v, torch.Tensor) else v.__array_interface__["data"][0]
                key.append(("arr", id(v), ptr, tuple(v.shape), str(v.dtype)))
                keep.append(v)
            else:
                key.append(None if v is None else float(v))
        return tuple(key), keep

    def _plan(self, env_properties=None, batch=None):
        if env_properties is None or env_properties is self.env_properties:
            env_properties, batch = self.env_properties, self.batch_size
        else:
            batch = batch or 1
        pkey, keep = self._props_key(env_properties, batch)
        key = (pkey, batch, float(self.tau), self._solver_id, self.dtype, tuple(self.control_state), self.time_rule,
               self._plan_extra_key(env_properties))
        plan = self._plans.pop(key, None)
        if plan is None:
            plan = self._make_plan(env_properties, batch)
            plan.keep_key = keep
            while len(self._plans) >= self._PLAN_CACHE_SIZE:  # least recently used first (dicts keep insertion order)
                self._plans.pop(next(iter(self._plans)))
        self._plans[key] = plan  # (re-)inserted last: most recently used
        return plan

    def _plan_extra_key(self, env_properties):
        """Hook for PMSM (saturated flag, motor LUT identity)."""
        return None

    # ---- small cached constants -----------------------------------------------------------------------------------
    def _constant(self, kind, B):
        key = (kind, B)
        if key not in self._const:
            if kind == "nan":
                t = torch.full((B,), float("nan"), dtype=self.dtype, device=self.device)
            elif kind == "true":
                t = torch.ones((B,), dtype=torch.bool, device=self.device)
            elif kind == "false":
                t = torch.zeros((B,), dtype=torch.bool, device=self.device)
            elif kind == "one":
                t = torch.ones((B,), dtype=self.dtype, device=self.device)
            self._const[key] = t
        return self._const[key]

    def _as(self, x, shape=None):
        if (type(x) is torch.Tensor and x.dtype == self.dtype and x.device == self._device and x.is_contiguous()
                and (shape is None or x.shape == shape)):
            return x  # the common case of a step loop: the leaf came out of the previous call
        t = torch.as_tensor(x)
        if t.device != self.device or t.dtype != self.dtype:
            t = t.to(device=self.device, dtype=self.dtype)
        if shape is not None and tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t.contiguous()

    def _phys_leaves(self, physical_state, B):
        return [self._as(getattr(physical_state, n), (B,)) for n in self.PHYS_FIELDS]

    def _new_state(self, phys_leaves, prng, solver_state, active, ref_leaves):
        return self.State(
            physical_state=self.PhysicalState(*phys_leaves),
            PRNGKey=prng,
            additions=self.Additions(solver_state=solver_state, active_solver_state=active),
            reference=self.PhysicalState(*ref_leaves),
        )

    def _solver_state_dummy(self, B):
        """``solver.init(...) * nan`` of the reference's init_state (e.g. pendulum_env.py:289-292): ``None`` for
        Euler, a tuple of NaN derivative leaves for Tsit5."""
        if self._solver_id == _lib.EULER:
            return None
        nan = self._constant("nan", B)
        return tuple(nan for _ in range(self.N_ODE))

    # ---- vmap_reset / state packing ---------------------------------------------------------------------------------
    def _reset_impl(self, env_properties, B, mode, values=None, state=None):
        plan = self._plan(env_properties, B)
        D = self.physical_state_dim
        a = _lib.ExenvResetArgs()
        a.batch, a.mode = B, mode
        keep = []
        if mode == _lib.RESET_FROM_STATE:
            phys = self._phys_leaves(state.physical_state, B)
            refs = [self._as(getattr(state.reference, n), (B,)) for n in self.PHYS_FIELDS] if self.control_state else [
                None] * D
        else:
            phys = [torch.empty(B, dtype=self.dtype, device=self.device) for _ in range(D)]
            refs = [torch.empty(B, dtype=self.dtype, device=self.device) for _ in range(D)]
            if values is not None:
                v = self._as(values)
                keep.append(v)
                a.values = v.data_ptr()
        obs = torch.empty((B, self.obs_dim), dtype=self.dtype, device=self.device)
        for i in range(D):
            a.state_out[i] = phys[i].data_ptr()
            a.ref_out[i] = refs[i].data_ptr() if refs[i] is not None else None
        a.obs = obs.data_ptr()
        _lib.check(_lib.lib().exenv_reset_pack(plan.handle, C.byref(a), self._stream()))
        return obs, phys, refs

    def _draw_norm_state(self, rng, B):
        """Normalised random initial state, uniform(-1, 1) per field (e.g. pendulum_env.py:271-276)."""
        gen = _as_generator(rng, self.device)
        return torch.rand((B, self.physical_state_dim), generator=gen, device=self.device, dtype=self.dtype) * 2 - 1

    def vmap_init_state(self, rng=None):
        """Initial state for all batches (reference core_env.py:649-662)."""
        return self.vmap_reset(rng)[1]

    def vmap_reset(self, rng=None, initial_state=None):
        """Resets all batches to the default, a random or a passed initial state (reference core_env.py:664-687).

        Returns ``(obs (B, O'), state)``."""
        B = self.batch_size
        if initial_state is not None:
            assert tree_structure(self._template_state()) == tree_structure(
                initial_state
            ), "initial_state should have the same dataclass structure as self.vmap_init_state()"
            obs, _, _ = self._reset_impl(self.env_properties, B, _lib.RESET_FROM_STATE, state=initial_state)
            return obs, initial_state
        if rng is None:
            obs, phys, refs = self._reset_impl(self.env_properties, B, _lib.RESET_DEFAULT)
            prng = self._constant("nan", B)
        else:
            values = self._draw_norm_state(rng, B)
            obs, phys, refs = self._reset_impl(self.env_properties, B, _lib.RESET_FROM_NORM, values=values)
            prng = _key_tensor(rng, B, self.device)
        state = self._new_state(phys, prng, self._solver_state_dummy(B), self._constant("false", B), refs)
        return obs, state

    def _template_state(self):
        nan = 0.0
        sst = None if self._solver_id == _lib.EULER else tuple(nan for _ in range(self.N_ODE))
        return self._new_state([nan] * self.physical_state_dim, nan, sst, False, [nan] * self.physical_state_dim)

    def vmap_generate_state_from_observation(self, obs, key=None):
        """State for each batch from a given observation (reference core_env.py:689-705)."""
        B = self.batch_size
        obs = self._as(obs, (B, self.obs_dim))
        _, phys, refs = self._reset_impl(self.env_properties, B, _lib.RESET_FROM_OBS, values=obs)
        prng = self._constant("nan", B) if key is None else key
        return self._new_state(phys, prng, self._solver_state_dummy(B), self._constant("false", B), refs)

    def vmap_generate_observation(self, state):
        """Observation of a batched state (``jax.vmap(env.generate_observation)`` in the reference's tests)."""
        return self._reset_impl(self.env_properties, self.batch_size, _lib.RESET_FROM_STATE, state=state)[0]

    # ---- vmap_step ---------------------------------------------------------------------------------------------------
    def _step_impl(self, env_properties, B, state, action, epilogue=False, inplace=False):
        plan = self._plan(env_properties, B)
        D, A, NODE = self.physical_state_dim, self.action_dim, self.N_ODE
        phys_in = self._phys_leaves(state.physical_state, B)
        action = self._as(action, (B, A))
        dev, dt = self.device, self.dtype
        tsit = self._solver_id == _lib.TSIT5
        # one allocation for the D new leaves (+ the Tsit5 derivative leaves): rows of a (D [+ NODE], B) block
        rows = (0 if inplace else D) + (NODE if tsit else 0)
        if rows and (B * phys_in[0].element_size()) % 16 == 0:  # rows stay 16-byte aligned for the vector path
            block = torch.empty((rows, B), dtype=dt, device=dev).unbind(0)
        else:
            block = tuple(torch.empty(B, dtype=dt, device=dev) for _ in range(rows))
        phys_out = phys_in if inplace else list(block[:D])
        obs = torch.empty((B, self.obs_dim), dtype=dt, device=dev)
        a = _lib.ExenvStepArgs()
        a.batch = B
        refs = None
        if self.control_state:
            refs = [self._as(getattr(state.reference, n), (B,)) for n in self.PHYS_FIELDS]
        for i in range(D):
            a.state_in[i] = phys_in[i].data_ptr()
            a.state_out[i] = phys_out[i].data_ptr()
            if refs:
                a.ref[i] = refs[i].data_ptr()
        a.action, a.obs = action.data_ptr(), obs.data_ptr()
        kout = None
        if tsit:
            kout = tuple(block[-NODE:])
            for i in range(NODE):
                a.k_out[i] = kout[i].data_ptr()
        rew = trunc = term = None
        if epilogue:
            rew = torch.empty((B, 1), dtype=dt, device=dev)
            trunc = torch.empty((B, self.n_trunc), dtype=torch.bool, device=dev)
            term = torch.empty((B, 1), dtype=torch.bool, device=dev)
            a.reward, a.truncated, a.terminated = rew.data_ptr(), trunc.data_ptr(), term.data_ptr()
        _lib.check(_lib.lib().exenv_step(plan.handle, C.byref(a), self._stream()))
        new_state = self.State(
            physical_state=self.PhysicalState(*phys_out),
            PRNGKey=state.PRNGKey,
            additions=self.Additions(solver_state=kout, active_solver_state=self._constant("true", B)),
            reference=state.reference,
        )
        if epilogue:
            return obs, new_state, rew, trunc, term
        return obs, new_state

    def vmap_step(self, state, action, inplace: bool = False):
        """One control step for all ``batch_size`` environments (reference core_env.py:533-569).

        ``action`` has shape ``(batch_size, action_dim)``; returns ``(obs, state)``.  ``inplace=True`` donates the
        physical-state leaves (the kernel reads each element before writing it)."""
        self._check_step_shapes(state, action, self.batch_size)
        return self._step_impl(self.env_properties, self.batch_size, state, action, inplace=inplace)

    def vmap_step_with_epilogue(self, state, action, inplace: bool = False):
        """``vmap_step`` + the GymWrapper epilogue (reward, truncated, terminated) fused in the same kernel
        (reference gym_wrapper.py:109-126).  Returns ``(obs, state, reward, truncated, terminated)``."""
        self._check_step_shapes(state, action, self.batch_size)
        return self._step_impl(self.env_properties, self.batch_size, state, action, epilogue=True, inplace=inplace)

    def _check_step_shapes(self, state, action, B):
        shape = tuple(torch.as_tensor(action).shape) if not isinstance(action, torch.Tensor) else tuple(action.shape)
        assert shape == (B, self.action_dim), (
            "The action needs to be of shape (batch_size, action_dim) which is "
            + f"{(B, self.action_dim)}, but {shape} is given"
        )
        phys = state.physical_state
        if type(phys) is self.PhysicalState:  # field count is fixed by the class: only the batch axis can be wrong
            first = getattr(phys, self.PHYS_FIELDS[0])
            ps = (tuple(first.shape)[0] if is_array(first) and len(first.shape) else 1, self.physical_state_dim)
        else:
            leaves = tree_flatten(phys)[0]
            ps = (tuple(leaves[0].shape)[0] if is_array(leaves[0]) else 1, len(leaves))
        assert ps == (B, self.physical_state_dim), (
            "The physical state needs to be of shape (batch_size, physical_state_dim) which is "
            + f"{(B, self.physical_state_dim)}, but {ps} is given"
        )

    # ---- vmap_sim_ahead ---------------------------------------------------------------------------------------------
    def sim_ahead_shape(self, n_actions, obs_stepsize, action_stepsize):
        """(number of saved samples N+1, number of solver steps) of the emulated ``diffeqsolve`` call."""
        ns, nst = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().exenv_sim_ahead_shape(self._plan().handle, n_actions, obs_stepsize, action_stepsize,
                                                    C.byref(ns), C.byref(nst)))
        return ns.value, nst.value

    def _sim_impl(self, env_properties, B, init_state, actions, obs_stepsize, action_stepsize, epilogue=False,
                  tile_steps=0, block_envs=0, out=None, layout="batch_major"):
        if layout not in ("batch_major", "time_major"):
            raise ValueError(f"layout must be 'batch_major' or 'time_major', got {layout!r}")
        tm = layout == "time_major"
        plan = self._plan(env_properties, B)
        D, A, NODE = self.physical_state_dim, self.action_dim, self.N_ODE
        dev, dt = self.device, self.dtype
        if tm:
            # physical (T, B, A): a view made by `time_major_actions` (or any tensor with these strides) is used as it
            # is; a batch-major tensor costs one transposing copy
            actions = torch.as_tensor(actions, dtype=dt, device=dev)
            actions_phys = actions.permute(1, 0, 2).contiguous()
        else:
            actions = self._as(actions)
        T = actions.shape[1]
        ns, nst = C.c_int64(), C.c_int64()
        # shape of the outputs + the schedule upload of a new call shape (so that the launch below never allocates: the
        # first call of a shape can already be captured into a CUDA graph)
        _lib.check(_lib.lib().exenv_sim_ahead_prepare(plan.handle, T, obs_stepsize, action_stepsize, tile_steps,
                                                      C.byref(ns), C.byref(nst)))
        NS = ns.value
        phys_in = self._phys_leaves(init_state.physical_state, B)
        out = out or {}
        obs = out.get("obs")
        if obs is None:
            obs = torch.empty((NS, B, self.obs_dim) if tm else (B, NS, self.obs_dim), dtype=dt, device=dev)
        traj = out.get("traj") or [torch.empty((NS, B) if tm else (B, NS), dtype=dt, device=dev) for _ in range(D)]
        last = [torch.empty(B, dtype=dt, device=dev) for _ in range(D)]
        a = _lib.ExenvSimArgs()
        a.batch, a.n_actions = B, T
        a.obs_stepsize, a.action_stepsize = obs_stepsize, action_stepsize
        a.max_steps = self.max_steps if self.max_steps else 0
        a.tile_steps, a.block_envs = tile_steps, block_envs
        a.layout = _lib.LAYOUT_TIME_MAJOR if tm else _lib.LAYOUT_BATCH_MAJOR
        refs = None
        if self.SIM_AHEAD_NAN_REFERENCE:
            init_ref = self.PhysicalState(*[self._constant("nan", B)] * D)
        else:
            init_ref = init_state.reference
        if self.control_state:
            refs = [self._as(getattr(init_ref, n), (B,)) for n in self.PHYS_FIELDS]
        for i in range(D):
            a.state_in[i] = phys_in[i].data_ptr()
            a.traj[i] = traj[i].data_ptr()
            a.last[i] = last[i].data_ptr()
            if refs:
                a.ref[i] = refs[i].data_ptr()
        a.actions, a.obs = (actions_phys if tm else actions).data_ptr(), obs.data_ptr()
        klast = None
        if self._solver_id == _lib.TSIT5:
            klast = tuple(torch.empty(B, dtype=dt, device=dev) for _ in range(NODE))
            for i in range(NODE):
                a.k_last[i] = klast[i].data_ptr()
        rew = trunc = term = None
        if epilogue:
            rew = torch.empty((NS - 1, B, 1) if tm else (B, NS - 1, 1), dtype=dt, device=dev)
            trunc = torch.empty((NS, B, self.n_trunc) if tm else (B, NS, self.n_trunc), dtype=torch.bool, device=dev)
            term = torch.empty((NS - 1, B, 1) if tm else (B, NS - 1, 1), dtype=torch.bool, device=dev)
            a.reward, a.truncated, a.terminated = rew.data_ptr(), trunc.data_ptr(), term.data_ptr()
        _lib.check(_lib.lib().exenv_sim_ahead(plan.handle, C.byref(a), self._stream()))
        if tm:  # hand out the reference's logical shapes: (B, N+1, O'), (B, N+1), (B, N, 1) as permuted views
            obs = obs.permute(1, 0, 2)
            traj = [t.t() for t in traj]
            if epilogue:
                rew, trunc, term = rew.permute(1, 0, 2), trunc.permute(1, 0, 2), term.permute(1, 0, 2)

        # broadcast leaves stay un-materialised (stride-0 views): reference, PRNGKey, solver_state, active flag
        def over_time(x):
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=dev)
            if t.ndim == 0:
                return t.expand(B, NS)
            return t.unsqueeze(1).expand(B, NS, *t.shape[1:])

        ref_t = [over_time(getattr(init_ref, n)) for n in self.PHYS_FIELDS]
        states = self.State(
            physical_state=self.PhysicalState(*traj),
            PRNGKey=over_time(init_state.PRNGKey),
            additions=self.Additions(
                solver_state=None if klast is None else tuple(over_time(k) for k in klast),
                active_solver_state=over_time(self._constant("true", B)),
            ),
            reference=self.PhysicalState(*ref_t),
        )
        # last_state: the reference stacks every leaf into one array first (core_env.py:485-486), so all leaves come
        # back in the floating dtype (the bool flag as 1.0)
        last_state = self.State(
            physical_state=self.PhysicalState(*last),
            PRNGKey=torch.as_tensor(init_state.PRNGKey, device=dev).to(dt),
            additions=self.Additions(solver_state=klast, active_solver_state=self._constant("one", B)),
            reference=self.PhysicalState(*[torch.as_tensor(getattr(init_ref, n), device=dev).to(dt)
                                           for n in self.PHYS_FIELDS]),
        )
        if epilogue:
            return obs, states, last_state, rew, trunc, term
        return obs, states, last_state

    def time_major_actions(self, n_action_steps, fill=None):
        """An ``actions`` tensor of logical shape ``(batch_size, n_action_steps, action_dim)`` whose memory is
        time-major ``(T, B, A)`` -- what ``vmap_sim_ahead(..., layout="time_major")`` consumes without a copy."""
        phys = torch.empty((n_action_steps, self.batch_size, self.action_dim), dtype=self.dtype, device=self.device)
        if fill is not None:
            phys.copy_(torch.as_tensor(fill, dtype=self.dtype, device=self.device).permute(1, 0, 2))
        return phys.permute(1, 0, 2)

    def vmap_sim_ahead(self, init_state, actions, obs_stepsize, action_stepsize, **kw):
        """Whole trajectories for all batches in one persistent kernel (reference core_env.py:571-616).

        ``actions``: ``(batch_size, n_action_steps, action_dim)``; returns ``(observations (B, N+1, O'), states with
        ``(B, N+1)`` leaves, last_state)``.

        ``layout="batch_major"`` (default): every returned array is C-contiguous like the reference's.
        ``layout="time_major"``: same shapes and values, but the memory behind them has the batch axis innermost
        (structure of arrays: ``observations.permute(1, 0, 2)`` / ``leaf.t()`` are contiguous) -- every warp access of
        the kernel is then one contiguous segment and no on-chip transposition is needed (about 2x the throughput)."""
        assert (
            obs_stepsize <= action_stepsize
        ), "The action stepsize should be greater or equal to the observation stepsize."
        shape = tuple(actions.shape)
        assert len(shape) == 3, "The actions need to have three dimensions: (batch_size, n_action_steps, action_dim)"
        assert (
            shape[0] == self.batch_size
        ), f"The first dimension does not correspond to the batch size which is {self.batch_size}, but {shape[0]} is given"
        assert (
            shape[-1] == self.action_dim
        ), f"The last dimension does not correspond to the action dim which is {self.action_dim}, but {shape[-1]} is given"
        return self._sim_impl(self.env_properties, self.batch_size, init_state, actions, obs_stepsize,
                              action_stepsize, **kw)

    def vmap_sim_ahead_with_epilogue(self, init_state, actions, obs_stepsize, action_stepsize, **kw):
        """``vmap_sim_ahead`` with ``vmap_generate_rew_trunc_term_ahead`` (reference core_env.py:618-647) fused into
        the trajectory kernel.  Returns ``(observations, states, last_state, reward, truncated, terminated)``."""
        return self._sim_impl(self.env_properties, self.batch_size, init_state, actions, obs_stepsize,
                              action_stepsize, epilogue=True, **kw)

    # ---- reward / truncated / terminated of materialised states ------------------------------------------------------
    def _epilogue_impl(self, phys_leaves, ref_leaves, n, env_properties=None):
        """generate_reward / generate_truncated / generate_terminated over `n` states given as flat leaves."""
        plan = self._plan(env_properties, n if env_properties is not None else None)
        dev, dt = self.device, self.dtype
        a = _lib.ExenvEpilogueArgs()
        a.batch = n
        keep = []
        for i in range(self.physical_state_dim):
            x = self._as(phys_leaves[i]).reshape(-1)
            keep.append(x)
            a.state[i] = x.data_ptr()
            if self.control_state:
                r = self._as(ref_leaves[i]).reshape(-1)
                keep.append(r)
                a.ref[i] = r.data_ptr()
        rew = torch.empty(n, dtype=dt, device=dev)
        trunc = torch.empty((n, self.n_trunc), dtype=torch.bool, device=dev)
        term = torch.empty(n, dtype=torch.bool, device=dev)
        a.reward, a.truncated, a.terminated = rew.data_ptr(), trunc.data_ptr(), term.data_ptr()
        _lib.check(_lib.lib().exenv_epilogue(plan.handle, C.byref(a), self._stream()))
        return rew, trunc, term

    def vmap_generate_reward_truncated_terminated(self, state):
        """The GymWrapper epilogue for a batched state (reference gym_wrapper.py:117-126):
        ``(reward (B, 1), truncated (B, n_trunc), terminated (B, 1))``."""
        B = self.batch_size
        phys = [getattr(state.physical_state, n) for n in self.PHYS_FIELDS]
        refs = [getattr(state.reference, n) for n in self.PHYS_FIELDS]
        rew, trunc, term = self._epilogue_impl(phys, refs, B)
        return rew.reshape(B, 1), trunc, term.reshape(B, 1)

    def vmap_generate_rew_trunc_term_ahead(self, states, actions):
        """Reward, truncated and terminated flags for data generated by ``vmap_sim_ahead`` (reference
        core_env.py:618-647): truncated for every saved state ``(B, N+1, n_trunc)``, reward / terminated for
        ``states[1:]`` ``(B, N, 1)``.  (``actions`` does not enter any reward of the six environments.)"""
        shape = tuple(actions.shape)
        assert len(shape) == 3, "The actions need to have three dimensions: (batch_size, n_action_steps, action_dim)"
        assert (
            shape[0] == self.batch_size
        ), f"The first dimension does not correspond to the batch size which is {self.batch_size}, but {shape[0]} is given"
        assert (
            shape[-1] == self.action_dim
        ), f"The last dimension does not correspond to the action dim which is {self.action_dim}, but {shape[-1]} is given"
        B = self.batch_size
        first = getattr(states.physical_state, self.PHYS_FIELDS[0])
        NS = first.shape[1]
        phys = [getattr(states.physical_state, n).contiguous() for n in self.PHYS_FIELDS]
        refs = [torch.as_tensor(getattr(states.reference, n), device=self.device).expand(B, NS).contiguous()
                for n in self.PHYS_FIELDS] if self.control_state else [None] * len(phys)
        rew, trunc, term = self._epilogue_impl(phys, refs, B * NS)
        return (rew.reshape(B, NS)[:, 1:, None], trunc.reshape(B, NS, self.n_trunc), term.reshape(B, NS)[:, 1:, None])

    # ---- single-environment API (reference core_env.py:361-488): the batched kernels with B = 1 -------------------
    def _wrap1(self, state):
        """Scalar-leaf State -> (1,) leaves."""
        leaves, struct = tree_flatten(state)
        dev = self.device
        up = []
        for x in leaves:
            t = torch.as_tensor(x, device=dev)
            up.append(t.reshape(1) if t.ndim == 0 else t)
        return tree_unflatten(struct, up)

    def _unwrap1(self, state):
        leaves, struct = tree_flatten(state)
        return tree_unflatten(struct, [x[0] if isinstance(x, torch.Tensor) and x.ndim >= 1 else x for x in leaves])

    def reset(self, env_properties, rng=None, initial_state=None, vmap_helper=None):
        if initial_state is not None:
            assert tree_structure(self._template_state()) == tree_structure(
                initial_state
            ), "initial_state should have the same dataclass structure as init_state()"
            s1 = self._wrap1(initial_state)
            obs, _, _ = self._reset_impl(env_properties, 1, _lib.RESET_FROM_STATE, state=s1)
            return obs[0], initial_state
        if rng is None:
            obs, phys, refs = self._reset_impl(env_properties, 1, _lib.RESET_DEFAULT)
            prng = self._constant("nan", 1)
        else:
            obs, phys, refs = self._reset_impl(env_properties, 1, _lib.RESET_FROM_NORM, values=self._draw_norm_state(rng, 1))
            prng = _key_tensor(rng, 1, self.device)
        st = self._new_state(phys, prng, self._solver_state_dummy(1), self._constant("false", 1), refs)
        return obs[0], self._unwrap1(st)

    def init_state(self, env_properties, rng=None, vmap_helper=None):
        return self.reset(env_properties, rng)[1]

    def generate_observation(self, state, env_properties):
        return self._reset_impl(env_properties, 1, _lib.RESET_FROM_STATE, state=self._wrap1(state))[0][0]

    def generate_state_from_observation(self, obs, env_properties, key=None):
        obs = self._as(obs, (1, self.obs_dim))
        _, phys, refs = self._reset_impl(env_properties, 1, _lib.RESET_FROM_OBS, values=obs)
        prng = self._constant("nan", 1) if key is None else key
        return self._unwrap1(self._new_state(phys, prng, self._solver_state_dummy(1), self._constant("false", 1), refs))

    def step(self, state, action_norm, env_properties):
        shape = tuple(torch.as_tensor(action_norm).shape)
        assert shape == (self.action_dim,), (
            f"The action needs to be of shape (action_dim,) which is "
            + f"{(self.action_dim,)}, but {shape} is given"
        )
        obs, st = self._step_impl(env_properties, 1, self._wrap1(state), self._as(action_norm, (1, self.action_dim)))
        return obs[0], self._unwrap1(st)

    # single-environment epilogue (e.g. pendulum_env.py:297-309, 381-390; read by gym_wrapper.py:53-58): the batched
    # kernel with one state
    def _epilogue1(self, state, env_properties):
        st = self._wrap1(state)
        phys = [getattr(st.physical_state, n) for n in self.PHYS_FIELDS]
        refs = [getattr(st.reference, n) for n in self.PHYS_FIELDS]
        return self._epilogue_impl(phys, refs, 1, env_properties)

    def generate_reward(self, state, action, env_properties):
        """Reward of one environment's state, shape ``(1,)`` (reference ``jnp.array([reward])``)."""
        return self._epilogue1(state, env_properties)[0]

    def generate_truncated(self, state, env_properties):
        """``abs(observation) > 1`` of one environment, shape ``(n_trunc,)`` (fluid tank: ``[False]``; PMSM: current
        magnitude beyond 1, shape ``(1,)``)."""
        return self._epilogue1(state, env_properties)[1][0]

    def generate_terminated(self, state, reward, env_properties):
        """``reward == 0``, shape ``(1,)`` (fluid tank: ``[False]``; PMSM: the truncated flag).  ``reward`` is the value
        ``generate_reward`` returned for this state (recomputed here: the three hooks share one kernel)."""
        return self._epilogue1(state, env_properties)[2]

    def sim_ahead(self, init_state, actions, env_properties, obs_stepsize, action_stepsize):
        actions = self._as(actions)
        assert actions.ndim == 2, "The actions need to have two dimensions: (n_action_steps, action_dim)"
        assert (
            actions.shape[-1] == self.action_dim
        ), f"The last dimension does not correspond to the action dim which is {self.action_dim}, but {actions.shape[-1]} is given"
        obs, states, last = self._sim_impl(env_properties, 1, self._wrap1(init_state), actions[None], obs_stepsize,
                                           action_stepsize)
        return obs[0], self._unwrap1(states), self._unwrap1(last)

    # ---- normalisation helpers (reference core_env.py:292-359), elementwise on whatever array type is passed ---------
    def normalize_state(self, state, env_properties):
        pn = env_properties.physical_normalizations
        phys = {n: getattr(pn, n).normalize(getattr(state.physical_state, n)) for n in self.PHYS_FIELDS}
        ref = {n: getattr(pn, n).normalize(getattr(state.reference, n)) for n in self.PHYS_FIELDS}
        return self.State(self.PhysicalState(**phys), state.PRNGKey, state.additions, self.PhysicalState(**ref))

    def denormalize_state(self, norm_state, env_properties):
        pn = env_properties.physical_normalizations
        phys = {n: getattr(pn, n).denormalize(getattr(norm_state.physical_state, n)) for n in self.PHYS_FIELDS}
        ref = {n: getattr(pn, n).denormalize(getattr(norm_state.reference, n)) for n in self.PHYS_FIELDS}
        return self.State(self.PhysicalState(**phys), norm_state.PRNGKey, norm_state.additions,
                          self.PhysicalState(**ref))

    def denormalize_action(self, action_norm, env_properties):
        an = env_properties.action_normalizations
        cols = [getattr(an, n).denormalize(action_norm[..., i]) for i, n in enumerate(self.ACTION_FIELDS)]
        return torch.stack([torch.as_tensor(c) for c in cols], dim=-1)

    # ---- descriptions ---------------------------------------------------------------------------------------------
    @property
    def obs_description(self):
        import numpy as np

        return np.hstack([np.array(self._obs_names()), np.array([n + "_ref" for n in self.control_state])])

    @property
    def action_description(self):
        import numpy as np

        return np.array(list(self.ACTION_FIELDS))

    def _obs_names(self):
        return list(self.PHYS_FIELDS)


class _Plan:
    def __init__(self, handle, keep):
        self.handle, self.keep = handle, keep

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().exenv_plan_destroy(self.handle)
        except Exception:
            pass


def _as_generator(rng, device):
    if isinstance(rng, torch.Generator):
        return rng
    gen = torch.Generator(device=device)
    if isinstance(rng, torch.Tensor):
        seed = int(rng.reshape(-1)[-1].item()) if rng.numel() else 0
    else:
        seed = int(rng)
    gen.manual_seed(seed)
    return gen


def _key_tensor(rng, B, device):
    """The state's ``PRNGKey`` leaf after a random reset: a (B, 2) uint32-like key array."""
    if isinstance(rng, torch.Tensor) and rng.ndim == 2 and rng.shape[0] == B:
        return rng.to(device)
    gen = _as_generator(rng, device)
    return torch.randint(0, 2 ** 31 - 1, (B, 2), generator=gen, device=device, dtype=torch.int64)
