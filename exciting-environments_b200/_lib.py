"""ctypes binding of ``libexenv_b200.so`` (the C ABI declared in ``include/exenv_b200.h``).

The library is built in-tree (``csrc/Makefile`` -> ``lib/libexenv_b200.so``) for sm_100a only.  There is no CPU
implementation behind this module: if the shared library is missing, or no CUDA device is usable, every compute
call raises.  The field layouts below must match the header byte for byte (checked by tests/test_abi.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# EXENV_B200_LIB: load another build of the same library (A/B measurements of kernel variants)
LIB_PATH = Path(os.environ.get("EXENV_B200_LIB") or (_PKG / "lib" / "libexenv_b200.so"))
HEADER_PATH = _PKG.parent / "include" / "exenv_b200.h"

ABI_VERSION = 3
MAX_PARAMS, MAX_CONTROL, MAX_PHYS, MAX_ODE = 32, 8, 8, 4
EULER, TSIT5 = 0, 1
F32, F64 = 0, 1
FLAG_SATURATED = 1
LAYOUT_BATCH_MAJOR, LAYOUT_TIME_MAJOR = 0, 1
RESET_DEFAULT, RESET_FROM_NORM, RESET_FROM_OBS, RESET_FROM_STATE = 0, 1, 2, 3

STATUS = {0: "EXENV_OK", -1: "EXENV_ERR_INVALID", -2: "EXENV_ERR_NULL", -3: "EXENV_ERR_ALIGN",
          -4: "EXENV_ERR_UNSUPPORTED", -5: "EXENV_ERR_CUDA", -6: "EXENV_ERR_MAX_STEPS", -7: "EXENV_ERR_NOMEM"}


class ExenvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class ExenvDesc(C.Structure):
    _fields_ = [
        ("env", C.c_int32), ("solver", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32),
        ("tau", C.c_double),
        ("n_control", C.c_int32), ("control", C.c_int32 * MAX_CONTROL), ("time_rule", C.c_int32),
        ("p", C.c_double * MAX_PARAMS), ("pp", C.c_void_p * MAX_PARAMS),
        ("lut_gx", C.c_void_p), ("lut_gy", C.c_void_p), ("lut_maps", C.c_void_p),
        ("lut_nx", C.c_int32), ("lut_ny", C.c_int32),
    ]


class ExenvDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_phys", "n_ode", "n_act", "n_obs", "n_static", "n_trunc")]


class ExenvStepArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int64),
        ("state_in", C.c_void_p * MAX_PHYS), ("state_out", C.c_void_p * MAX_PHYS),
        ("action", C.c_void_p), ("obs", C.c_void_p),
        ("k_out", C.c_void_p * MAX_ODE), ("ref", C.c_void_p * MAX_PHYS),
        ("reward", C.c_void_p), ("truncated", C.c_void_p), ("terminated", C.c_void_p),
    ]


class ExenvSimArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int64), ("n_actions", C.c_int64),
        ("obs_stepsize", C.c_double), ("action_stepsize", C.c_double), ("max_steps", C.c_int64),
        ("state_in", C.c_void_p * MAX_PHYS), ("actions", C.c_void_p), ("obs", C.c_void_p),
        ("traj", C.c_void_p * MAX_PHYS), ("last", C.c_void_p * MAX_PHYS), ("k_last", C.c_void_p * MAX_ODE),
        ("ref", C.c_void_p * MAX_PHYS),
        ("reward", C.c_void_p), ("truncated", C.c_void_p), ("terminated", C.c_void_p),
        ("tile_steps", C.c_int32), ("block_envs", C.c_int32), ("layout", C.c_int32), ("reserved_", C.c_int32),
    ]


class ExenvEpilogueArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("state", C.c_void_p * MAX_PHYS), ("ref", C.c_void_p * MAX_PHYS),
                ("reward", C.c_void_p), ("truncated", C.c_void_p), ("terminated", C.c_void_p)]


class ExenvGymArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("state", C.c_void_p * MAX_PHYS), ("ref", C.c_void_p * MAX_PHYS),
                ("hold_steps", C.c_void_p), ("rng", C.c_void_p),
                ("hold_steps_min", C.c_int32), ("hold_steps_max", C.c_int32), ("force_new", C.c_int32),
                ("reserved_", C.c_int32),
                ("reward", C.c_void_p), ("truncated", C.c_void_p), ("terminated", C.c_void_p)]


class ExenvScheduleInfo(C.Structure):
    _fields_ = [("n_save", C.c_int64), ("n_steps", C.c_int64), ("n_tiles", C.c_int32), ("act_cap", C.c_int32),
                ("save_cap", C.c_int32), ("tile_steps", C.c_int32)]


class ExenvResetArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int64), ("mode", C.c_int32), ("values", C.c_void_p),
        ("state_out", C.c_void_p * MAX_PHYS), ("ref_out", C.c_void_p * MAX_PHYS), ("obs", C.c_void_p),
    ]


# every symbol include/exenv_b200.h declares
SYMBOLS = [
    "exenv_abi_version", "exenv_last_error", "exenv_env_dims", "exenv_launch_count", "exenv_plan_create",
    "exenv_plan_destroy", "exenv_step", "exenv_sim_ahead_shape", "exenv_sim_ahead", "exenv_sim_ahead_host",
    "exenv_reset_pack", "exenv_schedule_build", "exenv_epilogue", "exenv_trunc_thresholds", "exenv_sim_ahead_prepare", "exenv_gym_epilogue",
]


def build(verbose: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(_PKG / "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libexenv_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded shared library.  Raises (loudly) when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                f"`make -C {_PKG / 'csrc'}`. exciting-environments_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(str(LIB_PATH))
        L.exenv_last_error.restype = C.c_size_t
        L.exenv_last_error.argtypes = [C.c_char_p, C.c_size_t]
        L.exenv_launch_count.restype = C.c_uint64
        L.exenv_plan_create.argtypes = [C.POINTER(ExenvDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.exenv_plan_destroy.argtypes = [C.c_void_p]
        L.exenv_step.argtypes = [C.c_void_p, C.POINTER(ExenvStepArgs), C.c_void_p]
        L.exenv_sim_ahead.argtypes = [C.c_void_p, C.POINTER(ExenvSimArgs), C.c_void_p]
        L.exenv_sim_ahead_host.argtypes = [C.c_void_p, C.POINTER(ExenvSimArgs), C.c_int64]
        L.exenv_sim_ahead_shape.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64)]
        L.exenv_sim_ahead_prepare.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32,
                                              C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.exenv_reset_pack.argtypes = [C.c_void_p, C.POINTER(ExenvResetArgs), C.c_void_p]
        L.exenv_env_dims.argtypes = [C.c_int32, C.POINTER(ExenvDims)]
        L.exenv_epilogue.argtypes = [C.c_void_p, C.POINTER(ExenvEpilogueArgs), C.c_void_p]
        L.exenv_gym_epilogue.argtypes = [C.c_void_p, C.POINTER(ExenvGymArgs), C.c_void_p]
        L.exenv_schedule_build.argtypes = [C.POINTER(ExenvDesc), C.c_int64, C.c_double, C.c_double, C.c_int32,
                                           C.POINTER(ExenvScheduleInfo), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.exenv_trunc_thresholds.argtypes = [C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_double),
                                             C.POINTER(C.c_double)]
        assert L.exenv_abi_version() == ABI_VERSION, "libexenv_b200.so was built from another header: rebuild it"
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().exenv_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int):
    if rc != 0:
        raise ExenvError(rc, last_error())


def schedule(desc, n_actions, obs_stepsize, action_stepsize, tile_steps=32):
    """Host-only: the emulated diffeqsolve schedule as numpy arrays (words, dt, theta, tiles (n_tiles, 6))."""
    import numpy as np

    info = ExenvScheduleInfo()
    check(lib().exenv_schedule_build(C.byref(desc), n_actions, obs_stepsize, action_stepsize, tile_steps,
                                     C.byref(info), None, None, None, None))
    words = np.empty(info.n_steps, np.uint64)
    dt = np.empty(info.n_steps, np.float64)
    theta = np.empty(info.n_save, np.float64)
    tiles = np.empty((info.n_tiles, 6), np.int32)
    check(lib().exenv_schedule_build(C.byref(desc), n_actions, obs_stepsize, action_stepsize, tile_steps,
                                     C.byref(info), words.ctypes.data, dt.ctypes.data, theta.ctypes.data,
                                     tiles.ctypes.data))
    return info, words, dt, theta, tiles


def launch_count() -> int:
    return int(lib().exenv_launch_count())
