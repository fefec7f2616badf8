"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/exenv_b200.h declares, and the
ctypes struct layouts in exciting-environments_b200/_lib.py match the header byte for byte.  No compute calls."""
import ctypes as C
import re
import subprocess
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "exenv_b200.h"


def test_header_symbols_are_exported(excenvs):
    L = excenvs._lib
    text = HEADER.read_text()
    declared = set(re.findall(r"^(?:int|size_t|uint64_t)\s+(exenv_\w+)\s*\(", text, flags=re.M))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = L.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.exenv_abi_version() == excenvs._lib.ABI_VERSION == 3
    exported = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    for sym in declared:
        assert f" T {sym}" in exported


def test_struct_layouts_match_header(excenvs):
    L = excenvs._lib
    names = ["ExenvDesc", "ExenvDims", "ExenvStepArgs", "ExenvSimArgs", "ExenvResetArgs", "ExenvScheduleInfo",
             "ExenvEpilogueArgs", "ExenvGymArgs"]
    prog = '#include <stdio.h>\n#include "exenv_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    prog += 'printf("off_p %zu\\n", offsetof(ExenvDesc, p));printf("off_pp %zu\\n", offsetof(ExenvDesc, pp));'
    prog += 'printf("off_lut %zu\\n", offsetof(ExenvDesc, lut_gx));'
    prog += 'printf("off_actions %zu\\n", offsetof(ExenvSimArgs, actions));'
    prog += 'printf("off_tile %zu\\n", offsetof(ExenvSimArgs, tile_steps));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        src, exe = Path(td) / "t.c", Path(td) / "t"
        src.write_text(prog)
        subprocess.run(["gcc", "-I", str(HEADER.parent), "-o", str(exe), str(src)], check=True)
        out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(L, n)), n
    assert int(out["off_p"]) == L.ExenvDesc.p.offset and int(out["off_pp"]) == L.ExenvDesc.pp.offset
    assert int(out["off_lut"]) == L.ExenvDesc.lut_gx.offset
    assert int(out["off_actions"]) == L.ExenvSimArgs.actions.offset
    assert int(out["off_tile"]) == L.ExenvSimArgs.tile_steps.offset


def test_no_cpu_fallback_and_error_reporting(excenvs):
    """Without a CUDA device plan creation fails loudly (EXENV_ERR_CUDA); bad descriptors are rejected first."""
    import torch

    L = excenvs._lib
    lib = L.lib()
    dims = L.ExenvDims()
    assert lib.exenv_env_dims(5, C.byref(dims)) == 0
    assert (dims.n_phys, dims.n_ode, dims.n_act, dims.n_obs, dims.n_static) == (7, 3, 2, 8, 7)
    assert lib.exenv_env_dims(9, C.byref(dims)) == -1
    d = L.ExenvDesc()
    d.env, d.solver, d.dtype, d.tau = 0, 7, 0, 1e-4
    h = C.c_void_p()
    assert lib.exenv_plan_create(C.byref(d), 0, C.byref(h)) == -1 and "solver" in L.last_error()
    if not torch.cuda.is_available():
        d.solver = 0
        rc = lib.exenv_plan_create(C.byref(d), 0, C.byref(h))
        assert rc == -5 and "no CPU fallback" in L.last_error()
        env = excenvs.Pendulum(batch_size=4)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            env.vmap_reset()
    assert lib.exenv_step(None, None, None) == -2


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, include, link or load it."""
    pkg = ROOT / "exciting-environments_b200"
    pat = re.compile(r"(from|import)\s+oracle\b|exenv_oracle|libexenv_oracle|oracle/_|[\"']oracle[\"']")
    files = [f for ext in ("*.py", "*.cu", "*.cuh", "*.inl", "*.h", "Makefile") for f in pkg.rglob(ext)]
    assert len(files) > 10
    for f in files:
        assert not pat.search(f.read_text()), f


def test_xla_ffi_shim_compiles_against_the_api_stub_and_covers_every_entry_point(excenvs):
    """csrc/xla_ffi_shim.cc cannot be built against the real XLA headers in this image (no jax); it is compiled against
    tests/ffi_stub (a stub of the part of the public xla::ffi API it uses) so that handler names, arity and the C-ABI
    structs they fill stay consistent with include/exenv_b200.h, and every launching entry point has a handler that
    jax_ffi.py registers."""
    import re

    root = HEADER.parent.parent
    shim = root / "exciting-environments_b200" / "csrc" / "xla_ffi_shim.cc"
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", str(root / "tests" / "ffi_stub"),
                        "-I", "/usr/local/cuda/include", str(shim)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    src = shim.read_text()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    from exciting_environments_b200 import jax_ffi

    assert handlers == set(jax_ffi.TARGETS.values())
    for entry in ("exenv_step(", "exenv_sim_ahead(", "exenv_reset_pack(", "exenv_epilogue(", "exenv_gym_epilogue("):
        assert entry in src, entry
