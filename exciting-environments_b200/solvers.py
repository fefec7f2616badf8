"""Stand-ins for the two diffrax solver objects the reference accepts as ``solver=`` (core_env.py:41,
README.md:52-68).  ``diffrax`` is not installable in this image; real ``diffrax.Euler()`` / ``diffrax.Tsit5()``
instances are accepted as well (matched by class name)."""
from __future__ import annotations

from . import _lib


class Euler:
    """Explicit Euler, ``y1 = y0 + (t1 - t0) * f(t0, y0)`` (diffrax.Euler)."""

    def __repr__(self):
        return "Euler()"


class Tsit5:
    """Tsitouras' 5(4) explicit Runge-Kutta method with FSAL and 4th-order dense output (diffrax.Tsit5)."""

    def __repr__(self):
        return "Tsit5()"


def solver_id(solver) -> int:
    name = type(solver).__name__
    if name == "Euler":
        return _lib.EULER
    if name == "Tsit5":
        return _lib.TSIT5
    raise ValueError(f"unsupported solver {solver!r}: the B200 path implements Euler and Tsit5")
