"""Importable alias of the package directory ``exciting-environments_b200/`` (a hyphen cannot appear in an
``import`` statement): ``import exciting_environments_b200 as excenvs``."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("exciting-environments_b200")
sys.modules[__name__] = _pkg
