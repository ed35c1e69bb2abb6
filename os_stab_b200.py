"""Import shim: the package directory is named ``os-stab_b200`` (not a valid
Python identifier), so ``import os_stab_b200`` loads it from there."""
import importlib.util
import pathlib
import sys

_pkg = pathlib.Path(__file__).resolve().parent / "os-stab_b200"
_spec = importlib.util.spec_from_file_location(
    "os_stab_b200", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["os_stab_b200"] = _mod
_spec.loader.exec_module(_mod)
