"""Import alias: the package directory is ``factored-neus_b200/`` (not a Python
identifier), so ``import factored_neus_b200`` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "factored-neus_b200")
_spec = importlib.util.spec_from_file_location(
    "factored_neus_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["factored_neus_b200"] = _mod
_spec.loader.exec_module(_mod)
