"""Import alias: ``import frx`` loads the package that lives in the directory
``p4-fr-sorry-math-but-love-you_b200`` (not a valid Python identifier)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p4-fr-sorry-math-but-love-you_b200")
_spec = importlib.util.spec_from_file_location("frx", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["frx"] = _mod
_spec.loader.exec_module(_mod)
