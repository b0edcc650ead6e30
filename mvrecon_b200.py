"""Import shim: the package directory is named ``multiview-reconstruction_b200`` (hyphen), which is not a Python
identifier.  ``import mvrecon_b200`` loads that package under this module's name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multiview-reconstruction_b200")
_spec = importlib.util.spec_from_file_location("mvrecon_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mvrecon_b200"] = _mod
_spec.loader.exec_module(_mod)
