"""Import shim: the package directory is named after the reference repository
(`binary-fluctuating-lattice-boltzmann_b200/`), which is not a valid Python identifier; this module
loads it under the importable name ``bflbm_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "binary-fluctuating-lattice-boltzmann_b200")
_spec = importlib.util.spec_from_file_location("bflbm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bflbm_b200"] = _mod
_spec.loader.exec_module(_mod)
