"""Import alias: `import rbunet` loads the package that lives in `eusipco-2026-robust-unet_b200/`
(the directory name required by the repo layout is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eusipco-2026-robust-unet_b200")
_spec = importlib.util.spec_from_file_location("rbunet", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rbunet"] = _mod
_spec.loader.exec_module(_mod)
