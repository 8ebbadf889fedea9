"""TEST INFRASTRUCTURE — loader for the unmodified reference (`/root/reference`).

Only usable in the build container (the GPU box has no `/root/reference`).  It is
used by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/` and by the CPU tests that validate the restatement in
`oracle/robust_unet_ref.py` against the real reference when it is present.

The reference scripts import matplotlib / GDAL / tkinter at module scope
(`Main_Final.py:19`); none of those is used by the hot path, so they are stubbed.
Nothing here is imported by the product package.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("RBU_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Main_Final.py"))


def load_reference():
    """Import `Main_Final` from the read-only reference tree and return the module."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_DIR}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "matplotlib.gridspec", "osgeo", "osgeo.gdal"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    sys.modules["matplotlib"].rcParams = {}
    sys.modules["matplotlib.colors"].ListedColormap = object
    sys.modules["osgeo"].gdal = sys.modules["osgeo.gdal"]
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import Main_Final  # noqa: E402  (prints "Using device: cpu")
    return Main_Final
