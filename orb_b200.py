"""Import shim: the package directory is named ``gpu-load-balance_b200`` (hyphens are not
importable), so this module loads it under the name ``gpu_load_balance_b200`` and re-exports it.

    import orb_b200 as orb
    ctx = orb.Orb(n_local, d)
"""
import importlib.util
import sys
from pathlib import Path

_NAME = "gpu_load_balance_b200"
if _NAME not in sys.modules:
    _dir = Path(__file__).resolve().parent / "gpu-load-balance_b200"
    _spec = importlib.util.spec_from_file_location(_NAME, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
