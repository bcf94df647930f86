"""Importable alias of the `pika-zoo_b200/` package (a hyphen is not a valid identifier).

    import pikazoo_b200
    from pikazoo_b200 import pikazoo_v0, PikaVecEnv
    from pikazoo_b200.wrappers import SimplifyAction
"""

import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_real = _importlib.import_module("pika-zoo_b200")
_sys.modules[__name__] = _real
for _name in ("pikazoo_v0", "wrappers", "vec_env", "dist", "spaces", "_lib"):
    _sys.modules[f"{__name__}.{_name}"] = _importlib.import_module(f"pika-zoo_b200.{_name}")
