"""Import shim: the package directory is named ``universaldiffeq.jl_b200`` (a dot cannot appear in a Python
module name), so ``import ude_b200`` loads it under the module name ``universaldiffeq_jl_b200`` and re-exports it."""
import importlib.util
import os
import sys

_NAME = "universaldiffeq_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "universaldiffeq.jl_b200")


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load()
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
PACKAGE_DIR = _DIR
