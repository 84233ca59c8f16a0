"""Import shim: loads the package directory ``gutzwiller.jl_b200/`` (whose name is not a valid
Python identifier) under the module name ``gutzwiller_jl_b200`` and re-exports it.

    import gutzwiller_b200 as gw
"""
import importlib.util
import os
import sys

_NAME = "gutzwiller_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gutzwiller.jl_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:
        del sys.modules[_NAME]
        raise

_pkg = sys.modules[_NAME]
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
