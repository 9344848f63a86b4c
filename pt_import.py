"""Import helper: the package directory is named `pawsometracker.jl_b200`
(a dot is not legal in a Python module name), so it is loaded by path and
registered as `pawsometracker_jl_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "pawsometracker.jl_b200")
MOD_NAME = "pawsometracker_jl_b200"


def load():
    if MOD_NAME in sys.modules:
        return sys.modules[MOD_NAME]
    spec = importlib.util.spec_from_file_location(
        MOD_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[MOD_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(MOD_NAME, None)
        raise
    return mod
