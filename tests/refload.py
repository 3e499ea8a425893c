"""Load the UNMODIFIED reference package under the alias ``velocity_asr_ref``.

The product package in this repo is also called ``velocity_asr`` (it is a drop-in), so
the reference is imported under another name.  Looked for in ``/root/reference`` (build
container only) and ``baseline/_ref`` (pip --target install, git-ignored, travels with
gpurun).  Returns None when neither exists; tests that need it then skip.
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ALIAS = "velocity_asr_ref"


def load_reference():
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    for root in ("/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        init = os.path.join(root, "velocity_asr", "__init__.py")
        if os.path.exists(init):
            spec = importlib.util.spec_from_file_location(
                _ALIAS, init, submodule_search_locations=[os.path.dirname(init)])
            mod = importlib.util.module_from_spec(spec)
            sys.modules[_ALIAS] = mod
            spec.loader.exec_module(mod)
            return mod
    return None
