"""Import shim: the package directory is named ``genparticlefilters.jl_b200`` (with a dot), which the
import system cannot address directly; load it under the module name ``genpf_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "genparticlefilters.jl_b200")
_spec = importlib.util.spec_from_file_location("genpf_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["genpf_b200"] = _mod
_spec.loader.exec_module(_mod)
