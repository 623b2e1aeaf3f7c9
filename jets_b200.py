"""Import shim: ``import jets_b200`` loads the package that lives in ``./jets.jl_b200/`` (a
directory name Python cannot import directly because of the dot)."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "jets.jl_b200")
_spec = _u.spec_from_file_location("jets_b200", _os.path.join(_d, "__init__.py"),
                                   submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["jets_b200"] = _mod
_spec.loader.exec_module(_mod)
