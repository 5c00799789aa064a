"""Import shim: the package directory is named ``ltransv.2b_b200`` (after the
reference, LTRANS v.2b), which is not a valid Python identifier.  ``import
ltrans_b200`` loads that directory as the package ``ltrans_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ltransv.2b_b200")
_spec = _ilu.spec_from_file_location(
    "ltrans_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["ltrans_b200"] = _mod
_spec.loader.exec_module(_mod)
