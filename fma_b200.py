"""Import alias: ``import fma_b200`` loads the package that lives in the contract-named directory
``llm-d-fast-model-actuation_b200/`` (hyphens are not importable)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "llm-d-fast-model-actuation_b200")
_spec = _u.spec_from_file_location("fma_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["fma_b200"] = _mod
_spec.loader.exec_module(_mod)
