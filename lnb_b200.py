"""Import shim: the product package lives in the directory `llama-nuts-and-bolts_b200/`
(the name the project mandates), which is not a valid Python identifier.  `import lnb_b200`
loads that directory as a regular package under the importable name `lnb_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "llama-nuts-and-bolts_b200")
_spec = importlib.util.spec_from_file_location(
    "lnb_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lnb_b200"] = _mod
_spec.loader.exec_module(_mod)
