"""Import shim: the product package lives in the directory `adseismic.jl_b200/` (the name the project layout
prescribes), which is not a valid Python identifier.  `import adseis_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adseismic.jl_b200")
_spec = importlib.util.spec_from_file_location("adseis_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["adseis_b200"] = _mod
_spec.loader.exec_module(_mod)
