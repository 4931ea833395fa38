"""Import alias: the product package lives in the directory `vehicle-cv-adas_b200/` (the name the
build contract asks for), which is not a valid Python identifier.  `import adas_b200` loads that
directory as a regular package under the name `adas_b200`."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vehicle-cv-adas_b200")
_spec = importlib.util.spec_from_file_location("adas_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["adas_b200"] = _mod
_spec.loader.exec_module(_mod)
