"""Imports the package directory `sos-slam_b200/` (hyphenated, so not importable by name) as module
`sos_slam_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    if "sos_slam_b200" in sys.modules:
        return sys.modules["sos_slam_b200"]
    path = os.path.join(_ROOT, "sos-slam_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("sos_slam_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sos_slam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
