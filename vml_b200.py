"""Importable alias for the package directory ``video-moment-localization_b200/``.

The directory name (fixed by the project layout) contains hyphens and cannot be
imported with a plain ``import`` statement; this shim registers it under the
module name ``vml_b200`` so that ``import vml_b200`` / ``from vml_b200.smin import
SMIN`` work from the repo root.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-moment-localization_b200")


def _load():
    spec = importlib.util.spec_from_file_location(
        "vml_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vml_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
