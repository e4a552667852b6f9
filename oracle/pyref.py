"""Import hook for the byte-compiled reference callers under oracle/_ref/pyref (TEST INFRASTRUCTURE ONLY).

oracle/build_ref.py byte-compiles the reference's Python callers of the operator where they lie (/root/reference) into
sourceless byte-code files `*.refpyc` (the Python analogue of oracle/_ref/*.so).  This finder makes them importable under the
module names the reference itself uses (`gaussian_renderer`, `scene.gaussian_model`, `utils.*` -- train_seq.py puts
./gaussian_splatting on sys.path -- and `gaussian_splatting.*`, `gaustar_scene.*`, `gaustar_utils.*`); directories without an
`__init__` become namespace packages, so no package initialiser of the reference runs that the callers do not import themselves.
Never imported by the product package.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(HERE, "_ref", "pyref")
SUFFIX = ".refpyc"
ROOTS = [os.path.join(PYREF, "gaustar"), os.path.join(PYREF, "gaussian_splatting"), PYREF]
TOP_LEVEL = {"gaussian_renderer", "scene", "utils", "gaussian_splatting", "gaustar_scene", "gaustar_utils"}


def available() -> bool:
    return os.path.exists(os.path.join(PYREF, "gaussian_splatting", "gaussian_renderer", "__init__" + SUFFIX))


class RefPycFinder:
    @staticmethod
    def find_spec(fullname, path=None, target=None):
        if fullname.split(".")[0] not in TOP_LEVEL:
            return None
        rel = fullname.replace(".", os.sep)
        for root in ROOTS:
            base = os.path.join(root, rel)
            if os.path.isfile(base + SUFFIX):
                return importlib.util.spec_from_file_location(fullname, base + SUFFIX, loader=importlib.machinery.SourcelessFileLoader(fullname, base + SUFFIX))
            init = os.path.join(base, "__init__" + SUFFIX)
            if os.path.isfile(init):
                return importlib.util.spec_from_file_location(fullname, init, loader=importlib.machinery.SourcelessFileLoader(fullname, init),
                                                              submodule_search_locations=[base])
        dirs = [os.path.join(root, rel) for root in ROOTS if os.path.isdir(os.path.join(root, rel))]
        if dirs:  # a namespace package: the reference's own __init__ (if any) is deliberately not part of the build
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = dirs
            return spec
        return None


def install():
    if RefPycFinder not in sys.meta_path:
        sys.meta_path.insert(0, RefPycFinder)


def uninstall():
    if RefPycFinder in sys.meta_path:
        sys.meta_path.remove(RefPycFinder)


def load_file(name: str, relpath: str, submodule_search_locations=None):
    """Load one byte-compiled file under an explicit module name (the stock operator wrapper: `ref_operator/...`)."""
    path = os.path.join(PYREF, relpath + SUFFIX)
    spec = importlib.util.spec_from_file_location(name, path, loader=importlib.machinery.SourcelessFileLoader(name, path),
                                                  submodule_search_locations=submodule_search_locations)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    return spec, mod
