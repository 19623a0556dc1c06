"""Loader for the package directory `3d-pointcloudreconstruction_b200/` (its name is not a Python identifier).

    import psd_b200
    pkg = psd_b200.load()            # -> module with chamfer_3DDist, emdModule, Loss, Metrics, ...
    psd_b200.add_to_sys_path()       # reference-style: `from dist_chamfer_3D import chamfer_3DDist`
"""
import importlib.util
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-pointcloudreconstruction_b200")
PKG_NAME = "pointcloudreconstruction_b200"


def load():
    """Import the package; (re)build libpsd_b200.so first when it is missing or older than its sources (needs nvcc; a box
    without nvcc keeps the shipped library, and a missing library still fails loudly in _lib.py)."""
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    import shutil
    if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) and not os.environ.get("PSD_B200_LIB"):
        build(force=False)
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(PKG_NAME, None)
        raise
    return mod


def build(force: bool = False, verbose: bool = False) -> str:
    spec = importlib.util.spec_from_file_location("_psd_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def build_pybind(verbose: bool = False) -> str:
    """Compile the pybind modules `chamfer_3D` / `emd` over the C ABI (3d-pointcloudreconstruction_b200/pybind)."""
    spec = importlib.util.spec_from_file_location("_psd_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_pybind(verbose=verbose)


def add_to_sys_path():
    """What the reference does with metric/chamfer3D and metric/emd (loss/loss.py:3-4)."""
    if PKG_DIR not in sys.path:
        sys.path.insert(0, PKG_DIR)
