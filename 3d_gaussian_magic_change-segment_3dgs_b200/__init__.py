"""B200-native (sm_100a) differentiable Gaussian-splatting rasterizer + simple-knn.

The directory name is not a Python identifier; import it with
`importlib.import_module("3d_gaussian_magic_change-segment_3dgs_b200")`, or -- the drop-in use -- put this
directory on `sys.path` so that the reference's own imports resolve here:

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer   # gaussian_renderer/__init__.py:14
    from simple_knn._C import distCUDA2                                                         # scene/gaussian_model.py:20

Everything runs in libgsr.so (csrc/, C ABI in include/gsr.h); importing fails loudly if it is not built.
"""
import os
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))


def install_dropins():
    """Make `diff_gaussian_rasterization` and `simple_knn` importable as top-level modules (what the reference imports)."""
    if PACKAGE_DIR not in sys.path:
        sys.path.insert(0, PACKAGE_DIR)


install_dropins()
import diff_gaussian_rasterization  # noqa: E402
import simple_knn  # noqa: E402
from diff_gaussian_rasterization import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,  # noqa: E402,F401
                                         mark_visible, export_state)
from simple_knn._C import distCUDA2  # noqa: E402,F401

_lib = sys.modules["_gsr_b200_lib"]


def build(force=False, verbose=False):
    from importlib import util as _u

    spec = _u.spec_from_file_location("_gsr_b200_build", os.path.join(PACKAGE_DIR, "build.py"))
    mod = _u.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)
