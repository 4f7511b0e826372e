"""Generate tests/golden/pyref_sh_cov.npz by IMPORTING the reference's own Python (run in the build container, where
/root/reference exists): utils/sh_utils.py:eval_sh (+0.5, clamp_min 0 as gaussian_renderer/__init__.py:353-357) and
scene/gaussian_model.py:28-32 build_covariance_from_scaling_rotation (utils/general_utils.py build_scaling_rotation,
strip_symmetric). These are the reference's in-tree cross-checks of the kernel's SH colour and cov3D.

    python tests/golden/make_golden_pyref.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from utils.general_utils import build_scaling_rotation, strip_symmetric  # noqa: E402  (reference code)
from utils.sh_utils import eval_sh  # noqa: E402  (reference code)

syn = importlib.import_module("3d_gaussian_magic_change-segment_3dgs_b200.synthetic")

P, W, H, seed = 3000, 160, 112, 3
gs, cam = syn.make_scene(P, W, H, seed=seed)
out = dict(means3D=gs["means3D"].numpy(), opacities=gs["opacities"].numpy(), shs=gs["shs"].numpy(), scales=gs["scales"].numpy(),
           rotations=gs["rotations"].numpy())
for k in ["W", "H", "tanfovx", "tanfovy"]:
    out["cam_" + k] = np.asarray(cam[k])
for k in ["viewmatrix", "projmatrix", "campos"]:
    out["cam_" + k] = cam[k].numpy()

# build_scaling_rotation allocates on "cuda"; run it on CPU by patching the device argument
import utils.general_utils as gu  # noqa: E402

_zeros = torch.zeros
gu.torch.zeros = lambda *a, **k: _zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
L = build_scaling_rotation(1.0 * gs["scales"], gs["rotations"])
out["cov3D"] = strip_symmetric(L @ L.transpose(1, 2)).numpy()
gu.torch.zeros = _zeros

shs_view = gs["shs"].transpose(1, 2).view(-1, 3, 16)
dir_pp = gs["means3D"] - cam["campos"].repeat(P, 1)
dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
for deg in range(4):
    sh2rgb = eval_sh(deg, shs_view, dir_pp_normalized)
    out["rgb_deg%d" % deg] = torch.clamp_min(sh2rgb + 0.5, 0.0).numpy()

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pyref_sh_cov.npz"), **out)
print("wrote pyref_sh_cov.npz")
