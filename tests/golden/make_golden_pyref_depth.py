"""Generate tests/golden/pyref_depth.npz by IMPORTING the reference's own utils/loss_utils.compute_depth_loss (run in the build
container, where /root/reference exists, on CPU): inputs, loss value and autograd gradient for

  * compute_depth_loss(x, gt, lambda)                                   utils/loss_utils.py:88-102     (gsr_depth_loss mode 0)
  * compute_depth_loss(1 / (d / (d.max() + 1e-5)).clamp(1e-6), gt, lambda)   gaussian_renderer/__init__.py:375 + train.py:120 (mode 1)

    python tests/golden/make_golden_pyref_depth.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
from utils.loss_utils import compute_depth_loss  # noqa: E402  (reference code)

out = {}
g = torch.Generator().manual_seed(29)
for tag, (H, W) in {"a": (61, 97), "b": (40, 52), "c": (1, 7)}.items():
    gt = torch.rand(1, H, W, generator=g)
    # mode 0: an inverse-depth-like positive signal correlated with the prior
    x = (0.5 + 3.0 * gt + 0.4 * torch.randn(1, H, W, generator=g)).abs().requires_grad_(True)
    loss = compute_depth_loss(x, gt, 0.1)
    loss.backward()
    out["m0_%s_x" % tag], out["m0_%s_gt" % tag] = x.detach().numpy(), gt.numpy()
    out["m0_%s_loss" % tag], out["m0_%s_grad" % tag] = np.float64(float(loss)), x.grad.numpy()
    # mode 1: a raw depth image (some empty pixels with depth 0, like uncovered background)
    d = (2.0 + 6.0 * (1 - gt) + 0.5 * torch.randn(1, H, W, generator=g)).clamp_min(0.3)
    d[0, 0, : max(1, W // 9)] = 0.0
    d = d.requires_grad_(True)
    dn = d / (d.max() + 1e-5)  # gaussian_renderer/__init__.py:375
    loss = compute_depth_loss(1 / dn.clamp(1e-6), gt, 0.1)  # train.py:120
    loss.backward()
    out["m1_%s_d" % tag], out["m1_%s_gt" % tag] = d.detach().numpy(), gt.numpy()
    out["m1_%s_loss" % tag], out["m1_%s_grad" % tag] = np.float64(float(loss)), d.grad.numpy()

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyref_depth.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: getattr(v, "shape", ()) for k, v in out.items()})
