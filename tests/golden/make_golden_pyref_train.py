"""Generate tests/golden/pyref_train.npz by IMPORTING the reference's own Python (run in the build container, where
/root/reference exists, on CPU):

  * utils/loss_utils.py l1_loss / ssim and the training loss of train.py:110-111, with its autograd gradient -- golden vector of
    gsr_image_loss;
  * utils/general_utils.py get_expon_lr_func (the xyz learning-rate schedule of scene/gaussian_model.py:178-190), inverse_sigmoid
    and build_rotation -- golden vectors of trainer.get_expon_lr_func, trainer.NativeTrainer.reset_opacity's formula and
    optim.build_rotation.

    python tests/golden/make_golden_pyref_train.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)

import utils.general_utils as gu  # noqa: E402  (reference code)
from utils.loss_utils import l1_loss, ssim  # noqa: E402  (reference code)

out = {}
g = torch.Generator().manual_seed(11)
for tag, (C, H, W) in {"a": (3, 37, 53), "b": (3, 64, 80)}.items():
    gt = torch.rand(C, H, W, generator=g)
    img = (gt + 0.2 * torch.randn(C, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    lam = 0.2
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)  # train.py:110-111
    loss.backward()
    out["loss_%s_img" % tag], out["loss_%s_gt" % tag] = img.detach().numpy(), gt.numpy()
    out["loss_%s_stats" % tag] = np.array([float(Ll1), float(s), float(loss)], np.float64)
    out["loss_%s_grad" % tag] = img.grad.numpy()

sched = gu.get_expon_lr_func(lr_init=0.00008 * 3.7, lr_final=0.0000016 * 3.7, lr_delay_mult=0.01, max_steps=30_000)
steps = np.array([0, 1, 2, 10, 100, 999, 1000, 7000, 15000, 29999, 30000, 40000])
out["lr_steps"], out["lr_values"] = steps, np.array([sched(int(t)) for t in steps], np.float64)

x = torch.rand(64, 1, generator=g) * 0.98 + 0.01
out["invsig_in"], out["invsig_out"] = x.numpy(), gu.inverse_sigmoid(torch.min(x, torch.ones_like(x) * 0.01)).numpy()  # reset_opacity

_zeros = torch.zeros
gu.torch.zeros = lambda *a, **k: _zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})  # build_rotation allocates on "cuda"
q = torch.randn(50, 4, generator=g) * 1.5
out["rot_q"], out["rot_R"] = q.numpy(), gu.build_rotation(q).numpy()
gu.torch.zeros = _zeros

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyref_train.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
