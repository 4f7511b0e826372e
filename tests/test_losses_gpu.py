"""gsr_image_loss against the reference's loss written with torch ops exactly as utils/loss_utils.py:104-150
and train.py:110-111 do (grouped F.conv2d with the 11x11 sigma-1.5 window, zero padding), loss value and autograd gradient."""
import importlib
from math import exp

import pytest
import torch
import torch.nn.functional as F

import helpers as H

pytestmark = pytest.mark.gpu


def _window(channel):
    g = torch.Tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous()


def _ref_loss(img1, img2, lam):
    channel = img1.size(-3)
    window = _window(channel).to(img1.device).type_as(img1)
    mu1 = F.conv2d(img1, window, padding=5, groups=channel)
    mu2 = F.conv2d(img2, window, padding=5, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=5, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=5, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=5, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    l1 = torch.abs(img1 - img2).mean()
    ssim = ssim_map.mean()
    return (1.0 - lam) * l1 + lam * (1.0 - ssim), l1, ssim


@pytest.mark.parametrize("C,Hh,W,lam", [(3, 131, 213, 0.2), (3, 64, 64, 0.2), (1, 7, 9, 0.5), (3, 1080, 1920, 0.2)])
def test_l1_ssim_loss_and_gradient_match_torch(C, Hh, W, lam):
    H.pkg()
    losses = importlib.import_module(H.PKG_NAME + ".losses")
    g = torch.Generator().manual_seed(C * 1000 + W)
    gt = torch.rand(C, Hh, W, generator=g).cuda()
    # a "render" correlated with the ground truth, like a partly trained model
    img = (gt + 0.15 * torch.randn(C, Hh, W, generator=g).cuda()).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    want, l1, ssim = _ref_loss(a, gt, lam)
    want.backward()
    b = img.clone().requires_grad_(True)
    got = losses.l1_ssim_loss(b, gt, lam)
    got.backward()
    stats, _ = losses.l1_ssim_loss_and_grad(img, gt, lam, want_grad=False)
    assert abs(float(got.detach()) - float(want.detach())) <= 1e-5, (float(got.detach()), float(want.detach()))
    assert abs(float(stats[0]) - float(l1.detach())) <= 1e-5 and abs(float(stats[1]) - float(ssim.detach())) <= 1e-5
    assert H.rel_linf(b.grad, a.grad) <= 1e-4, H.rel_linf(b.grad, a.grad)
    # deterministic: no float atomics anywhere
    got2 = losses.l1_ssim_loss(img.clone().requires_grad_(True), gt, lam)
    assert float(got2.detach()) == float(got.detach())


def test_l1_ssim_identical_images():
    H.pkg()
    losses = importlib.import_module(H.PKG_NAME + ".losses")
    img = torch.rand(3, 40, 50).cuda()
    stats, grad = losses.l1_ssim_loss_and_grad(img, img.clone(), 0.2)
    assert float(stats[0]) == 0.0 and abs(float(stats[1]) - 1.0) <= 1e-6 and abs(float(stats[2])) <= 1e-6
    assert float(grad.abs().max()) <= 1e-6
    with pytest.raises(RuntimeError, match="no CPU path"):
        losses.l1_ssim_loss_and_grad(img.cpu(), img.cpu())


@pytest.mark.parametrize("tag", ["a", "b"])
def test_l1_ssim_matches_reference_python_fixture(tag):
    """tests/golden/pyref_train.npz: loss, L1, SSIM and dloss/dimage produced by IMPORTING the reference's utils/loss_utils.py
    (l1_loss, ssim) and combining them as train.py:110-111 does, on CPU (tests/golden/make_golden_pyref_train.py)."""
    import os

    import numpy as np

    H.pkg()
    losses = importlib.import_module(H.PKG_NAME + ".losses")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyref_train.npz"))
    img, gt = torch.from_numpy(z["loss_%s_img" % tag]).cuda(), torch.from_numpy(z["loss_%s_gt" % tag]).cuda()
    stats, grad = losses.l1_ssim_loss_and_grad(img, gt, 0.2)
    want = z["loss_%s_stats" % tag]
    for i in range(3):
        assert abs(float(stats[i]) - want[i]) <= 1e-5, (i, float(stats[i]), want[i])
    assert H.rel_linf(grad, torch.from_numpy(z["loss_%s_grad" % tag]).cuda()) <= 1e-4
