"""gsr_depth_loss: the depth-supervision loss of train.py:118-121 / utils/loss_utils.py:88-102, value and gradient,
against (1) tests/golden/pyref_depth.npz, produced by IMPORTING the reference's own compute_depth_loss on CPU
(tests/golden/make_golden_pyref_depth.py), and (2) the same formulas in torch on the GPU at 1080p (torch.median / torch.quantile,
autograd), where the order statistics of the native radix select must be the very elements torch's sorts pick."""
import importlib
import os

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.ROOT, "tests", "golden", "pyref_depth.npz")


def _torch_depth_loss(dyn_depth, gt_depth, lambda_depth):
    """utils/loss_utils.py:88-102, expression by expression (the masked assignment written out-of-place)."""
    dyn_depth = dyn_depth.view(1, -1)
    gt_depth = gt_depth.view(1, -1)
    t_d = torch.median(dyn_depth, dim=-1, keepdim=True).values
    s_d = torch.mean(torch.abs(dyn_depth - t_d), dim=-1, keepdim=True)
    dyn_depth_norm = (dyn_depth - t_d) / s_d
    t_gt = torch.median(gt_depth, dim=-1, keepdim=True).values
    s_gt = torch.mean(torch.abs(gt_depth - t_gt), dim=-1, keepdim=True)
    gt_depth_norm = (gt_depth - t_gt) / s_gt
    arr = (dyn_depth_norm - gt_depth_norm) ** 2
    arr = torch.where(arr > torch.quantile(arr, 0.8, dim=1)[..., None], torch.zeros_like(arr), arr)
    return arr.mean() * lambda_depth


def _losses():
    H.pkg()
    return importlib.import_module(H.PKG_NAME + ".losses")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_depth_loss_matches_the_reference_python_golden(tag):
    L = _losses()
    z = np.load(GOLD)
    x = torch.from_numpy(z["m0_%s_x" % tag]).cuda().requires_grad_(True)
    gt = torch.from_numpy(z["m0_%s_gt" % tag]).cuda()
    loss = L.depth_loss(x, gt, 0.1)
    loss.backward()
    assert abs(float(loss.detach()) - float(z["m0_%s_loss" % tag])) <= 2e-6 * max(1.0, abs(float(z["m0_%s_loss" % tag])))
    assert H.rel_linf(x.grad, torch.from_numpy(z["m0_%s_grad" % tag])) <= 1e-4
    d = torch.from_numpy(z["m1_%s_d" % tag]).cuda().requires_grad_(True)
    gt1 = torch.from_numpy(z["m1_%s_gt" % tag]).cuda()
    loss1 = L.depth_supervision_loss(d, gt1, 0.1)
    loss1.backward()
    assert abs(float(loss1.detach()) - float(z["m1_%s_loss" % tag])) <= 2e-6 * max(1.0, abs(float(z["m1_%s_loss" % tag])))
    assert H.rel_linf(d.grad, torch.from_numpy(z["m1_%s_grad" % tag])) <= 1e-4


@pytest.mark.parametrize("Hh,W,seed", [(1080, 1920, 1), (840, 1297, 2), (33, 65, 3)])
def test_depth_loss_matches_torch_at_full_size(Hh, W, seed):
    L = _losses()
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(1, Hh, W, generator=g).cuda()
    d0 = (1.0 + 9.0 * (1 - gt.cpu()) + torch.randn(1, Hh, W, generator=g)).clamp_min(0.25).cuda()
    # the generic entry: compute_depth_loss on an arbitrary positive signal
    x0 = 1 / (d0 / (d0.max() + 1e-5)).clamp(1e-6)
    a = x0.clone().requires_grad_(True)
    want = _torch_depth_loss(a, gt, 0.1)
    want.backward()
    b = x0.clone().requires_grad_(True)
    got = L.depth_loss(b, gt, 0.1)
    got.backward()
    assert abs(float(got.detach()) - float(want.detach())) <= 2e-6 * max(1.0, abs(float(want.detach())))
    assert H.rel_linf(b.grad, a.grad) <= 1e-4
    # the fused entry: raw rasterizer depth in, gradient w.r.t. it, through max-normalisation, clamp and reciprocal
    a1 = d0.clone().requires_grad_(True)
    want1 = _torch_depth_loss(1 / (a1 / (a1.max() + 1e-5)).clamp(1e-6), gt, 0.1)
    want1.backward()
    b1 = d0.clone().requires_grad_(True)
    got1 = L.depth_supervision_loss(b1, gt, 0.1)
    got1.backward()
    assert abs(float(got1.detach()) - float(want1.detach())) <= 2e-6 * max(1.0, abs(float(want1.detach())))
    assert H.rel_linf(b1.grad, a1.grad) <= 1e-4
    # deterministic (no float atomics), and value-only mode agrees
    again, _ = L.depth_loss_and_grad(x0, gt, 0.1, want_grad=False)
    assert float(again[0]) == float(got.detach())
