"""Loss kernels of the training step. No CPU path.

* photometric: loss = (1 - lambda) * L1 + lambda * (1 - SSIM), train.py:110-111 with utils/loss_utils.py:104-150 (11x11 Gaussian
  window, sigma 1.5, zero padding), forward and gradient in two CUDA launches (gsr_image_loss);
* depth supervision: compute_depth_loss (utils/loss_utils.py:88-102: median / mean-absolute-deviation normalisation of the rendered
  inverse depth and of the monocular prior, squared difference, the largest 20 % dropped by a quantile) and its gradient by radix
  select instead of torch's two full sorts (gsr_depth_loss)."""
import ctypes

import torch

from . import _lib


def l1_ssim_loss_and_grad(image, gt, lambda_dssim=0.2, grad_scale=1.0, want_grad=True):
    """Native call. image, gt: [C,H,W] fp32 CUDA. Returns (stats f32[3] = {L1, SSIM, loss} on the device, dloss/dimage * grad_scale
    or None)."""
    if not image.is_cuda:
        raise RuntimeError("image must be a CUDA tensor; libgsr has no CPU path")
    if image.dim() != 3 or image.shape != gt.shape or image.dtype != torch.float32 or gt.dtype != torch.float32:
        raise RuntimeError("image and gt must be float32 tensors of the same [C,H,W] shape")
    L = _lib.lib()
    dev = image.device
    C, H, W = (int(v) for v in image.shape)
    x, y = image.detach().contiguous(), gt.detach().contiguous()
    with torch.cuda.device(dev):
        n = int(L.gsr_image_loss_scratch_bytes(C, H, W))
        scratch = torch.empty(n, dtype=torch.uint8, device=dev)
        stats = torch.empty(3, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x) if want_grad else None
        rc = L.gsr_image_loss(x.data_ptr(), y.data_ptr(), C, H, W, float(lambda_dssim), float(grad_scale), stats.data_ptr(),
                              grad.data_ptr() if want_grad else None, scratch.data_ptr(), n, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "gsr_image_loss")
    return stats, grad


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        stats, grad = l1_ssim_loss_and_grad(image, gt, lambda_dssim, 1.0, want_grad=image.requires_grad)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(stats)
        return stats[2].clone(), stats

    @staticmethod
    def backward(ctx, g_loss, g_stats):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None


def l1_ssim_loss(image, gt, lambda_dssim=0.2):
    """Drop-in for `(1.0 - opt.lambda_dssim) * l1_loss(image, gt) + opt.lambda_dssim * (1.0 - ssim(image, gt))` (train.py:110-111).
    Differentiable with respect to `image` only (the ground truth carries no gradient in the reference either)."""
    return _L1SSIM.apply(image, gt, lambda_dssim)[0]


def depth_loss_and_grad(values, gt_depth, lambda_depth, grad_scale=1.0, want_grad=True, fused_from_raw_depth=False):
    """Native call. values: compute_depth_loss's dyn_depth (any shape, fp32 CUDA) -- or, with fused_from_raw_depth=True, the
    rasterizer's RAW depth image, in which case render()'s `depth / (depth.max() + 1e-5)` (gaussian_renderer/__init__.py:375) and
    train.py:120's `1 / depth.clamp(1e-6)` run inside the kernels too. Returns (loss f32[1] on the device,
    grad_scale * dloss/dvalues shaped like `values`, or None)."""
    if not values.is_cuda:
        raise RuntimeError("depth_loss needs CUDA tensors; libgsr has no CPU path")
    if values.numel() != gt_depth.numel() or values.dtype != torch.float32 or gt_depth.dtype != torch.float32:
        raise RuntimeError("dyn_depth and gt_depth must be float32 tensors with the same number of elements")
    L = _lib.lib()
    dev = values.device
    x, g = values.detach().contiguous(), gt_depth.detach().to(dev).contiguous()
    n = int(x.numel())
    with torch.cuda.device(dev):
        nb = int(L.gsr_depth_loss_scratch_bytes(n))
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x) if want_grad else None
        rc = L.gsr_depth_loss(x.data_ptr(), g.data_ptr(), n, float(lambda_depth), float(grad_scale), 1 if fused_from_raw_depth else 0,
                              loss.data_ptr(), grad.data_ptr() if want_grad else None, scratch.data_ptr(), nb,
                              torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "gsr_depth_loss")
    return loss, grad


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dyn_depth, gt_depth, lambda_depth, fused):
        loss, grad = depth_loss_and_grad(dyn_depth, gt_depth, lambda_depth, 1.0, want_grad=dyn_depth.requires_grad, fused_from_raw_depth=fused)
        ctx.save_for_backward(grad)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g_loss):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None


def depth_loss(dyn_depth, gt_depth, lambda_depth):
    """Drop-in for `compute_depth_loss(dyn_depth, gt_depth, lambda_depth)` (utils/loss_utils.py:88-102); differentiable with respect
    to dyn_depth (the prior carries no gradient in the reference either)."""
    return _DepthLoss.apply(dyn_depth, gt_depth, lambda_depth, False)


def depth_supervision_loss(raw_depth, gt_depth, lambda_depth):
    """`compute_depth_loss(1 / (raw_depth / (raw_depth.max() + 1e-5)).clamp(1e-6), gt_depth, lambda_depth)` -- the whole depth term
    of train.py:118-121 from the rasterizer's raw depth output -- in one native call; differentiable with respect to raw_depth."""
    return _DepthLoss.apply(raw_depth, gt_depth, lambda_depth, True)
