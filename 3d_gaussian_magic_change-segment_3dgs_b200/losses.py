"""Fused photometric loss: loss = (1 - lambda) * L1 + lambda * (1 - SSIM), train.py:110-111 with
utils/loss_utils.py:104-150 (11x11 Gaussian window, sigma 1.5, zero padding), forward and gradient in two CUDA launches
(gsr_image_loss). No CPU path."""
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
