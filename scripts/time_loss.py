"""L1 + SSIM loss forward + gradient at 1080p: torch (reference formulas, autograd) vs gsr_image_loss."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import test_losses_gpu as T
H.pkg()
losses = importlib.import_module(H.PKG_NAME + ".losses")
gt = torch.rand(3, 1080, 1920, device="cuda"); img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
def t_torch():
    a = img.clone().requires_grad_(True)
    T._ref_loss(a, gt, 0.2)[0].backward()
def t_ours():
    losses.l1_ssim_loss_and_grad(img, gt, 0.2)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("1080p L1+SSIM fwd+grad: torch %.3f ms, gsr_image_loss %.3f ms" % (timeit(t_torch), timeit(t_ours)))
