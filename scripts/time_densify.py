"""densify_and_prune at 6M Gaussians: the reference's per-tensor masks/cats (restated, tests/test_densify_gpu.py) vs one row selection."""
import importlib, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import test_densify_gpu as T
H.pkg()
optim = importlib.import_module(H.PKG_NAME + ".optim"); mv = importlib.import_module(H.PKG_NAME + ".multiview")
P = 6_000_000
g = torch.Generator().manual_seed(0)
init = {"means3D": torch.randn(P, 3, generator=g) * 3, "features_dc": torch.randn(P, 1, 3, generator=g), "features_rest": torch.randn(P, 15, 3, generator=g) * 0.1,
        "segments": torch.randn(P, 2, generator=g), "opacities": torch.randn(P, 1, generator=g) * 3, "scales": torch.randn(P, 3, generator=g) * 1.2 - 3.0,
        "rotations": torch.randn(P, 4, generator=g)}
init = {k: v.cuda() for k, v in init.items()}
lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "segment": 0.01, "scaling": 5e-3, "rotation": 1e-3}
acc = torch.rand(P, 1, generator=g).cuda() * 0.002; den = torch.randint(1, 4, (P, 1), generator=g).float().cuda()
def run_ref():
    ref = T.RefModel(init, lrs, 0.01)
    for n, k in T.NAMES.items(): ref.p[n].grad = torch.zeros_like(ref.p[n])
    ref.optimizer.step()
    ref.xyz_gradient_accum, ref.denom = acc.clone(), den.clone()
    torch.cuda.synchronize(); t = time.time()
    ref.densify_and_prune(0.0002, 0.005, 5.0, 20, optim.build_rotation)
    torch.cuda.synchronize(); return (time.time() - t) * 1e3, ref.p["xyz"].shape[0]
def run_ours():
    params = optim.FlatParameters.from_tensors(init); grads = mv.FlatGradients(P, "cuda", split_sh=True)
    opt = optim.FusedAdam(params, grads, lrs); opt.step()
    torch.cuda.synchronize(); t = time.time()
    np_, ng, idx = optim.densify_and_prune(params, opt, acc.clone(), den.clone(), 0.0002, 0.005, 5.0, 20)
    torch.cuda.synchronize(); return (time.time() - t) * 1e3, idx.numel()
run_ref(); run_ours()
r = [run_ref() for _ in range(3)]; o = [run_ours() for _ in range(3)]
print("densify_and_prune at 6M -> %d rows: reference algorithm (torch masks/cats) %.1f ms, one row selection per buffer %.1f ms" % (o[0][1], min(x[0] for x in r), min(x[0] for x in o)))
