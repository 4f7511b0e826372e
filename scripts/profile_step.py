"""One small program that launches every kernel of the library a few times (default cfg3), for `ncu --set full` (profiles/README.md)
and, at cfg1, for `compute-sanitizer --tool memcheck`:  python scripts/profile_step.py [iterations] [workload]
classic forward + dense backward, the raw-parameter entry, packets backward + gather over 2 views, L1+SSIM and depth-supervision losses, fused Adam."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
WORKLOAD = sys.argv[2] if len(sys.argv) > 2 else "cfg3"  # cfg1 (100k Gaussians, 800x800) is small enough for compute-sanitizer
Pk = H.pkg()
mv = importlib.import_module(H.PKG_NAME + ".multiview")
optim = importlib.import_module(H.PKG_NAME + ".optim")
losses = importlib.import_module(H.PKG_NAME + ".losses")
D = Pk.diff_gaussian_rasterization
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS[WORKLOAD]
gs, cam = syn.make_scene(WORKLOAD)
gs = H.to_dev(gs)
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True))
rs = H.settings(cam, torch.zeros(3))
cams = [cam, syn.make_camera(W, Hh, yaw_deg=45.0)]
eps = 1e-6
logit = lambda p: torch.log(p.clamp(eps, 1 - eps) / (1 - p.clamp(eps, 1 - eps)))
params = optim.FlatParameters.from_tensors({"means3D": gs["means3D"], "features_dc": gs["shs"][:, :1].contiguous(),
                                            "features_rest": gs["shs"][:, 1:].contiguous(), "segments": logit(gs["segments"]),
                                            "opacities": logit(gs["opacities"]), "scales": torch.log(gs["scales"]), "rotations": gs["rotations"] * 1.3})
rgrads = mv.FlatGradients(P, "cuda", split_sh=True)
opt = optim.FusedAdam(params, rgrads, {"xyz": 1e-7, "f_dc": 1e-6, "f_rest": 1e-7, "opacity": 1e-6, "segment": 1e-6, "scaling": 1e-7, "rotation": 1e-7})
flat = mv.FlatGradients(P, "cuda")
gt = torch.rand(3, Hh, W, device="cuda")
gt_depth = torch.rand(1, Hh, W, device="cuda")
campos = [c["campos"].cuda() for c in cams]
with torch.no_grad():
    for it in range(ITERS):
        # classic entry: forward + dense backward (the bench.py step)
        fwd = mv.native_view_forward(D, gs, rs)
        mv.native_view_backward(D, gs, rs, fwd, ug, flat, first=True)
        # raw-parameter entry + loss + Adam (the native training step)
        fwd_r = mv.native_view_forward(D, params.views, rs)
        stats, g_color = losses.l1_ssim_loss_and_grad(fwd_r[1], gt, 0.2)
        losses.depth_loss_and_grad(fwd_r[2], gt_depth, 0.1, fused_from_raw_depth=True)  # depth supervision: radix-select median / quantile
        mv.native_view_backward(D, params.views, rs, fwd_r, {"color": g_color}, rgrads, first=True)
        opt.step()
        # multi-view exchange format: packets of two views + one gather pass
        sets = []
        for c in cams:
            r2 = H.settings(c, torch.zeros(3))
            f2 = mv.native_view_forward(D, gs, r2)
            sets.append(mv.native_view_backward_packets(D, gs, r2, f2, ug))
        mv.exchange_packets(D, None, flat, gs, sets, [campos], 3, world=1)
        torch.cuda.synchronize()
print("profile_step ok: %d iterations, launches=%d" % (ITERS, int(Pk._lib.lib().gsr_launch_count())))
