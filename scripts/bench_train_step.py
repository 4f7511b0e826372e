"""SURVEY.md 8f-1/2/3 at cfg3 (6 M Gaussians, 1080p): one whole optimisation step
    render (fwd) -> loss = 0.8 L1 + 0.2 (1 - SSIM) (train.py:110-111) -> backward to the RAW parameters -> Adam (7 groups, eps 1e-15)
three ways:
  reference       torch activations + cat, the REFERENCE's CUDA rasterizer (oracle/_ref), torch loss (utils/loss_utils.py), torch Adam
  classic         the same torch seam / loss / optimiser around THIS library's classic drop-in entry
  native          raw-parameter entry (fused activations) + gsr_image_loss + backward into the flat gradient buffer + gsr_adam_step
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import test_losses_gpu as TL  # noqa: E402
import test_raw_params_gpu as T  # noqa: E402

Pk = H.pkg()
optim = importlib.import_module(H.PKG_NAME + ".optim")
mv = importlib.import_module(H.PKG_NAME + ".multiview")
losses = importlib.import_module(H.PKG_NAME + ".losses")
D = Pk.diff_gaussian_rasterization
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS["cfg3"]
gs, cam = syn.make_scene("cfg3")
eps = 1e-6
logit = lambda p: torch.log(p.clamp(eps, 1 - eps) / (1 - p.clamp(eps, 1 - eps)))
init = {"means3D": gs["means3D"], "features_dc": gs["shs"][:, :1].contiguous(), "features_rest": gs["shs"][:, 1:].contiguous(),
        "segments": logit(gs["segments"]), "opacities": logit(gs["opacities"]), "scales": torch.log(gs["scales"]), "rotations": gs["rotations"] * 1.7}
init = {k: v.cuda() for k, v in init.items()}
del gs
gt = torch.rand(3, Hh, W, generator=torch.Generator().manual_seed(7)).cuda()
rs = H.settings(cam, torch.zeros(3))
rast = Pk.GaussianRasterizer(rs)
lrs = {"xyz": 1.6e-6, "f_dc": 2.5e-5, "f_rest": 2.5e-5 / 20.0, "opacity": 5e-4, "segment": 1e-4, "scaling": 5e-5, "rotation": 1e-5}  # small: the scene stays put
LAM = 0.2


def timeit(fn, n=30, w=10):
    import gc
    for _ in range(w):
        fn()
    gc.collect()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_style(rasterize):
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    topt = torch.optim.Adam([{"params": [tp[optim.GROUPS[n]]], "lr": lr, "name": n} for n, lr in lrs.items()], lr=0.0, eps=1e-15)
    raw_t = {"xyz": tp["means3D"], "features_dc": tp["features_dc"], "features_rest": tp["features_rest"], "segment": tp["segments"],
             "opacity": tp["opacities"], "scaling": tp["scales"], "rotation": tp["rotations"]}

    def step():
        topt.zero_grad(set_to_none=True)
        act = T._activate(raw_t)
        m2 = torch.zeros_like(raw_t["xyz"], requires_grad=True)
        color = rasterize(act, m2)[0]
        loss = TL._ref_loss(color, gt, LAM)[0]
        loss.backward()
        topt.step()
    t = timeit(step)
    del tp, topt, raw_t
    torch.cuda.empty_cache()
    return t


out = {"workload": "cfg3: render + 0.8 L1 + 0.2 (1-SSIM) + backward to raw parameters + Adam (61 floats x 6M)"}
C = H.ref_dgr()
if C is not None:
    import bench
    e_ = torch.empty(0)
    out["reference_ms"] = round(torch_style(lambda act, m2: bench.RefRasterize.apply(C, act["means3D"], m2, act["shs"], e_, act["segments"],
                                                                                      act["opacities"], act["scales"], act["rotations"], e_, rs)), 4)
out["classic_ms"] = round(torch_style(lambda act, m2: rast(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"],
                                                           segments=act["segments"], scales=act["scales"], rotations=act["rotations"])), 4)

params = optim.FlatParameters.from_tensors(init)
grads = mv.FlatGradients(P, "cuda", split_sh=True)
opt = optim.FusedAdam(params, grads, lrs)
m2g = torch.empty(P, 3, device="cuda")


def step_native():
    with torch.no_grad():
        fwd = mv.native_view_forward(D, params.views, rs)
        stats, g_color = losses.l1_ssim_loss_and_grad(fwd[1], gt, LAM)
        mv.native_view_backward(D, params.views, rs, fwd, {"color": g_color}, grads, first=True, means2D_grad=m2g)
        opt.step()


out["native_ms"] = round(timeit(step_native), 4)
out["adam_ms"] = round(timeit(lambda: opt.step()), 4)
out["loss_ms"] = round(timeit(lambda: losses.l1_ssim_loss_and_grad(gt * 0.9, gt, LAM)), 4)
if "reference_ms" in out:
    out["native_vs_reference"] = round(out["reference_ms"] / out["native_ms"], 2)
out["native_vs_classic"] = round(out["classic_ms"] / out["native_ms"], 2)
print(json.dumps(out))
