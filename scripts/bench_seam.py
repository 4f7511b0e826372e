"""the seam between the model's raw parameters and the rasterizer at cfg3 (6 M Gaussians, 1080p), fwd + bwd to
the RAW parameters. classic = torch activations + torch.cat (scene/gaussian_model.py:100-124) + classic entry through autograd;
fused = rasterize_gaussians_raw (activations inside the preprocess kernels, no cat, gradients written once)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import test_raw_params_gpu as T  # noqa: E402

Pk = H.pkg()
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS["cfg3"]
gs, cam = syn.make_scene("cfg3")
eps = 1e-6
logit = lambda p: torch.log(p.clamp(eps, 1 - eps) / (1 - p.clamp(eps, 1 - eps)))
raw = {"xyz": gs["means3D"], "features_dc": gs["shs"][:, :1].contiguous(), "features_rest": gs["shs"][:, 1:].contiguous(),
       "segment": logit(gs["segments"]), "opacity": logit(gs["opacities"]), "scaling": torch.log(gs["scales"]), "rotation": gs["rotations"] * 1.7}
raw = {k: v.cuda().requires_grad_(True) for k, v in raw.items()}
del gs
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True))
rs = H.settings(cam, torch.zeros(3))
rast = Pk.GaussianRasterizer(rs)


def zero():
    for v in raw.values():
        v.grad = None


def classic():
    zero()
    act = T._activate(raw)
    m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
    color, radii, depth, alpha, segment = rast(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"],
                                               segments=act["segments"], scales=act["scales"], rotations=act["rotations"])
    torch.autograd.backward([color, depth], [ug["color"], ug["depth"]])


def fused():
    zero()
    m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
    color, radii, depth, alpha, segment = rast.forward_raw(raw["xyz"], m2, raw["features_dc"], raw["features_rest"], raw["segment"],
                                                           raw["opacity"], raw["scaling"], raw["rotation"])
    torch.autograd.backward([color, depth], [ug["color"], ug["depth"]])


def timeit(fn, n=50, w=15):
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


t_c = timeit(classic)
g_c = {k: v.grad.clone() for k, v in raw.items()}
t_f = timeit(fused)
err = {k: "%.1e" % H.rel_linf(raw[k].grad, g_c[k]) for k in raw}
print(json.dumps({"workload": "cfg3 fwd+bwd to raw parameters", "classic_torch_activations_ms": round(t_c, 4), "fused_raw_entry_ms": round(t_f, 4),
                  "speedup": round(t_c / t_f, 3), "grad_rel_linf_fused_vs_classic": err}))
