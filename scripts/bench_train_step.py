"""SURVEY.md 8f-1/8f-2 at cfg3: one optimisation step = rasterize fwd + bwd to the raw parameters + Adam.
reference-style: torch activations + cat + classic entry through autograd + torch.optim.Adam (seven groups, eps 1e-15);
fused: raw-parameter entry writing straight into the flat gradient buffer + gsr_adam_step on the flat parameter buffer."""
import importlib
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
optim = importlib.import_module(H.PKG_NAME + ".optim")
mv = importlib.import_module(H.PKG_NAME + ".multiview")
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
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True))
rs = H.settings(cam, torch.zeros(3))
rast = Pk.GaussianRasterizer(rs)
lrs = {"xyz": 1.6e-6, "f_dc": 2.5e-5, "f_rest": 2.5e-5 / 20.0, "opacity": 5e-4, "segment": 1e-4, "scaling": 5e-5, "rotation": 1e-5}  # small: the scene stays put


def timeit(fn, n=40, w=12):
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


# ---- reference-style step
tp = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
topt = torch.optim.Adam([{"params": [tp[optim.GROUPS[n]]], "lr": lr, "name": n} for n, lr in lrs.items()], lr=0.0, eps=1e-15)
raw_t = {"xyz": tp["means3D"], "features_dc": tp["features_dc"], "features_rest": tp["features_rest"], "segment": tp["segments"],
         "opacity": tp["opacities"], "scaling": tp["scales"], "rotation": tp["rotations"]}


def step_classic():
    topt.zero_grad(set_to_none=True)
    act = T._activate(raw_t)
    m2 = torch.zeros_like(raw_t["xyz"], requires_grad=True)
    color, radii, depth, alpha, segment = rast(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"],
                                               segments=act["segments"], scales=act["scales"], rotations=act["rotations"])
    torch.autograd.backward([color, depth], [ug["color"], ug["depth"]])
    topt.step()


t_classic = timeit(step_classic)
t_adam_torch = timeit(lambda: topt.step())
del tp, topt, raw_t
torch.cuda.empty_cache()

# ---- fused step
params = optim.FlatParameters.from_tensors(init)
grads = mv.FlatGradients(P, "cuda", split_sh=True)
opt = optim.FusedAdam(params, grads, lrs)
v = params.views
e = torch.empty(0)
m2g = torch.empty(P, 3, device="cuda")


def step_fused():
    with torch.no_grad():
        fwd = D._forward_native(v["means3D"], v["features_dc"], e, v["segments"], v["opacities"], v["scales"], v["rotations"], e, rs,
                                sh_rest=v["features_rest"], raw_params=True)
        R, color, depth, segment, alpha, radii, geom, binb, img = fwd
        D._backward_native(rs, v["means3D"], radii, e, v["segments"], v["scales"], v["rotations"], e, ug["color"], None, ug["depth"], None,
                           v["features_dc"], geom, R, binb, img, alpha, out=grads.backward_out(m2g), sh_rest=v["features_rest"],
                           raw_params=True, opacities=v["opacities"])
        opt.step()


t_fused = timeit(step_fused)
t_adam = timeit(lambda: opt.step())
nbytes = 28 * params.buffer.numel()
print(json.dumps({"workload": "cfg3: fwd + bwd to raw parameters + Adam (61 floats x 6M)", "reference_style_step_ms": round(t_classic, 4),
                  "fused_step_ms": round(t_fused, 4), "speedup": round(t_classic / t_fused, 3), "torch_adam_ms": round(t_adam_torch, 4),
                  "gsr_adam_step_ms": round(t_adam, 4), "adam_alg_bytes": nbytes, "adam_GBps": round(nbytes / (t_adam * 1e-3) / 1e9, 1)}))
