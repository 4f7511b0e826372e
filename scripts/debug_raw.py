import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import test_raw_params_gpu as T
Pk = H.pkg(); syn = H.synthetic()
P, W, Hh, seed, deg = 40_000, 400, 304, 5, 3
raw0, cam = T._raw_scene(P, W, Hh, seed)
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True))
rs = H.settings(cam, torch.tensor([0.2, 0.1, 0.4]), sh_degree=deg)
def ours_classic(act, m2, rs):
    return Pk.GaussianRasterizer(rs)(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"],
                                     segments=act["segments"], scales=act["scales"], rotations=act["rotations"])
o1, g1, m1 = T._run_classic(ours_classic, raw0, rs, ug)
o2, g2, m2_ = T._run_classic(ours_classic, raw0, rs, ug)
print("classic vs classic (atomic noise):", {k: "%.2e" % H.rel_linf(g1[k], g2[k]) for k in T.RAW})
raw = {k: v.clone().requires_grad_(True) for k, v in raw0.items()}
m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
of = Pk.GaussianRasterizer(rs).forward_raw(raw["xyz"], m2, raw["features_dc"], raw["features_rest"], raw["segment"], raw["opacity"], raw["scaling"], raw["rotation"])
T._loss(of, ug).backward()
print("radii equal:", torch.equal(of[1], o1[1]), "img diffs:", [float((a.detach() - b.detach()).abs().max()) for a, b in zip(of, o1) if a.dtype.is_floating_point])
print("fused vs classic:", {k: "%.2e" % H.rel_linf(raw[k].grad, g1[k]) for k in T.RAW}, "m2 %.2e" % H.rel_linf(m2.grad, m1))
act = T._activate(raw0)
D = Pk.diff_gaussian_rasterization
print("op bits equal:", torch.equal(torch.sigmoid(raw0["opacity"]), 1.0 / (1.0 + torch.exp(-raw0["opacity"]))))
