"""GPU tests of the fused-activation entry (GsrGaussians.raw_params / rasterize_gaussians_raw): raw model
parameters in, activations (scene/gaussian_model.py:100-124: sigmoid / exp / normalize / cat) inside the preprocess kernels.
Oracle: the same raw leaves pushed through the torch activations and the CLASSIC entry (ours, and the reference CUDA build
when present) -- outputs must match, raw-parameter gradients must match autograd's to <= 1e-4 relative."""
import importlib

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

RAW = ["xyz", "features_dc", "features_rest", "segment", "opacity", "scaling", "rotation"]


def _raw_scene(P, W, Hh, seed):
    """Raw parameters whose activations reproduce the synthetic scene (rotations get a random length: normalize matters)."""
    syn = H.synthetic()
    gs, cam = syn.make_scene(P, W, Hh, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    eps = 1e-6
    logit = lambda p: torch.log(p.clamp(eps, 1 - eps) / (1 - p.clamp(eps, 1 - eps)))
    raw = {
        "xyz": gs["means3D"].clone(),
        "features_dc": gs["shs"][:, :1].contiguous(),
        "features_rest": gs["shs"][:, 1:].contiguous(),
        "segment": logit(gs["segments"]),
        "opacity": logit(gs["opacities"]),
        "scaling": torch.log(gs["scales"]),
        "rotation": gs["rotations"] * (0.25 + 3.0 * torch.rand(P, 1, generator=g)),
    }
    return {k: v.cuda() for k, v in raw.items()}, cam


def _activate(raw):
    """scene/gaussian_model.py:100-124"""
    return {"means3D": raw["xyz"], "shs": torch.cat((raw["features_dc"], raw["features_rest"]), dim=1),
            "segments": torch.sigmoid(raw["segment"]), "opacities": torch.sigmoid(raw["opacity"]), "scales": torch.exp(raw["scaling"]),
            "rotations": torch.nn.functional.normalize(raw["rotation"])}


def _loss(outs, ug):
    color, radii, depth, alpha, segment = outs
    return (color * ug["color"]).sum() + (depth * ug["depth"]).sum() + (alpha * ug["alpha"]).sum() + (segment * ug["segment"]).sum()


def _run_classic(rasterize, raw0, rs, ug):
    raw = {k: v.clone().requires_grad_(True) for k, v in raw0.items()}
    act = _activate(raw)
    m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
    outs = rasterize(act, m2, rs)
    _loss(outs, ug).backward()
    return outs, {k: raw[k].grad for k in RAW}, m2.grad


@pytest.mark.parametrize("P,W,Hh,seed,deg", [(40_000, 400, 304, 5, 3), (9_000, 213, 131, 6, 1)])
def test_fused_activations_match_torch_activations(P, W, Hh, seed, deg):
    Pk = H.pkg()
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    raw0, cam = _raw_scene(P, W, Hh, seed)
    ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True))
    rs = H.settings(cam, torch.tensor([0.2, 0.1, 0.4]), sh_degree=deg)

    def ours_classic(act, m2, rs):
        return Pk.GaussianRasterizer(rs)(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"],
                                         segments=act["segments"], scales=act["scales"], rotations=act["rotations"])

    outs_c, g_c, m2_c = _run_classic(ours_classic, raw0, rs, ug)

    raw = {k: v.clone().requires_grad_(True) for k, v in raw0.items()}
    m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
    outs_f = Pk.GaussianRasterizer(rs).forward_raw(raw["xyz"], m2, raw["features_dc"], raw["features_rest"], raw["segment"], raw["opacity"],
                                                   raw["scaling"], raw["rotation"])
    _loss(outs_f, ug).backward()

    # activations are spelled like ATen's kernels: identical radii, images equal to rounding
    assert torch.equal(outs_f[1], outs_c[1]), "radii differ: %d" % int((outs_f[1] != outs_c[1]).sum())
    for a, b, name in zip(outs_f, outs_c, ["color", "radii", "depth", "alpha", "segment"]):
        if name != "radii":
            assert float((a.detach() - b.detach()).abs().max()) <= 1e-5, name
    for k in RAW:
        assert raw[k].grad is not None and raw[k].grad.shape == raw0[k].shape, k
        assert H.rel_linf(raw[k].grad, g_c[k]) <= 1e-4, (k, H.rel_linf(raw[k].grad, g_c[k]))
    assert H.rel_linf(m2.grad, m2_c) <= 1e-4

    # and against the reference's own CUDA rasterizer behind the same torch activations
    C = H.ref_dgr()
    if C is not None:
        import os
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench

        def ref_classic(act, m2, rs):
            e = torch.empty(0)
            return bench.RefRasterize.apply(C, act["means3D"], m2, act["shs"], e, act["segments"], act["opacities"], act["scales"],
                                            act["rotations"], e, rs)

        outs_r, g_r, _ = _run_classic(ref_classic, raw0, rs, ug)
        assert torch.equal(outs_f[1], outs_r[1])
        assert float((outs_f[0].detach() - outs_r[0].detach()).abs().max()) <= 1e-5
        assert float((outs_f[2].detach() - outs_r[2].detach()).abs().max()) <= 1e-5
        for k in RAW:
            assert H.rel_linf(raw[k].grad, g_r[k]) <= 1e-4, (k, H.rel_linf(raw[k].grad, g_r[k]))


def test_fused_activations_packets_and_split_flat_buffer():
    """raw-parameter gradients through the multi-view exchange format: packets + gather into a flat buffer with separate
    features_dc / features_rest blocks equal the dense raw-parameter backward; accumulate mode sums two views."""
    Pk = H.pkg()
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh = 30_000, 320, 240
    raw, cam0 = _raw_scene(P, W, Hh, 12)
    cams = [cam0, syn.make_camera(W, Hh, yaw_deg=90.0)]
    ug = H.to_dev(syn.upstream_grads(W, Hh, 12, with_depth=True, with_segment=True, with_alpha=True))
    bg = torch.tensor([0.1, 0.2, 0.3])
    e = torch.empty(0)
    rp = {"sh_rest": raw["features_rest"], "opacities": raw["opacity"]}
    dense, blobs, campos = [], [], []
    acc = mv.FlatGradients(P, "cuda", split_sh=True)
    for vi, cam in enumerate(cams):
        rs = H.settings(cam, bg)
        fwd = D._forward_native(raw["xyz"], raw["features_dc"], e, raw["segment"], raw["opacity"], raw["scaling"], raw["rotation"], e, rs,
                                sh_rest=raw["features_rest"], raw_params=True)
        R, color, depth, segment, alpha, radii, geom, binb, img = fwd
        args = (rs, raw["xyz"], radii, e, raw["segment"], raw["scaling"], raw["rotation"], e, ug["color"], ug["segment"], ug["depth"],
                ug["alpha"], raw["features_dc"], geom, R, binb, img, alpha)
        dense.append(D._backward_native(*args, sh_rest=raw["features_rest"], raw_params=True, opacities=raw["opacity"]))
        D._backward_native(*args, sh_rest=raw["features_rest"], raw_params=True, opacities=raw["opacity"], out=acc.backward_out(),
                           accumulate=vi > 0)
        blob, cnt = D._backward_packets_native(rs, raw["xyz"], radii, raw["segment"], raw["scaling"], raw["rotation"], ug["color"],
                                               ug["segment"], ug["depth"], ug["alpha"], raw["features_dc"], geom, R, binb, img, alpha,
                                               capacity=D.last_num_visible(), raw_params=rp)
        blobs.append((blob, cnt, D.last_num_visible()))
        campos.append(cam["campos"].cuda())
    names = {"means3D": "means3D", "features_dc": "sh", "features_rest": "sh_rest", "segments": "segments", "opacities": "opacities",
             "scales": "scales", "rotations": "rotations"}
    flat = mv.FlatGradients(P, "cuda", split_sh=True)
    flat.buffer.fill_(9.0)
    leaves = {"means3D": raw["xyz"], "features_dc": raw["features_dc"], "features_rest": raw["features_rest"]}
    mv.exchange_packets(D, None, flat, leaves, blobs, [campos], 3, world=1)
    for leaf, nat in names.items():
        want = dense[0][nat] + dense[1][nat]
        assert H.rel_linf(flat.views[leaf], want) <= 2e-5, leaf
        assert H.rel_linf(acc.views[leaf], want) <= 2e-5, leaf
