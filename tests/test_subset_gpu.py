"""Index-list rendering (GsrGaussians.subset): rendering the Gaussians listed in `subset` must equal rendering
materialised masked copies of every tensor -- what the reference's viewer does for its bbox mask
(gaussian_renderer/__init__.py:239-268) -- bit for bit in the forward, and scatter the same gradients into full-size rows."""
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

KEYS = ["means3D", "shs", "segments", "opacities", "scales", "rotations"]


def test_subset_equals_materialised_copies():
    Pk = H.pkg()
    syn = H.synthetic()
    P, W, Hh = 40_000, 400, 304
    gs, cam = syn.make_scene(P, W, Hh, seed=31)
    gs = H.to_dev(gs)
    ug = H.to_dev(syn.upstream_grads(W, Hh, 31, with_depth=True, with_segment=True, with_alpha=True))
    rs = H.settings(cam, torch.tensor([0.3, 0.2, 0.1]))
    # a bounding-box style mask: strictly ascending indices
    m = (gs["means3D"][:, 0] > -1.0) & (gs["means3D"][:, 1] < 1.5) & (torch.rand(P, generator=torch.Generator().manual_seed(0)).cuda() < 0.7)
    idx = torch.nonzero(m).flatten().to(torch.int32)
    assert 0 < idx.numel() < P

    def run(subset, tensors):
        leaves = {k: tensors[k].clone().requires_grad_(True) for k in KEYS}
        m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
        out = Pk.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"], shs=leaves["shs"],
                                        segments=leaves["segments"], scales=leaves["scales"], rotations=leaves["rotations"], subset=subset)
        color, radii, depth, alpha, segment = out
        ((color * ug["color"]).sum() + (depth * ug["depth"]).sum() + (alpha * ug["alpha"]).sum() + (segment * ug["segment"]).sum()).backward()
        return out, {k: leaves[k].grad for k in KEYS}, m2.grad

    out_c, g_c, m2_c = run(None, {k: gs[k][idx.long()].contiguous() for k in KEYS})  # the reference's way: masked copies
    out_s, g_s, m2_s = run(idx, gs)                                                  # index list, no copies
    for a, b, name in zip(out_s, out_c, ["color", "radii", "depth", "alpha", "segment"]):
        assert a.shape == b.shape and torch.equal(a, b), name  # same values in the same order: identical bits
    assert out_s[1].numel() == idx.numel()
    rest = torch.ones(P, dtype=torch.bool, device="cuda")
    rest[idx.long()] = False
    for k in KEYS:
        assert g_s[k].shape == gs[k].shape
        assert H.rel_linf(g_s[k][idx.long()], g_c[k]) <= 2e-5, k  # fp32 atomic-order noise between two backward runs
        assert float(g_s[k][rest].abs().max()) == 0.0, k          # rows outside the list: exact zeros
    assert H.rel_linf(m2_s[idx.long()], m2_c) <= 2e-5 and float(m2_s[rest].abs().max()) == 0.0


def test_subset_edge_cases():
    Pk = H.pkg()
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh = 5_000, 128, 96
    gs, cam = syn.make_scene(P, W, Hh, seed=32)
    gs = H.to_dev(gs)
    rs = H.settings(cam, torch.tensor([1.0, 1.0, 1.0]))
    e = torch.empty(0)
    full = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
    everything = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs,
                                   subset=torch.arange(P, dtype=torch.int32, device="cuda"))
    for a, b in zip(full[1:6], everything[1:6]):
        assert torch.equal(a, b)  # the identity list is the plain render
    none = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs,
                             subset=torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert none[0] == 0 and none[5].numel() == 0 and float(none[1].abs().max()) == 0.0  # like P == 0: images stay zero
    one = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs,
                            subset=torch.tensor([int(torch.nonzero(full[5] > 0)[0])], dtype=torch.int32, device="cuda"))
    assert one[5].numel() == 1 and int(one[5][0]) > 0 and one[0] > 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs,
                          subset=torch.arange(3, dtype=torch.int32))
