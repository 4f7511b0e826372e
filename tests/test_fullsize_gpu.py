"""Full-size checks at BASELINE.json cfg3 (6 M Gaussians, 1920x1080), the configuration the headline metric is quoted on:
bit-exact integer state and image / gradient parity against the reference CUDA build when it is present, and
size-independent properties that need no oracle (sortedness, range partition, determinism, linearity, packets == dense)."""
import importlib

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg3():
    syn = H.synthetic()
    P, W, Hh, seed = syn.CONFIGS["cfg3"]
    gs, cam = syn.make_scene("cfg3")
    ug = syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True)
    rs = H.settings(cam, torch.tensor([0.05, 0.1, 0.15]))
    return H.to_dev(gs), cam, H.to_dev(ug), rs, (P, W, Hh)


def test_cfg3_parity_vs_reference_cuda(cfg3):
    if H.ref_dgr() is None:
        pytest.skip("oracle/_ref not built")
    gs, cam, ug, rs, (P, W, Hh) = cfg3
    ours = H.run_ours(gs, rs, ug)
    ref = H.run_ref(gs, rs, ug)
    torch.cuda.synchronize()
    assert ours["num_rendered"] == ref["num_rendered"] > 0
    assert torch.equal(ours["radii"], ref["radii"])
    so, sr = ours["state"], ref["state"]
    assert torch.equal(so["point_keys"], sr["point_keys"]) and torch.equal(so["point_list"], sr["point_list"])
    assert torch.equal(so["ranges"], sr["ranges"]) and torch.equal(so["n_contrib"], sr["n_contrib"])
    for k in ["color", "depth", "alpha", "segment"]:
        assert float((ours[k] - ref[k]).abs().max()) <= 1e-5, k
    for k in ["means3D", "means2D", "sh", "segments", "opacities", "scales", "rotations"]:
        a, b = ours["grads"][k], ref["grads"][k]
        assert H.rel_linf(a, b.reshape(a.shape)) <= 1e-4, k


def test_cfg3_state_properties_and_determinism(cfg3):
    gs, cam, ug, rs, (P, W, Hh) = cfg3
    a = H.run_ours(gs, rs)
    b = H.run_ours(gs, rs)
    R = a["num_rendered"]
    st = a["state"]
    keys = st["point_keys"][:R]
    assert bool((keys[1:] >= keys[:-1]).all())  # sorted by (tile, depth bits)
    tiles = (keys >> 32).to(torch.int64)
    T = ((W + 15) // 16) * ((Hh + 15) // 16)
    assert int(tiles.min()) >= 0 and int(tiles.max()) < T
    rng = st["ranges"].to(torch.int64)
    cnt = torch.bincount(tiles, minlength=T)
    assert torch.equal(rng[:, 1] - rng[:, 0], cnt)  # ranges partition the sorted list by tile
    nz = cnt > 0
    assert torch.equal(rng[nz, 0], (torch.cumsum(cnt, 0) - cnt)[nz])
    assert int(st["tiles_touched"].to(torch.int64).sum()) == R
    assert int((st["n_contrib"].to(torch.int64).view(Hh, W) > 0).sum()) > 0
    # every output of a second run is bit-identical (integer atomics only in the forward)
    for k in ["color", "depth", "alpha", "segment", "radii"]:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(st["point_list"], b["state"]["point_list"])
    # alpha is a probability, depth is non-negative, background blending keeps colour finite
    assert float(a["alpha"].min()) >= 0.0 and float(a["alpha"].max()) <= 1.0 + 1e-6
    assert bool(torch.isfinite(a["color"]).all()) and float(a["depth"].min()) >= 0.0


def test_cfg3_backward_linearity_and_packet_exchange(cfg3):
    gs, cam, ug, rs, (P, W, Hh) = cfg3
    Pk = H.pkg()
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    D = Pk.diff_gaussian_rasterization
    fwd = mv.native_view_forward(D, gs, rs)
    one = mv.FlatGradients(P, "cuda")
    two = mv.FlatGradients(P, "cuda")
    mv.native_view_backward(D, gs, rs, fwd, ug, one, first=True)
    ug2 = {k: (2.0 * v if v is not None else None) for k, v in ug.items()}
    mv.native_view_backward(D, gs, rs, fwd, ug2, two, first=True)
    # the backward is linear in the pixel gradients; scaling by 2 is exact in fp32, only the atomic order differs
    assert H.rel_linf(two.buffer, 2.0 * one.buffer) <= 2e-5
    # accumulate mode: a second pass adds the same rows
    mv.native_view_backward(D, gs, rs, fwd, ug, one, first=False)
    assert H.rel_linf(one.buffer, two.buffer) <= 2e-5
    # gradient packets + one gather pass rebuild the dense rows (SH rows from basis x dL/dRGB)
    flat = mv.FlatGradients(P, "cuda")
    flat.buffer.fill_(1.0)
    sets = [mv.native_view_backward_packets(D, gs, rs, fwd, ug2)]
    assert sets[0][2] == int((fwd[5] > 0).sum())
    mv.exchange_packets(D, None, flat, gs, sets, [[cam["campos"].cuda()]], 3, world=1)
    assert H.rel_linf(flat.packed(), two.packed()) <= 2e-5
    vis = fwd[5] > 0
    assert float(flat.views["shs"][~vis].abs().max()) == 0.0 and float(flat.views["means3D"][~vis].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------- cfg2 and cfg5
def _assert_state_equal(ours, ref):
    assert ours["num_rendered"] == ref["num_rendered"] > 0
    assert torch.equal(ours["radii"], ref["radii"])
    so, sr = ours["state"], ref["state"]
    assert torch.equal(so["point_keys"], sr["point_keys"]) and torch.equal(so["point_list"], sr["point_list"])
    assert torch.equal(so["ranges"], sr["ranges"]) and torch.equal(so["n_contrib"], sr["n_contrib"])
    assert torch.equal(so["tiles_touched"], sr["tiles_touched"])


def test_cfg2_full_parity_vs_reference_cuda():
    """BASELINE.json configs[1] at full size (3 M Gaussians, 1297x840 -- not a multiple of the tile size), forward + backward
    against the reference CUDA build: integer state bit-exact, images <= 1e-5, gradients <= 1e-4 relative."""
    if H.ref_dgr() is None:
        pytest.skip("oracle/_ref not built")
    syn = H.synthetic()
    P, W, Hh, seed = syn.CONFIGS["cfg2"]
    gs, cam = syn.make_scene("cfg2")
    ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True))
    gs = H.to_dev(gs)
    rs = H.settings(cam, torch.tensor([0.2, 0.1, 0.3]))
    ours = H.run_ours(gs, rs, ug)
    ref = H.run_ref(gs, rs, ug)
    torch.cuda.synchronize()
    _assert_state_equal(ours, ref)
    for k in ["color", "depth", "alpha", "segment"]:
        assert float((ours[k] - ref[k]).abs().max()) <= 1e-5, k
    for k in ["means3D", "means2D", "sh", "segments", "opacities", "scales", "rotations"]:
        a, b = ours["grads"][k], ref["grads"][k]
        assert H.rel_linf(a, b.reshape(a.shape)) <= 1e-4, k


def test_cfg5_forward_parity_and_fused_parts():
    """BASELINE.json configs[4] at full size (10 M Gaussians in four sub-scenes, 3840x2160, forward only): the concatenated scene
    against the reference CUDA build (integer state bit-exact, images <= 1e-5), and the concatenation-free multi-part entry
    (GsrGaussians.parts) bit-identical to rendering the concatenated tensors."""
    Pk = H.pkg()
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh, seed = syn.CONFIGS["cfg5"]
    parts = [H.to_dev(syn.make_gaussians(P // 4, seed + i, scale_P=P)) for i in range(4)]
    cam = syn.make_camera(W, Hh)
    rs = H.settings(cam, torch.tensor([0.0, 0.0, 0.0]))
    with torch.no_grad():
        fused = D._forward_parts_native(parts, rs)
        gs = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
        scene, _ = syn.make_scene("cfg5")
        assert torch.equal(gs["means3D"].cpu(), scene["means3D"])  # the recipe's merged scene
        del scene
        ours = H.run_ours(gs, rs)
    for i, k in enumerate(["num_rendered", "color", "depth", "segment", "alpha", "radii"]):
        a, b = fused[i], ours[k]
        assert (a == b) if k == "num_rendered" else torch.equal(a, b), k
    st = D.export_state(P, W, Hh, fused[6], fused[7], fused[8], fused[0])
    assert torch.equal(st["point_list"], ours["state"]["point_list"]) and torch.equal(st["n_contrib"], ours["state"]["n_contrib"])
    del fused, st
    rast = Pk.GaussianRasterizer(rs)
    c2 = rast.forward_parts(parts)[0]
    assert torch.equal(c2, ours["color"])
    if H.ref_dgr() is None:
        pytest.skip("oracle/_ref not built: fused == concatenated checked, reference comparison skipped")
    with torch.no_grad():
        ref = H.run_ref(gs, rs)
    torch.cuda.synchronize()
    _assert_state_equal(ours, ref)
    for k in ["color", "depth", "alpha", "segment"]:
        assert float((ours[k] - ref[k]).abs().max()) <= 1e-5, k


def test_count_work_matches_state(cfg3):
    """gsr_count_work (the roofline's work counters): E_b equals the sum of n_contrib, Cc <= E_b <= E <= 256 R, and every counted
    quantity is reproduced by a second run."""
    gs, cam, ug, rs, (P, W, Hh) = cfg3
    D = H.pkg().diff_gaussian_rasterization
    out = H.run_ours(gs, rs)
    w = D.count_work(P, W, Hh, out["geom"], out["binning"], out["img"], out["num_rendered"])
    assert w["E_b"] == int(out["state"]["n_contrib"].to(torch.int64).sum())
    assert 0 < w["Cc"] <= w["E_b"] <= w["E"] <= 256 * out["num_rendered"]
    assert w == D.count_work(P, W, Hh, out["geom"], out["binning"], out["img"], out["num_rendered"])
