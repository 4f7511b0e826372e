"""GPU tests of the drop-in Python surface (GaussianRasterizationSettings / GaussianRasterizer / rasterize_gaussians /
markVisible) and of the edge cases of SURVEY.md Appendix D that do not need the reference build."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _mk(P, W, Hh, seed):
    syn = H.synthetic()
    gs, cam = syn.make_scene(P, W, Hh, seed=seed)
    return H.to_dev(gs), cam


def test_module_forward_backward_autograd():
    Pk = H.pkg()
    gs, cam = _mk(20_000, 256, 192, 3)
    rs = H.settings(cam, torch.zeros(3))
    leaves = {k: gs[k].clone().requires_grad_(True) for k in ["means3D", "shs", "segments", "opacities", "scales", "rotations"]}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rast = Pk.GaussianRasterizer(rs)
    color, radii, depth, alpha, segment = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves["shs"],
                                               segments=leaves["segments"], scales=leaves["scales"], rotations=leaves["rotations"])
    assert color.shape == (3, 192, 256) and depth.shape == (1, 192, 256) and alpha.shape == (1, 192, 256)
    assert segment.shape == (2, 192, 256) and radii.shape == (20_000,) and radii.dtype == torch.int32
    loss = color.mean() + 0.1 * (depth / (depth.max() + 1e-5)).mean() + 0.05 * segment.mean()
    loss.backward()
    for k, v in leaves.items():
        assert v.grad is not None and v.grad.shape == v.shape and torch.isfinite(v.grad).all(), k
    assert means2D.grad is not None and float(means2D.grad[:, 2].abs().max()) == 0.0
    vis = radii > 0
    assert float(leaves["means3D"].grad[~vis].abs().max()) == 0.0  # dense grads, zero rows for invisible Gaussians
    assert float(leaves["shs"].grad[~vis].abs().max()) == 0.0
    assert float(leaves["opacities"].grad[vis].abs().max()) > 0.0


def test_validation_messages():
    Pk = H.pkg()
    gs, cam = _mk(100, 64, 64, 1)
    rast = Pk.GaussianRasterizer(H.settings(cam, torch.zeros(3)))
    m2 = torch.zeros_like(gs["means3D"])
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=gs["means3D"], means2D=m2, opacities=gs["opacities"], scales=gs["scales"], rotations=gs["rotations"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        rast(means3D=gs["means3D"], means2D=m2, opacities=gs["opacities"], shs=gs["shs"], scales=gs["scales"])
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rast(means3D=gs["means3D"].reshape(-1), means2D=m2, opacities=gs["opacities"], shs=gs["shs"], scales=gs["scales"],
             rotations=gs["rotations"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        rast(means3D=gs["means3D"].cpu(), means2D=m2, opacities=gs["opacities"], shs=gs["shs"], scales=gs["scales"],
             rotations=gs["rotations"])


def test_empty_input_returns_zero_images():
    Pk = H.pkg()
    _, cam = _mk(10, 64, 48, 1)
    rs = H.settings(cam, torch.ones(3))
    e = lambda *s: torch.zeros(*s, device="cuda")
    rast = Pk.GaussianRasterizer(rs)
    color, radii, depth, alpha, segment = rast(means3D=e(0, 3), means2D=e(0, 3), opacities=e(0, 1), shs=e(0, 16, 3), segments=e(0, 2),
                                               scales=e(0, 3), rotations=e(0, 4))
    assert float(color.abs().max()) == 0.0 and radii.numel() == 0  # background NOT applied when P == 0 (rasterize_points.cu:87)


def test_nothing_visible_gives_background():
    gs, cam = _mk(1000, 96, 64, 2)
    gs["means3D"][:, 2] = -50.0  # all behind the camera
    rs = H.settings(cam, torch.tensor([0.3, 0.6, 0.9]))
    ug = H.to_dev(H.synthetic().upstream_grads(96, 64, 2))
    out = H.run_ours(gs, rs, ug)
    assert out["num_rendered"] == 0 and int(out["radii"].abs().sum()) == 0
    assert torch.allclose(out["color"][:, 0, 0], torch.tensor([0.3, 0.6, 0.9], device="cuda"))
    assert float(out["depth"].abs().max()) == 0 and float(out["alpha"].abs().max()) == 0 and float(out["segment"].abs().max()) == 0
    assert int(out["state"]["n_contrib"].sum()) == 0
    for k, v in out["grads"].items():
        if v is not None:
            assert float(v.abs().max()) == 0.0, k


def test_segments_none_renders_zero_segment():
    Pk = H.pkg()
    gs, cam = _mk(5000, 128, 96, 4)
    rast = Pk.GaussianRasterizer(H.settings(cam, torch.zeros(3)))
    color, radii, depth, alpha, segment = rast(means3D=gs["means3D"], means2D=torch.zeros_like(gs["means3D"]), opacities=gs["opacities"],
                                               shs=gs["shs"], scales=gs["scales"], rotations=gs["rotations"])
    assert segment.shape == (2, 96, 128) and float(segment.abs().max()) == 0.0 and float(color.abs().max()) > 0


def test_mark_visible_matches_view_depth():
    Pk = H.pkg()
    gs, cam = _mk(50_000, 64, 64, 6)
    rast = Pk.GaussianRasterizer(H.settings(cam, torch.zeros(3)))
    vis = rast.markVisible(gs["means3D"])
    O = H.cpu_oracle()
    exp = O.mark_visible(gs["means3D"].cpu().numpy(), cam["viewmatrix"].numpy(), cam["projmatrix"].numpy())
    got = vis.cpu().numpy()
    assert got.dtype == np.bool_ and (got != exp).mean() <= 1e-4  # fp32 FMA vs non-FMA at the z == 0.2 boundary


def test_mark_visible_bit_exact_vs_reference_cuda():
    """markVisible against the reference's own _C.mark_visible (rasterizer_impl.cu:141-153) on the same device: bit for bit, at
    a plumbing size and at cfg3 size, for a rotated camera too."""
    C = H.ref_dgr()
    if C is None:
        pytest.skip("oracle/_ref not built")
    Pk = H.pkg()
    syn = H.synthetic()
    for P, seed, yaw in [(50_000, 6, 0.0), (6_000_000, 2, 135.0), (1, 3, 0.0), (257, 4, 45.0)]:
        gs = syn.make_gaussians(P, seed)
        cam = syn.make_camera(320, 200, yaw_deg=yaw)
        m = gs["means3D"].cuda()
        rs = H.settings(cam, torch.zeros(3))
        got = Pk.GaussianRasterizer(rs).markVisible(m)
        exp = C.mark_visible(m, rs.viewmatrix, rs.projmatrix)
        assert got.dtype == torch.bool and got.shape == exp.shape and torch.equal(got, exp), (P, yaw)
        assert 0 < int(got.sum()) <= P


def test_prefiltered_culled_point_raises():
    gs, cam = _mk(1000, 64, 64, 8)
    rs = H.settings(cam, torch.zeros(3), prefiltered=True)
    with pytest.raises(RuntimeError, match="prefiltered"):
        H.run_ours(gs, rs)


def test_debug_mode_runs():
    gs, cam = _mk(2000, 96, 64, 9)
    rs = H.settings(cam, torch.zeros(3), debug=True)
    out = H.run_ours(gs, rs, H.to_dev(H.synthetic().upstream_grads(96, 64, 9)))
    assert torch.isfinite(out["color"]).all()


def test_tile_lists_are_sorted_partition_full_size():
    """Size-independent properties at a BASELINE-size case (cfg2 shape, 3M Gaussians, 1297x840): ranges partition
    [0,R), keys are non-decreasing, sum(tiles_touched) == R, every listed Gaussian is visible."""
    gs, cam = _mk(3_000_000, 1297, 840, 1)
    rs = H.settings(cam, torch.zeros(3))
    out = H.run_ours(gs, rs)
    st, R = out["state"], out["num_rendered"]
    assert int(st["tiles_touched"].long().sum()) == R and R > 0
    keys = st["point_keys"]
    assert bool((keys[1:] >= keys[:-1]).all())
    rng = st["ranges"].long()
    ne = rng[(rng[:, 1] - rng[:, 0]) > 0]
    assert int((ne[:, 1] - ne[:, 0]).sum()) == R
    assert bool((ne[1:, 0] == ne[:-1, 1]).all()) and int(ne[0, 0]) == 0 and int(ne[-1, 1]) == R
    assert bool((out["radii"][st["point_list"].long()] > 0).all())
    assert bool((st["n_contrib"].view(840, 1297).long() <= (rng[:, 1] - rng[:, 0]).max()).all())
    assert torch.isfinite(out["color"]).all() and float(out["alpha"].max()) <= 1.0 + 1e-4


def test_accumulate_mode_sums_views_into_flat_buffer():
    """gsr_backward(accumulate): two views accumulated into one flat buffer == sum of the two dense gradients; untouched rows
    keep what the first (overwrite) view wrote; densification inputs stay per view."""
    import importlib

    Pk = H.pkg()
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh = 30_000, 320, 240
    gs, cam0 = syn.make_scene(P, W, Hh, seed=61, yaw_deg=0.0)
    cam1 = syn.make_camera(W, Hh, yaw_deg=45.0)
    gs = H.to_dev(gs)
    ug = H.to_dev(syn.upstream_grads(W, Hh, 61, with_depth=True, with_segment=True, with_alpha=True))
    bg = torch.tensor([0.1, 0.2, 0.3])
    flat = mv.FlatGradients(P, "cuda")
    flat.buffer.fill_(123.0)  # stale content must be overwritten by the first view
    e = torch.empty(0)
    dense = []
    for vi, cam in enumerate([cam0, cam1]):
        rs = H.settings(cam, bg)
        fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
        m2 = torch.zeros(P, 3, device="cuda")
        mv.native_view_backward(D, gs, rs, fwd, ug, flat, first=(vi == 0), means2D_grad=m2)
        R, color, depth, segment, alpha, radii, geom, binb, img = fwd
        g = D._backward_native(rs, gs["means3D"], radii, e, gs["segments"], gs["scales"], gs["rotations"], e, ug["color"], ug["segment"],
                               ug["depth"], ug["alpha"], gs["shs"], geom, R, binb, img, alpha)
        dense.append(g)
        # two separate backward runs differ by the order of the fp32 atomics in the compositing backward: ~1e-6 relative
        assert H.rel_linf(m2, g["means2D"]) <= 2e-5
    names = {"means3D": "means3D", "shs": "sh", "segments": "segments", "opacities": "opacities", "scales": "scales", "rotations": "rotations"}
    for leaf, nat in names.items():
        exp = dense[0][nat] + dense[1][nat]
        assert H.rel_linf(flat.views[leaf], exp) <= 2e-5, leaf
    assert sum(c for _, c in flat.offsets().values()) == 61 * P and flat.buffer.numel() >= 61 * P


def test_gradient_packets_rebuild_dense_rows():
    """gsr_backward_packets + gsr_gather_packets (the multi-GPU exchange format): rebuilding dense rows from
    a view's packets reproduces the dense backward (SH rows are rebuilt as basis(direction) x dL/dRGB) up to the fp32
    atomic-order noise between two backward runs; two views sum; the gather form overwrites every row."""
    import importlib

    Pk = H.pkg()
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh = 30_000, 320, 240
    gs, cam0 = syn.make_scene(P, W, Hh, seed=71, yaw_deg=0.0)
    cams = [cam0, syn.make_camera(W, Hh, yaw_deg=90.0)]
    gs = H.to_dev(gs)
    ug = H.to_dev(syn.upstream_grads(W, Hh, 71, with_depth=True, with_segment=True, with_alpha=True))
    bg = torch.tensor([0.1, 0.2, 0.3])
    e = torch.empty(0)
    flat = mv.FlatGradients(P, "cuda")
    sets, dense, campos = [], [], []
    for cam in cams:
        rs = H.settings(cam, bg)
        fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
        m2 = torch.zeros(P, 3, device="cuda")
        sets.append(mv.native_view_backward_packets(D, gs, rs, fwd, ug, means2D_grad=m2))
        R, color, depth, segment, alpha, radii, geom, binb, img = fwd
        g = D._backward_native(rs, gs["means3D"], radii, e, gs["segments"], gs["scales"], gs["rotations"], e, ug["color"], ug["segment"],
                               ug["depth"], ug["alpha"], gs["shs"], geom, R, binb, img, alpha)
        dense.append(g)
        campos.append(cam["campos"].cuda())
        assert int(sets[-1][1]) == sets[-1][2] == int((radii > 0).sum())
        pk, bits, first = D.packet_blob_views(sets[-1][0], P)
        vis_ids = torch.nonzero(radii > 0).flatten()
        assert pk.shape[1] == 16  # 64-byte packets, no id word: ascending Gaussian order + the visibility index address them
        vis = torch.zeros(32 * bits.numel(), dtype=torch.bool, device="cuda")
        vis[:P] = radii > 0
        grp = vis.view(-1, 32)
        want_bits = (grp.long() << torch.arange(32, device="cuda")).sum(1)
        assert torch.equal(bits.long() & 0xFFFFFFFF, want_bits)
        before = torch.cumsum(grp.sum(1), 0) - grp.sum(1)  # packets before each group of 32
        nz = grp.any(1)
        assert torch.equal(first[nz].long(), before[nz])
        assert H.rel_linf(m2, g["means2D"]) <= 2e-5
    names = {"means3D": "means3D", "shs": "sh", "segments": "segments", "opacities": "opacities", "scales": "scales", "rotations": "rotations"}
    # one view; the gather pass writes every row, so stale contents must not survive
    flat.buffer.fill_(123.0)
    mv.exchange_packets(D, None, flat, gs, sets[:1], [campos[:1]], 3, world=1)
    for leaf, nat in names.items():
        assert H.rel_linf(flat.views[leaf], dense[0][nat]) <= 2e-5, leaf
    # rows of Gaussians invisible in the view stay exactly zero
    inv = ~(dense[0]["means2D"].abs().sum(1) > 0)
    assert float(flat.views["shs"][inv].abs().max()) == 0.0
    # two views of one rank: the sum
    flat.buffer.fill_(-7.0)
    mv.exchange_packets(D, None, flat, gs, sets, [campos], 3, world=1)
    for leaf, nat in names.items():
        assert H.rel_linf(flat.views[leaf], dense[0][nat] + dense[1][nat]) <= 2e-5, leaf
    # sticky capacity: the second step's blobs are produced at the exchange's capacity and gathered without repacking
    st = {}
    mv.exchange_packets(D, None, flat, gs, sets, [campos], 3, world=1, state=st)
    assert st["cap"] >= max(s[2] for s in sets)
    sets2 = []
    for cam in cams:
        rs = H.settings(cam, bg)
        fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
        sets2.append(mv.native_view_backward_packets(D, gs, rs, fwd, ug, capacity=st["cap"]))
    assert all(D.packet_blob_capacity(s[0], P) == st["cap"] for s in sets2)
    flat.buffer.fill_(5.0)
    mv.exchange_packets(D, None, flat, gs, sets2, [campos], 3, world=1, state=st)
    for leaf, nat in names.items():
        assert H.rel_linf(flat.views[leaf], dense[0][nat] + dense[1][nat]) <= 2e-5, leaf
    # the stacked-blob entry (what follows an ncclAllGather of equally sized blobs): gsr_gather_packets
    flat.buffer.fill_(9.0)
    D.gather_packets(gs["means3D"], torch.stack(campos).contiguous(), 3, gs["shs"].size(1), torch.stack([s[0] for s in sets2]), flat.backward_out())
    for leaf, nat in names.items():
        assert H.rel_linf(flat.views[leaf], dense[0][nat] + dense[1][nat]) <= 2e-5, leaf


def test_packet_capacity_comes_from_the_forward_it_belongs_to():
    """Several views forwarded before their backwards (or an eval render in between): the packet buffer of a view is sized by
    THAT view's visible count, and a buffer that is too small is refused (GSR_ERR_OVERFLOW) instead of dropping packets."""
    import importlib

    Pk = H.pkg()
    D = Pk.diff_gaussian_rasterization
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    syn = H.synthetic()
    gs = H.to_dev(syn.make_gaussians(40_000, 21))
    cam_a, cam_b = syn.make_camera(160, 120, yaw_deg=90.0, radius=40.0), syn.make_camera(160, 120)  # a: far away, sees everything
    rs_a, rs_b = H.settings(cam_a, torch.zeros(3)), H.settings(cam_b, torch.zeros(3))
    ug = H.to_dev(syn.upstream_grads(160, 120, 5))
    fwd_a = mv.native_view_forward(D, gs, rs_a)
    fwd_b = mv.native_view_forward(D, gs, rs_b)  # the thread's "last forward" is now b
    Va, Vb = int((fwd_a[5] > 0).sum()), int((fwd_b[5] > 0).sum())
    assert fwd_a[0].num_visible == Va and fwd_b[0].num_visible == Vb and Va > Vb
    blob, count, nvis = mv.native_view_backward_packets(D, gs, rs_a, fwd_a, ug)
    assert nvis == Va == int(count.item()) and D.packet_blob_capacity(blob, 40_000) >= Va
    flat, dense = mv.FlatGradients(40_000, "cuda"), mv.FlatGradients(40_000, "cuda")
    mv.exchange_packets(D, None, flat, gs, [(blob, count, nvis)], [[cam_a["campos"].cuda()]], 3, world=1)
    mv.native_view_backward(D, gs, rs_a, fwd_a, ug, dense, first=True)
    assert H.rel_linf(flat.packed(), dense.packed()) <= 2e-5
    with pytest.raises(RuntimeError, match="do not fit"):
        D._backward_packets_native(rs_a, gs["means3D"], fwd_a[5], gs["segments"], gs["scales"], gs["rotations"], ug["color"], None, ug.get("depth"),
                                   None, gs["shs"], fwd_a[6], fwd_a[0], fwd_a[7], fwd_a[8], fwd_a[4], capacity=max(Va // 2, 1))


def test_unaligned_views_of_a_flat_buffer_are_accepted():
    """Contiguous inputs that start at an odd offset of a larger allocation (the reference's scalar loads accept them)."""
    Pk = H.pkg()
    gs, cam = _mk(3000, 96, 64, 12)
    rs = H.settings(cam, torch.zeros(3))
    ref = H.run_ours(gs, rs, export=False)
    shifted = {}
    for k, v in gs.items():
        buf = torch.empty(v.numel() + 3, device="cuda")
        buf[1:1 + v.numel()] = v.reshape(-1)
        shifted[k] = buf[1:1 + v.numel()].view(v.shape)
        assert shifted[k].data_ptr() % 16 != 0 and shifted[k].is_contiguous()
    out = H.run_ours(shifted, rs, export=False)
    for k in ["color", "depth", "alpha", "segment", "radii"]:
        assert torch.equal(out[k], ref[k]), k


@pytest.mark.parametrize("N", [1, 5, 29])
def test_runtime_class_count_equals_channel_pairs_of_the_two_class_path(N):
    """num_class at run time (the reference compiles NUM_CLASS = 2 in, config.h:16, although its ModelParams default to 29): an
    N-channel segment render equals the stack of 2-class renders of its channel pairs -- bit for bit in the forward -- and its
    gradients equal the colour/depth/alpha gradient of one 2-class backward plus the segment-only gradients of every pair."""
    Pk = H.pkg()
    syn = H.synthetic()
    P, W, Hh = 20_000, 208, 144
    gs = H.to_dev(syn.make_gaussians(P, 33, num_class=2))
    g = torch.Generator().manual_seed(N)
    segN = torch.sigmoid(torch.randn(P, N, generator=g)).cuda()
    cam = syn.make_camera(W, Hh)
    rs = H.settings(cam, torch.tensor([0.1, 0.0, 0.2]))
    up = H.to_dev(syn.upstream_grads(W, Hh, 3, with_depth=True, with_alpha=True))
    up_seg = (torch.randn(N, Hh, W, generator=g) / (N * W * Hh)).cuda()
    names = ["means3D", "opacities", "shs", "scales", "rotations"]

    def run(seg, gseg, with_color):
        lv = {k: gs[k].clone().requires_grad_(True) for k in names}
        sg = seg.clone().requires_grad_(True)
        m2 = torch.zeros(P, 3, device="cuda", requires_grad=True)
        color, radii, depth, alpha, segment = Pk.GaussianRasterizer(rs)(means3D=lv["means3D"], means2D=m2, opacities=lv["opacities"], shs=lv["shs"],
                                                                        segments=sg, scales=lv["scales"], rotations=lv["rotations"])
        loss = (segment * gseg).sum()
        if with_color:
            loss = loss + (color * up["color"]).sum() + (depth * up["depth"]).sum() + (alpha * up["alpha"]).sum()
        loss.backward()
        grads = {k: lv[k].grad for k in names}
        grads["means2D"], grads["segments"] = m2.grad, sg.grad
        return (color, radii, depth, alpha, segment), grads

    outN, gN = run(segN, up_seg, True)
    assert outN[4].shape == (N, Hh, W) and gN["segments"].shape == (P, N)
    want = {k: torch.zeros_like(v) for k, v in gN.items() if k != "segments"}
    want_seg = torch.zeros(P, N, device="cuda")
    for c0 in range(0, N, 2):
        n = min(2, N - c0)
        seg2 = torch.zeros(P, 2, device="cuda")
        seg2[:, :n] = segN[:, c0:c0 + n]
        g2 = torch.zeros(2, Hh, W, device="cuda")
        g2[:n] = up_seg[c0:c0 + n]
        out2, gr2 = run(seg2, g2, c0 == 0)
        assert torch.equal(out2[4][:n], outN[4][c0:c0 + n]), c0  # every pair: bit-identical to the 2-class render
        if c0 == 0:
            for i in range(4):
                assert torch.equal(out2[i], outN[i])
        for k in want:
            want[k] += gr2[k]
        want_seg[:, c0:c0 + n] = gr2["segments"][:, :n]
    for k in want:
        assert H.rel_linf(gN[k], want[k]) <= 3e-5, k
    assert H.rel_linf(gN["segments"], want_seg) <= 3e-5
