"""trainer.NativeTrainer: the native optimisation loop against the reference's loop written with the classic
drop-in API, torch activations, the torch loss of utils/loss_utils.py and torch.optim.Adam (train.py:85-186)."""
import importlib

import pytest
import torch

import helpers as H
import test_losses_gpu as TL
import test_raw_params_gpu as TR

pytestmark = pytest.mark.gpu


def _cam_dict(cam, W, Hh):
    return dict(cam, image_width=W, image_height=Hh)


def test_native_trainer_matches_reference_style_loop():
    Pk = H.pkg()
    trainer = importlib.import_module(H.PKG_NAME + ".trainer")
    optim = importlib.import_module(H.PKG_NAME + ".optim")
    D = Pk.diff_gaussian_rasterization
    syn = H.synthetic()
    P, W, Hh = 30_000, 320, 240
    raw0, cam0 = TR._raw_scene(P, W, Hh, 21)
    cams = [_cam_dict(cam0, W, Hh), _cam_dict(syn.make_camera(W, Hh, yaw_deg=30.0), W, Hh)]
    g = torch.Generator().manual_seed(5)
    gts = [torch.rand(3, Hh, W, generator=g).cuda() for _ in cams]
    bg = torch.tensor([0.0, 0.0, 0.0])
    o = trainer.OptimizationParams()
    init = {"means3D": raw0["xyz"], "features_dc": raw0["features_dc"], "features_rest": raw0["features_rest"], "segments": raw0["segment"],
            "opacities": raw0["opacity"], "scales": raw0["scaling"], "rotations": raw0["rotation"]}
    tr = trainer.NativeTrainer(D, init, o, cameras_extent=5.0, bg=bg)
    tr.active_sh_degree = 3
    # reference-style loop
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    lrs = {"xyz": o.position_lr_init, "f_dc": o.feature_lr, "f_rest": o.feature_lr / 20.0, "opacity": o.opacity_lr, "segment": o.segment_lr,
           "scaling": o.scaling_lr, "rotation": o.rotation_lr}
    topt = torch.optim.Adam([{"params": [tp[optim.GROUPS[n]]], "lr": lr, "name": n} for n, lr in lrs.items()], lr=0.0, eps=1e-15)
    sched = trainer.get_expon_lr_func(o.position_lr_init, o.position_lr_final, lr_delay_mult=o.position_lr_delay_mult,
                                      max_steps=o.position_lr_max_steps)
    acc = torch.zeros(P, 1, device="cuda")
    losses_native, losses_ref = [], []
    for it in range(1, 7):
        cam, gt = cams[it % 2], gts[it % 2]
        losses_native.append(float(tr.train_step([cam], [gt])))
        for grp in topt.param_groups:
            if grp["name"] == "xyz":
                grp["lr"] = sched(it)
        raw_t = {"xyz": tp["means3D"], "features_dc": tp["features_dc"], "features_rest": tp["features_rest"], "segment": tp["segments"],
                 "opacity": tp["opacities"], "scaling": tp["scales"], "rotation": tp["rotations"]}
        act = TR._activate(raw_t)
        m2 = torch.zeros_like(tp["means3D"], requires_grad=True)
        rs = H.settings(cam, bg)
        color, radii, depth, alpha, segment = Pk.GaussianRasterizer(rs)(means3D=act["means3D"], means2D=m2, opacities=act["opacities"],
                                                                        shs=act["shs"], segments=act["segments"], scales=act["scales"],
                                                                        rotations=act["rotations"])
        loss = TL._ref_loss(color, gt, o.lambda_dssim)[0]
        loss.backward()
        losses_ref.append(float(loss.detach()))
        vis = radii > 0
        acc[vis] += torch.norm(m2.grad[vis, :2], dim=-1, keepdim=True)  # add_densification_stats
        topt.step()
        topt.zero_grad(set_to_none=True)
    for a, b in zip(losses_native, losses_ref):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (losses_native, losses_ref)
    # Adam with eps = 1e-15 turns ANY non-zero gradient into a full-size step, so a Gaussian whose gradient is exactly zero in one
    # run and 1e-12 in the other (an alpha < 1/255 decision flipped by fp32 atomic-order noise in an earlier step) moves by ~lr:
    # compare robustly -- all but a 2e-3 fraction of the entries within 1e-4 of their scale, and nothing further than a few steps
    for k, view in tr.params.views.items():
        want = tp[k].detach().reshape(view.shape)
        diff = (view - want).abs()
        scale = max(1.0, float(want.abs().max()))
        frac_bad = float((diff > 1e-4 * scale).float().mean())
        assert frac_bad <= 2e-3, (k, frac_bad)
        assert float(diff.max()) <= 6 * 0.05 + 1e-6, (k, float(diff.max()))
    assert H.rel_linf(tr.stats.xyz_gradient_accum, acc) <= 1e-3
    tr.close()


def test_native_trainer_depth_supervised_step_matches_reference_style_loop():
    """The using_depth step (train.py:115-121, 'localrf'): render -> L1 + SSIM + compute_depth_loss(1 / depth.clamp(1e-6), prior) ->
    backward, natively (gsr_depth_loss fused with render()'s depth normalisation) against the same step written with torch ops."""
    import test_depth_loss_gpu as TD

    Pk = H.pkg()
    trainer = importlib.import_module(H.PKG_NAME + ".trainer")
    D = Pk.diff_gaussian_rasterization
    P, W, Hh = 30_000, 320, 240
    raw0, cam0 = TR._raw_scene(P, W, Hh, 23)
    cam = _cam_dict(cam0, W, Hh)
    g = torch.Generator().manual_seed(9)
    gt, prior = torch.rand(3, Hh, W, generator=g).cuda(), torch.rand(1, Hh, W, generator=g).cuda()
    o = trainer.OptimizationParams()
    init = {"means3D": raw0["xyz"], "features_dc": raw0["features_dc"], "features_rest": raw0["features_rest"], "segments": raw0["segment"],
            "opacities": raw0["opacity"], "scales": raw0["scaling"], "rotations": raw0["rotation"]}
    tr = trainer.NativeTrainer(D, init, o, cameras_extent=5.0, bg=torch.zeros(3))
    tr.active_sh_degree = 3
    p0 = {k: v.clone() for k, v in tr.params.views.items()}
    native = float(tr.train_step([cam], [gt], [prior]))
    # the same step in torch: activations -> classic rasterizer -> render()'s depth normalisation -> the reference's losses
    tp = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    raw_t = {"xyz": tp["means3D"], "features_dc": tp["features_dc"], "features_rest": tp["features_rest"], "segment": tp["segments"],
             "opacity": tp["opacities"], "scaling": tp["scales"], "rotation": tp["rotations"]}
    act = TR._activate(raw_t)
    rs = H.settings(cam, torch.zeros(3))
    color, radii, depth, alpha, segment = Pk.GaussianRasterizer(rs)(means3D=act["means3D"], means2D=torch.zeros_like(tp["means3D"]),
                                                                    opacities=act["opacities"], shs=act["shs"], segments=act["segments"],
                                                                    scales=act["scales"], rotations=act["rotations"])
    dn = depth / (depth.max() + 1e-5)  # gaussian_renderer/__init__.py:375
    loss = TL._ref_loss(color, gt, o.lambda_dssim)[0] + TD._torch_depth_loss(1 / dn.clamp(1e-6), prior, o.lambda_depth)
    loss.backward()
    assert abs(native - float(loss.detach())) <= 2e-5 * max(1.0, abs(float(loss.detach())))
    # the gradient the native step produced (it is still in the flat buffer) against autograd's
    for k, view in tr.grads.views.items():
        assert H.rel_linf(view, tp[k].grad.reshape(view.shape)) <= 2e-4, k
    assert any(not torch.equal(tr.params.views[k], p0[k]) for k in p0)  # and the step was applied
    tr.close()


def test_native_trainer_densifies_and_keeps_running():
    Pk = H.pkg()
    trainer = importlib.import_module(H.PKG_NAME + ".trainer")
    D = Pk.diff_gaussian_rasterization
    P, W, Hh = 20_000, 256, 192
    raw0, cam0 = TR._raw_scene(P, W, Hh, 22)
    cam = _cam_dict(cam0, W, Hh)
    gt = torch.rand(3, Hh, W, generator=torch.Generator().manual_seed(1)).cuda()
    o = trainer.OptimizationParams(densify_from_iter=2, densification_interval=3, opacity_reset_interval=4, densify_until_iter=100)
    init = {"means3D": raw0["xyz"], "features_dc": raw0["features_dc"], "features_rest": raw0["features_rest"], "segments": raw0["segment"],
            "opacities": raw0["opacity"], "scales": raw0["scaling"], "rotations": raw0["rotation"]}
    tr = trainer.NativeTrainer(D, init, o, cameras_extent=5.0, bg=torch.zeros(3))
    sizes, vals = [], []
    for it in range(1, 8):
        vals.append(float(tr.train_step([cam], [gt])))
        sizes.append(tr.P)
    assert all(torch.isfinite(torch.tensor(vals))) and len(set(sizes)) > 1  # the model was rebuilt at iterations 3 and 6
    assert tr.grads.buffer.numel() == tr.params.buffer.numel() == tr.opt.exp_avg.numel() >= 61 * tr.P
    assert float(torch.sigmoid(tr.params.views["opacities"]).max()) <= 0.05  # opacity reset at iteration 4, a few steps ago
    tr.close()
