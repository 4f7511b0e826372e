"""gsr_adam_step against torch.optim.Adam with the reference's seven parameter groups, eps=1e-15
(scene/gaussian_model.py:166-177), several steps, a rescheduled xyz learning rate and all-zero gradient rows."""
import importlib

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P", [1, 1003, 40_000])
def test_fused_adam_matches_torch_adam(P):
    H.pkg()
    optim = importlib.import_module(H.PKG_NAME + ".optim")
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    g = torch.Generator().manual_seed(P)
    init = {"means3D": torch.randn(P, 3, generator=g) * 3, "features_dc": torch.randn(P, 1, 3, generator=g),
            "features_rest": torch.randn(P, 15, 3, generator=g) * 0.1, "segments": torch.randn(P, 2, generator=g),
            "opacities": torch.randn(P, 1, generator=g), "scales": torch.randn(P, 3, generator=g) - 3, "rotations": torch.randn(P, 4, generator=g)}
    init = {k: v.cuda() for k, v in init.items()}
    lrs = {"xyz": 1.6e-4 * 5.0, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "opacity": 0.05, "segment": 0.01, "scaling": 5e-3, "rotation": 1e-3}
    # torch side: the reference's optimizer
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    groups = [{"params": [tp[optim.GROUPS[n]]], "lr": lr, "name": n} for n, lr in lrs.items()]
    ref = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    # ours
    params = optim.FlatParameters.from_tensors(init)
    grads = mv.FlatGradients(P, "cuda", split_sh=True)
    ours = optim.FusedAdam(params, grads, lrs)
    for k, v in params.views.items():
        assert torch.equal(v, init[k].reshape(v.shape))
    for step in range(1, 7):
        xyz_lr = lrs["xyz"] * (0.9 ** step)  # update_learning_rate
        for grp in ref.param_groups:
            if grp["name"] == "xyz":
                grp["lr"] = xyz_lr
        ours.set_lr("xyz", xyz_lr)
        for k, view in grads.views.items():
            gk = torch.randn(view.shape, generator=g).cuda() * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
            if P > 4:
                gk[P // 2:] = 0.0  # invisible Gaussians: zero gradient rows, still updated through the moments
            view.copy_(gk)
            tp[k].grad = gk.clone()
        ref.step()
        ours.step()
        for k, view in params.views.items():
            want = tp[k].detach()
            err = float((view - want).abs().max())
            scale = float(want.abs().max())
            assert err <= 2e-6 * max(scale, 1.0), (step, k, err, scale)
    # moments too
    offs = params.offsets()
    for n, (off, cnt) in offs.items():
        st = ref.state[tp[n]]
        assert H.rel_linf(ours.exp_avg[off:off + cnt].view(st["exp_avg"].shape), st["exp_avg"]) <= 1e-6
        assert H.rel_linf(ours.exp_avg_sq[off:off + cnt].view(st["exp_avg_sq"].shape), st["exp_avg_sq"]) <= 1e-6


def test_fused_adam_rejects_mismatched_layouts():
    H.pkg()
    optim = importlib.import_module(H.PKG_NAME + ".optim")
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    p = optim.FlatParameters(10, "cuda")
    with pytest.raises(ValueError):
        optim.FusedAdam(p, mv.FlatGradients(10, "cuda"), {"xyz": 1e-3})
    with pytest.raises(KeyError):
        optim.FusedAdam(p, mv.FlatGradients(10, "cuda", split_sh=True), {"nonsense": 1e-3})
