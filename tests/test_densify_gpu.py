"""optim.densify_and_prune / gsr_select_rows against the reference's algorithm restated with torch tensors and
a real torch.optim.Adam, following scene/gaussian_model.py:377-515 statement by statement (prune_points / _prune_optimizer,
cat_tensors_to_optimizer / densification_postfix, densify_and_clone, densify_and_split, densify_and_prune). Same RNG state =>
the same torch.normal samples, so parameters and Adam moments must come out BIT-identical."""
import importlib

import pytest
import torch
from torch import nn

import helpers as H

pytestmark = pytest.mark.gpu

NAMES = {"xyz": "means3D", "f_dc": "features_dc", "f_rest": "features_rest", "opacity": "opacities", "segment": "segments",
         "scaling": "scales", "rotation": "rotations"}


class RefModel:
    """The slice of GaussianModel that densification touches."""

    def __init__(self, init, lrs, percent_dense):
        self.p = {n: nn.Parameter(init[k].clone().requires_grad_(True)) for n, k in NAMES.items()}
        self.optimizer = torch.optim.Adam([{"params": [self.p[n]], "lr": lrs[n], "name": n} for n in NAMES], lr=0.0, eps=1e-15)
        self.percent_dense = percent_dense
        P = init["means3D"].shape[0]
        self.xyz_gradient_accum = torch.zeros((P, 1), device="cuda")
        self.denom = torch.zeros((P, 1), device="cuda")
        self.max_radii2D = torch.zeros(P, device="cuda")

    get_scaling = property(lambda s: torch.exp(s.p["scaling"]))
    get_opacity = property(lambda s: torch.sigmoid(s.p["opacity"]))

    def _prune_optimizer(self, mask):  # :377-393
        for group in self.optimizer.param_groups:
            st = self.optimizer.state.get(group["params"][0], None)
            if st is not None:
                st["exp_avg"] = st["exp_avg"][mask]
                st["exp_avg_sq"] = st["exp_avg_sq"][mask]
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
                self.optimizer.state[group["params"][0]] = st
            else:
                group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            self.p[group["name"]] = group["params"][0]

    def prune_points(self, mask):  # :395-412
        valid = ~mask
        self._prune_optimizer(valid)
        self.xyz_gradient_accum = self.xyz_gradient_accum[valid]
        self.denom = self.denom[valid]
        self.max_radii2D = self.max_radii2D[valid]

    def densification_postfix(self, new):  # :414-463
        for group in self.optimizer.param_groups:
            ext = new[group["name"]]
            st = self.optimizer.state.get(group["params"][0], None)
            if st is not None:
                st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
                st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
                self.optimizer.state[group["params"][0]] = st
            else:
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
            self.p[group["name"]] = group["params"][0]
        P = self.p["xyz"].shape[0]
        self.xyz_gradient_accum = torch.zeros((P, 1), device="cuda")
        self.denom = torch.zeros((P, 1), device="cuda")
        self.max_radii2D = torch.zeros(P, device="cuda")

    def densify_and_split(self, grads, thr, extent, build_rotation, N=2):  # :465-481
        n_init = self.p["xyz"].shape[0]
        padded = torch.zeros(n_init, device="cuda")
        padded[:grads.shape[0]] = grads.squeeze()
        sel = torch.where(padded >= thr, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values > self.percent_dense * extent)
        stds = self.get_scaling[sel].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device="cuda")
        samples = torch.normal(mean=means, std=stds)
        rots = build_rotation(self.p["rotation"][sel]).repeat(N, 1, 1)
        new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.p["xyz"][sel].repeat(N, 1),
               "scaling": torch.log(self.get_scaling[sel].repeat(N, 1) / (0.8 * N)), "rotation": self.p["rotation"][sel].repeat(N, 1),
               "f_dc": self.p["f_dc"][sel].repeat(N, 1, 1), "f_rest": self.p["f_rest"][sel].repeat(N, 1, 1),
               "opacity": self.p["opacity"][sel].repeat(N, 1), "segment": self.p["segment"][sel].repeat(N, 1)}
        self.densification_postfix(new)
        self.prune_points(torch.cat((sel, torch.zeros(N * sel.sum(), device="cuda", dtype=bool))))

    def densify_and_clone(self, grads, thr, extent):  # :483-498
        sel = torch.where(torch.norm(grads, dim=-1) >= thr, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * extent)
        self.densification_postfix({n: self.p[n][sel] for n in NAMES})

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, build_rotation):  # :500-515
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, build_rotation)
        prune_mask = (self.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size
            big_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
        self.prune_points(prune_mask)


@pytest.mark.parametrize("P,max_screen_size", [(20_000, None), (20_000, 20), (257, 20)])
def test_densify_and_prune_matches_reference_algorithm(P, max_screen_size):
    H.pkg()
    optim = importlib.import_module(H.PKG_NAME + ".optim")
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    g = torch.Generator().manual_seed(P + (max_screen_size or 0))
    init = {"means3D": torch.randn(P, 3, generator=g) * 3, "features_dc": torch.randn(P, 1, 3, generator=g),
            "features_rest": torch.randn(P, 15, 3, generator=g) * 0.1, "segments": torch.randn(P, 2, generator=g),
            "opacities": torch.randn(P, 1, generator=g) * 3, "scales": torch.randn(P, 3, generator=g) * 1.2 - 3.0,
            "rotations": torch.randn(P, 4, generator=g)}
    init = {k: v.cuda() for k, v in init.items()}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "segment": 0.01, "scaling": 5e-3, "rotation": 1e-3}
    extent, pd, thr, min_op = 5.0, 0.01, 0.0002, 0.005
    ref = RefModel(init, lrs, pd)
    params = optim.FlatParameters.from_tensors(init)
    grads = mv.FlatGradients(P, "cuda", split_sh=True)
    ours = optim.FusedAdam(params, grads, lrs)
    # two optimiser steps so that the moments are non-trivial (Adam parity itself is tests/test_optim_gpu.py; copy torch's state over
    # so that the comparison below can be bit-exact)
    for _ in range(2):
        for n, k in NAMES.items():
            gk = torch.randn(init[k].shape, generator=g).cuda() * 1e-3
            ref.p[n].grad = gk
            grads.views[k].copy_(gk)
        ref.optimizer.step()
        ours.step()
    offs = params.offsets()
    for n, k in NAMES.items():
        o, c = offs[k]
        params.buffer[o:o + c].copy_(ref.p[n].detach().reshape(-1))
        ours.exp_avg[o:o + c].copy_(ref.optimizer.state[ref.p[n]]["exp_avg"].reshape(-1))
        ours.exp_avg_sq[o:o + c].copy_(ref.optimizer.state[ref.p[n]]["exp_avg_sq"].reshape(-1))
    acc = torch.rand(P, 1, generator=g).cuda() * 0.002
    den = torch.randint(0, 4, (P, 1), generator=g).float().cuda()  # zeros -> NaN -> 0, as in the reference
    ref.xyz_gradient_accum, ref.denom = acc.clone(), den.clone()
    ref.max_radii2D = torch.rand(P, generator=g).cuda() * 40
    torch.manual_seed(1234)
    ref.densify_and_prune(thr, min_op, extent, max_screen_size, optim.build_rotation)
    torch.manual_seed(1234)
    new_params, new_grads, index = optim.densify_and_prune(params, ours, acc.clone(), den.clone(), thr, min_op, extent, max_screen_size,
                                                           percent_dense=pd)
    Pn = ref.p["xyz"].shape[0]
    assert new_params.views["means3D"].shape[0] == Pn == index.numel() and Pn != P
    assert new_grads.buffer.numel() == new_params.buffer.numel() >= 61 * Pn
    noffs = new_params.offsets()
    for n, k in NAMES.items():
        assert torch.equal(new_params.views[k], ref.p[n].detach().reshape(new_params.views[k].shape)), n
        st = ref.optimizer.state[ref.p[n]]
        o, c = noffs[k]
        assert torch.equal(ours.exp_avg[o:o + c], st["exp_avg"].reshape(-1)), n
        assert torch.equal(ours.exp_avg_sq[o:o + c], st["exp_avg_sq"].reshape(-1)), n
    # the optimiser keeps working on the new buffers
    new_grads.buffer.normal_()
    before = new_params.buffer.clone()
    ours.step()
    assert ours.params is new_params and not torch.equal(before, new_params.buffer)


def test_select_rows_basics():
    H.pkg()
    optim = importlib.import_module(H.PKG_NAME + ".optim")
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    P = 1000
    f = mv.FlatGradients(P, "cuda")
    f.buffer.copy_(torch.arange(f.buffer.numel(), dtype=torch.float32))
    idx = torch.tensor([5, -1, 999, 0, 5], device="cuda")
    out = optim.select_rows(f, idx)
    for name, view in out.views.items():
        src = f.views[name]
        assert torch.equal(view[0], src[5]) and torch.equal(view[2], src[999]) and torch.equal(view[3], src[0]) and torch.equal(view[4], src[5])
        assert float(view[1].abs().max()) == 0.0, name
    assert optim.select_rows(f, torch.zeros(0, dtype=torch.int64, device="cuda")).buffer.numel() == 0
