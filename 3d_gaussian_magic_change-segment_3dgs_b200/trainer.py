"""Batched multi-view training driver on the native path: the optimisation loop of train.py:85-186 taking B
cameras per step, built from this package's pieces only --

    flat raw parameters (optim.FlatParameters)
      -> fused-activation forward (gsr_forward, raw_params)                          render()            train.py:106
      -> L1 + SSIM loss and its image gradient (gsr_image_loss)                      train.py:110-111
      -> depth supervision from a monocular prior and its depth-image gradient        train.py:115-121 (using_depth, 'localrf'),
         (gsr_depth_loss: median / quantile by radix select, fused normalisation)     utils/loss_utils.py:88-102
      -> backward into the flat gradient buffer / as packets (gsr_backward[_packets]) loss.backward()     train.py:143
      -> multi-GPU sum over all ranks' views (peer-memory gather, NCCL fallback)     (new: the reference is single-GPU)
      -> densification statistics per view, densify / prune / opacity reset          train.py:169-180
      -> fused Adam with the exponential xyz schedule (gsr_adam_step)                 train.py:183-185, gaussian_model.py:184-190

The step's B views are split views[rank::world]; every view's loss is scaled by 1/B, so the summed gradient is that of the mean
loss over the batch (B = 1, world = 1 is the reference's step). Parameters are replicated and every rank applies the same
update, so replicas stay bitwise identical. No CPU path: the CUDA library is required."""
from typing import NamedTuple

import numpy as np
import torch

from . import losses, multiview as mv, optim


class OptimizationParams(NamedTuple):
    """arguments/__init__.py:90-112 (defaults of the reference)."""
    iterations: int = 30_000
    position_lr_init: float = 0.00008
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 30_000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    segment_lr: float = 0.05
    scaling_lr: float = 0.002
    rotation_lr: float = 0.001
    percent_dense: float = 0.01
    lambda_dssim: float = 0.2
    lambda_depth: float = 0.1  # weight of the depth-supervision term (using_depth, depth_loss_choice 'localrf')
    densification_interval: int = 100
    opacity_reset_interval: int = 3000
    densify_from_iter: int = 500
    densify_until_iter: int = 15_000
    densify_grad_threshold: float = 0.0002


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """utils/general_utils.py:37-70: log-linear interpolation from lr_init to lr_final with an optional eased-in delay."""

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return delay_rate * log_lerp

    return helper


class NativeTrainer:
    def __init__(self, D, init, opt_params, cameras_extent, bg, spatial_lr_scale=1.0, max_sh_degree=3, dist=None, rank=0, world=1,
                 white_background=False, grad_exchange="peer"):
        """D: the drop-in diff_gaussian_rasterization module; init: dict of RAW parameter tensors (means3D, features_dc,
        features_rest, segments, opacities, scales, rotations) on this rank's CUDA device, identical on every rank."""
        self.D, self.o, self.extent, self.bg = D, opt_params, float(cameras_extent), bg
        self.dist, self.rank, self.world = dist if world > 1 else None, rank, world
        self.device = init["means3D"].device
        self.params = optim.FlatParameters.from_tensors(init)
        self.grads = mv.FlatGradients(self.P, self.device, sh_coeffs=1 + init["features_rest"].size(1), num_class=init["segments"].size(1),
                                      split_sh=True)
        self._rng = torch.Generator(device=self.device)  # split samples of the densification: never touches the global RNG
        o = opt_params
        self.opt = optim.FusedAdam(self.params, self.grads, {"xyz": o.position_lr_init * spatial_lr_scale, "f_dc": o.feature_lr,
                                                             "f_rest": o.feature_lr / 20.0, "opacity": o.opacity_lr, "segment": o.segment_lr,
                                                             "scaling": o.scaling_lr, "rotation": o.rotation_lr})
        self.xyz_lr = get_expon_lr_func(o.position_lr_init * spatial_lr_scale, o.position_lr_final * spatial_lr_scale,
                                        lr_delay_mult=o.position_lr_delay_mult, max_steps=o.position_lr_max_steps)
        self.max_sh_degree, self.active_sh_degree = max_sh_degree, 0
        self.white_background = white_background
        self.iteration = 0
        self.stats = mv.DensificationStats(self.P, self.device)
        self.grad_exchange = grad_exchange
        self.px = None
        self._m2 = None
        self._make_exchange()

    P = property(lambda self: self.params.views["means3D"].size(0))

    def _make_exchange(self):
        """(Re)create the per-size state: peer-visible packet buffers (collective) and the screen-space gradient scratch."""
        if self.px is not None:
            self.px.close()
            self.px = None
        self._m2 = torch.zeros(self.P, 3, device=self.device)
        self._xstate = {}
        if self.dist is not None and self.grad_exchange == "peer":
            try:
                self.px = mv.PeerPacketExchange(self.D, self.dist, self.P, self._views_per_rank_hint(), self.rank, self.world, self.device)
            except mv.PeerUnavailable:
                self.px = None  # every rank falls back to the NCCL packet exchange together

    def _views_per_rank_hint(self):
        return getattr(self, "_vpr", 1)

    def _settings(self, cam):
        dev = self.device
        t = lambda k: cam[k].to(dev) if isinstance(cam[k], torch.Tensor) else torch.as_tensor(cam[k], device=dev)
        return self.D.GaussianRasterizationSettings(image_height=int(cam["image_height"]), image_width=int(cam["image_width"]),
                                                    tanfovx=float(cam["tanfovx"]), tanfovy=float(cam["tanfovy"]), bg=self.bg.to(dev),
                                                    scale_modifier=1.0, viewmatrix=t("viewmatrix"), projmatrix=t("projmatrix"),
                                                    sh_degree=self.active_sh_degree, campos=t("campos"), prefiltered=False, debug=False)

    def train_step(self, cams, gt_images, gt_depths=None):
        """One optimisation step over the batch `cams` (list of camera dicts: image_width/height, tanfovx/y, viewmatrix,
        projmatrix, campos -- the same list on every rank) with ground-truth images gt_images[i] ([3,H,W] CUDA, needed for this
        rank's views only) and, optionally, monocular depth priors gt_depths[i] ([1,H,W] or None): the reference's using_depth
        step, loss += compute_depth_loss(1 / render()["depth"].clamp(1e-6), gt_depth, lambda_depth) (train.py:115-121).
        Returns the batch-mean loss as a device scalar (all ranks' views, all-reduced)."""
        o, D, dist = self.o, self.D, self.dist
        self.iteration += 1
        it = self.iteration
        self.opt.set_lr("xyz", self.xyz_lr(it))  # update_learning_rate
        if it % 1000 == 0 and self.active_sh_degree < self.max_sh_degree:  # oneupSHdegree
            self.active_sh_degree += 1
        B = len(cams)
        if B % self.world:
            raise ValueError("the batch (%d views) must divide evenly over %d ranks" % (B, self.world))
        mine = mv.shard_views(B, self.rank, self.world)
        if dist is not None and self.px is not None and len(mine) != self.px.nv:
            self._vpr = len(mine)
            self._make_exchange()
        v = self.params.views
        total = torch.zeros((), device=self.device)
        sets, campos_mine = [], []
        track = it < o.densify_until_iter
        with torch.no_grad():
            for k, vi in enumerate(mine):
                rs = self._settings(cams[vi])
                fwd = mv.native_view_forward(D, v, rs)
                stats, g_color = losses.l1_ssim_loss_and_grad(fwd[1], gt_images[vi], o.lambda_dssim, grad_scale=1.0 / B)
                total += stats[2] / B
                up = {"color": g_color}
                if gt_depths is not None and gt_depths[vi] is not None:  # viewpoint_cam.depth is not None and dataset.using_depth
                    d_loss, g_depth = losses.depth_loss_and_grad(fwd[2], gt_depths[vi], o.lambda_depth, grad_scale=1.0 / B,
                                                                 fused_from_raw_depth=True)
                    total += d_loss[0] / B
                    up["depth"] = g_depth
                if dist is None:
                    mv.native_view_backward(D, v, rs, fwd, up, self.grads, first=(k == 0), means2D_grad=self._m2 if track else None)
                elif self.px is not None:
                    self.px.view_backward(v, rs, fwd, up, k, means2D_grad=self._m2 if track else None)
                else:
                    sets.append(mv.native_view_backward_packets(D, v, rs, fwd, up, means2D_grad=self._m2 if track else None,
                                                                capacity=self._xstate.get("cap", 0)))
                if track:  # train.py:170-172: statistics of the UNscaled per-view loss
                    self.stats.add_view(self._m2 * B, fwd[5])
            if dist is not None:
                all_campos = [[torch.as_tensor(cams[i]["campos"]).to(self.device) for i in mv.shard_views(B, r, self.world)]
                              for r in range(self.world)]
                if self.px is not None:
                    self.px.exchange(self.grads, v, all_campos, self.active_sh_degree)
                else:
                    mv.exchange_packets(D, dist, self.grads, v, sets, all_campos, self.active_sh_degree, self.world, state=self._xstate)
                dist.all_reduce(total)
            elif not mine:
                self.grads.buffer.zero_()
            skip = ()
            if track:
                if it > o.densify_from_iter and it % o.densification_interval == 0:
                    self.densify(20 if it > o.opacity_reset_interval else None)
                    skip = tuple(optim.GROUPS)  # every tensor was rebuilt: the reference's optimizer.step() finds .grad None everywhere
                if it % o.opacity_reset_interval == 0 or (self.white_background and it == o.densify_from_iter):
                    self.reset_opacity()
                    skip = skip or ("opacity",)
            if it < o.iterations:
                self.opt.step(skip=skip)
        return total

    def densify(self, size_threshold):
        """densify_and_prune with the statistics of ALL ranks; every rank draws the same samples (seeded by the iteration)."""
        if self.dist is not None:
            self.stats.allreduce(self.dist)
        self._rng.manual_seed(0x3D65 + self.iteration)  # the same seed on every rank: replicas draw identical samples
        self.params, self.grads, _ = optim.densify_and_prune(self.params, self.opt, self.stats.xyz_gradient_accum, self.stats.denom,
                                                             self.o.densify_grad_threshold, 0.005, self.extent, size_threshold,
                                                             percent_dense=self.o.percent_dense, generator=self._rng)
        # the step's gradient was computed for the old rows; train_step skips the optimiser step of this iteration, like the
        # reference, whose rebuilt nn.Parameters carry no .grad when optimizer.step() runs (train.py:183)
        self.stats = mv.DensificationStats(self.P, self.device)
        self._make_exchange()

    def reset_opacity(self):
        """gaussian_model.py:256-260: opacity <- min(opacity, 0.01) in logit space, Adam moments of the group zeroed."""
        op = self.params.views["opacities"]
        x = torch.min(torch.sigmoid(op), torch.ones_like(op) * 0.01)
        op.copy_(torch.log(x / (1 - x)))
        off, cnt = self.params.offsets()["opacities"]
        self.opt.exp_avg[off:off + cnt].zero_()
        self.opt.exp_avg_sq[off:off + cnt].zero_()

    def close(self):
        if self.px is not None:
            self.px.close()
            self.px = None
