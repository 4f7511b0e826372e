"""Flat raw-parameter storage + fused multi-tensor Adam.

The reference keeps seven nn.Parameters and steps them with torch.optim.Adam(l, lr=0.0, eps=1e-15) -- one parameter group each
with its own learning rate, the xyz rate rescheduled every iteration (scene/gaussian_model.py:166-190). Here the parameters
live in ONE flat fp32 buffer with the layout of multiview.FlatGradients(split_sh=True), so that

    rasterizer backward / multi-GPU gather  ->  flat gradient buffer  ->  gsr_adam_step  ->  flat parameter buffer

never touches per-tensor torch kernels. `views[name]` are ordinary tensors shaped like the reference's parameters and can be
handed to forward_raw / _forward_native(raw_params=True) directly. No CPU path: the CUDA library is required."""
import ctypes

import torch

from . import _lib
from .multiview import FlatGradients

# reference parameter-group name -> flat-buffer block (scene/gaussian_model.py:166-174)
GROUPS = {"xyz": "means3D", "f_dc": "features_dc", "f_rest": "features_rest", "opacity": "opacities", "segment": "segments",
          "scaling": "scales", "rotation": "rotations"}


class FlatParameters(FlatGradients):
    """Flat fp32 buffer: means3D | features_dc | features_rest | segments | opacities | scales | rotations (raw values, 61 floats
    per Gaussian, blocks 256-byte aligned like FlatGradients)."""

    def __init__(self, P, device, sh_coeffs=16, num_class=2):
        super().__init__(P, device, sh_coeffs=sh_coeffs, num_class=num_class, split_sh=True)

    @classmethod
    def from_tensors(cls, tensors):
        """tensors: dict block name -> initial value (e.g. {"means3D": pc._xyz, "features_dc": pc._features_dc, ...})."""
        P = tensors["means3D"].size(0)
        fp = cls(P, tensors["means3D"].device, sh_coeffs=1 + tensors["features_rest"].size(1), num_class=tensors["segments"].size(1))
        with torch.no_grad():
            for name, view in fp.views.items():
                view.copy_(tensors[name].reshape(view.shape))
        return fp


class FusedAdam:
    """torch.optim.Adam(param_groups, eps=1e-15) semantics (no weight decay, no amsgrad) over a FlatParameters /
    FlatGradients pair, one kernel launch per step. `lrs`: dict reference group name ("xyz", "f_dc", ...) -> learning rate."""

    def __init__(self, params, grads, lrs, betas=(0.9, 0.999), eps=1e-15):
        if params.buffer.numel() != grads.buffer.numel() or list(params.shapes) != list(grads.shapes):
            raise ValueError("parameter and gradient buffers must share one layout")
        if not params.buffer.is_cuda:
            raise RuntimeError("FusedAdam needs CUDA buffers; libgsr has no CPU path")
        self.params, self.grads = params, grads
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.lrs = {GROUPS.get(k, k): float(v) for k, v in lrs.items()}
        unknown = set(self.lrs) - set(params.shapes)
        if unknown:
            raise KeyError("unknown parameter group(s): %s" % sorted(unknown))
        self.exp_avg = torch.zeros_like(params.buffer)
        self.exp_avg_sq = torch.zeros_like(params.buffer)
        self.steps = {n: 0 for n in params.shapes}  # torch.optim.Adam keeps `step` per parameter: a skipped group lags behind

    def set_lr(self, group, lr):
        """update_learning_rate (scene/gaussian_model.py:184-190) sets the xyz group's rate every iteration."""
        self.lrs[GROUPS.get(group, group)] = float(lr)

    def step(self, skip=()):
        """One Adam step. `skip`: parameter groups left untouched this step (torch.optim.Adam skips a parameter whose .grad is
        None, which is what happens to a tensor the reference has just replaced -- reset_opacity, densification)."""
        L = _lib.lib()
        skip = {GROUPS.get(k, k) for k in skip}
        offs = self.params.offsets()
        names = [n for n in offs if n in self.lrs and n not in skip]
        if not names:
            return
        names = [n for n in names if offs[n][1] > 0]
        for n in names:
            self.steps[n] += 1
        arr = (_lib.GsrAdamGroup * len(names))(*[_lib.GsrAdamGroup(offs[n][0], offs[n][1], self.lrs[n], self.steps[n]) for n in names])
        dev = self.params.buffer.device
        with torch.cuda.device(dev):
            rc = L.gsr_adam_step(self.params.buffer.data_ptr(), self.grads.buffer.data_ptr(), self.exp_avg.data_ptr(),
                                 self.exp_avg_sq.data_ptr(), arr, len(names), self.betas[0], self.betas[1], self.eps, max(self.steps.values()),
                                 torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "gsr_adam_step")


# ---------------------------------------------------------------------------------------------------------------------
# Densify / prune as ONE index list
# ---------------------------------------------------------------------------------------------------------------------
def select_rows(flat, index):
    """New flat buffer (same class and block layout as `flat`) whose row j is row index[j] of `flat` in every block;
    index[j] == -1 gives a row of zeros. index: int64 CUDA tensor. One launch (gsr_select_rows)."""
    if not flat.buffer.is_cuda:
        raise RuntimeError("select_rows needs CUDA buffers; libgsr has no CPU path")
    L = _lib.lib()
    dev = flat.buffer.device
    n_out, n_src = int(index.numel()), int(flat.shapes["means3D"][0])
    idx = index.to(device=dev, dtype=torch.int64).contiguous()
    sh_coeffs = (flat.shapes["features_rest"][1] + 1) if flat.split_sh else flat.shapes["shs"][1]
    num_class = flat.shapes["segments"][1]
    if isinstance(flat, FlatParameters):
        out = FlatParameters(n_out, dev, sh_coeffs=sh_coeffs, num_class=num_class)
    else:
        out = FlatGradients(n_out, dev, sh_coeffs=sh_coeffs, num_class=num_class, split_sh=flat.split_sh)
    rows = [int(torch.Size(s[1:]).numel()) for s in flat.shapes.values()]
    nb = len(rows)
    arr = (ctypes.c_int32 * nb)(*rows)
    so = (ctypes.c_uint64 * nb)(*[flat.offsets()[k][0] for k in flat.shapes])
    do = (ctypes.c_uint64 * nb)(*[out.offsets()[k][0] for k in flat.shapes])
    with torch.cuda.device(dev):
        rc = L.gsr_select_rows(flat.buffer.data_ptr(), out.buffer.data_ptr(), idx.data_ptr(), n_out, n_src, arr, so, do, nb,
                               torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "gsr_select_rows")
    return out


def build_rotation(r):
    """utils/general_utils.py:86-107 (normalised quaternion -> rotation matrix), same expressions."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def densify_and_prune(params, opt, xyz_gradient_accum, denom, max_grad, min_opacity, extent, max_screen_size, percent_dense=0.01, N=2,
                      generator=None):
    """GaussianModel.densify_and_prune (scene/gaussian_model.py:500-515: clone, split, prune) on the flat buffers.

    The masks and the handful of new rows (sampled positions, shrunk scales of split Gaussians) are computed with the same torch
    expressions as the reference on the small selected subsets -- including the one torch.normal call, so the same RNG state
    gives the same samples -- but the seven parameter tensors and their two Adam moments are rebuilt by ONE row selection each
    (gsr_select_rows) from a composed index list instead of 2 torch.cat + 2 boolean masks per tensor and state.
    `generator`: a dedicated torch.Generator for the split samples (callers that must not disturb the global RNG stream).
    Returns (new FlatParameters, new FlatGradients, index) and re-targets `opt` (moments reindexed, fresh rows zero); the caller
    resets its densification statistics to zeros of the new size, as densification_postfix does (:443-463). Like the reference,
    the screen-size criterion sees max_radii2D AFTER that reset (all zeros), so only the world-size criterion can fire."""
    v = params.views
    dev = params.buffer.device
    P = v["means3D"].size(0)
    with torch.no_grad():
        grads = xyz_gradient_accum / denom
        grads[grads.isnan()] = 0.0
        scaling = torch.exp(v["scales"])
        max_scale = torch.max(scaling, dim=1).values
        # ---- densify_and_clone (:483-498)
        sel_clone = torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False)
        sel_clone = torch.logical_and(sel_clone, max_scale <= percent_dense * extent)
        idx_clone = torch.nonzero(sel_clone).flatten()
        src1 = torch.cat((torch.arange(P, device=dev), idx_clone))           # source row of every row after the clone
        fresh1 = torch.cat((torch.zeros(P, dtype=torch.bool, device=dev), torch.ones(idx_clone.numel(), dtype=torch.bool, device=dev)))
        # ---- densify_and_split (:465-481): the statistics cover the first P rows only, clones see a padded gradient of 0
        padded_grad = torch.zeros(src1.numel(), device=dev)
        padded_grad[:grads.shape[0]] = grads.squeeze()
        sel_split = torch.where(padded_grad >= max_grad, True, False)
        sel_split = torch.logical_and(sel_split, max_scale[src1] > percent_dense * extent)
        src_split = src1[sel_split]
        stds = scaling[src_split].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device=dev)
        samples = torch.normal(mean=means, std=stds, generator=generator)  # generator=None: the global RNG, like the reference
        rots = build_rotation(v["rotations"][src_split]).repeat(N, 1, 1)
        child_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + v["means3D"][src_split].repeat(N, 1)
        child_scale = torch.log(scaling[src_split].repeat(N, 1) / (0.8 * N))
        n_child = child_xyz.size(0)
        src2 = torch.cat((src1, src_split.repeat(N)))
        fresh2 = torch.cat((fresh1, torch.ones(n_child, dtype=torch.bool, device=dev)))
        child_of = torch.cat((torch.full((src1.numel(),), -1, dtype=torch.int64, device=dev), torch.arange(n_child, device=dev)))
        keep = ~torch.cat((sel_split, torch.zeros(n_child, device=dev, dtype=torch.bool)))  # the split originals go
        src3, fresh3, child3 = src2[keep], fresh2[keep], child_of[keep]
        # ---- prune (:507-513)
        is_child = child3 >= 0
        scale3 = torch.where(is_child[:, None], child_scale[child3.clamp(min=0)], v["scales"][src3])
        prune_mask = (torch.sigmoid(v["opacities"][src3]) < min_opacity).squeeze(-1)
        if max_screen_size:
            big_points_vs = torch.zeros_like(prune_mask)  # max_radii2D was reset to zeros by the densification above
            big_points_ws = torch.exp(scale3).max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
        keep2 = ~prune_mask
        index, fresh, child = src3[keep2], fresh3[keep2], child3[keep2]
        # ---- one row selection per buffer
        new_params = select_rows(params, index)
        ch = torch.nonzero(child >= 0).flatten()
        new_params.views["means3D"][ch] = child_xyz[child[ch]]
        new_params.views["scales"][ch] = child_scale[child[ch]]
        new_grads = FlatGradients(index.numel(), dev, sh_coeffs=1 + params.shapes["features_rest"][1], num_class=params.shapes["segments"][1],
                                  split_sh=True)
        moment_index = torch.where(fresh, torch.full_like(index, -1), index)
        holder = FlatGradients.__new__(FlatGradients)
        holder.shapes, holder.split_sh, holder._offsets = params.shapes, True, params.offsets()
        for name in ("exp_avg", "exp_avg_sq"):
            holder.buffer = getattr(opt, name)
            setattr(opt, name, select_rows(holder, moment_index).buffer)
        opt.params, opt.grads = new_params, new_grads
    return new_params, new_grads, index
