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
    """[61 * P] fp32 buffer: means3D | features_dc | features_rest | segments | opacities | scales | rotations (raw values)."""

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

    def offsets(self):
        off, out = 0, {}
        for name, shape in self.shapes.items():
            n = int(torch.Size(shape).numel())
            out[name] = (off, n)
            off += n
        return out


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
        self.step_count = 0

    def set_lr(self, group, lr):
        """update_learning_rate (scene/gaussian_model.py:184-190) sets the xyz group's rate every iteration."""
        self.lrs[GROUPS.get(group, group)] = float(lr)

    def step(self):
        L = _lib.lib()
        self.step_count += 1
        offs = self.params.offsets()
        names = [n for n in offs if n in self.lrs]
        arr = (_lib.GsrAdamGroup * len(names))(*[_lib.GsrAdamGroup(offs[n][0], offs[n][1], self.lrs[n]) for n in names])
        dev = self.params.buffer.device
        with torch.cuda.device(dev):
            rc = L.gsr_adam_step(self.params.buffer.data_ptr(), self.grads.buffer.data_ptr(), self.exp_avg.data_ptr(),
                                 self.exp_avg_sq.data_ptr(), arr, len(names), self.betas[0], self.betas[1], self.eps, self.step_count,
                                 torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "gsr_adam_step")
