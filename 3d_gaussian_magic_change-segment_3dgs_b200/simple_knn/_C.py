"""Drop-in `simple_knn._C` over libgsr: distCUDA2(points f32[P,3] cuda) -> f32[P], the mean squared distance to
the 3 nearest neighbours (reference: simple-knn/spatial.cu:15-26, simple_knn.cu:185-221; caller
scene/gaussian_model.py:143). No CPU path."""
import importlib.util
import os
import sys

import torch


def _load_lib_module():
    name = "_gsr_b200_lib"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_lib.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_lib = _load_lib_module()


def distCUDA2(points):
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor; libgsr has no CPU path")
    L = _lib.lib()
    device = points.device
    P = points.size(0)
    with torch.cuda.device(device):
        means = torch.zeros(P, dtype=torch.float32, device=device)
        if P == 0:
            return means
        pts = points.contiguous().float()
        nws = L.gsr_knn_workspace_bytes(P)
        ws = torch.empty(nws, dtype=torch.uint8, device=device)
        rc = L.gsr_knn_dist2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(), nws, torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "gsr_knn_dist2")
        return means
