"""Drop-in `diff_gaussian_rasterization` over libgsr (B200 / sm_100a).

Mirrors the public surface of the reference module
(submodules_local/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:21-235):

    GaussianRasterizationSettings   NamedTuple, same field order            (:168-180)
    GaussianRasterizer              nn.Module: markVisible(), forward()     (:182-235)
    rasterize_gaussians(...)        -> (color, radii, depth, alpha, segment) (:21-44, :102)

so `gaussian_renderer.render()`, `train.py` and `train_segment.py` run unchanged with this directory's parent on
`sys.path`. The native side is the C-ABI library libgsr.so (include/gsr.h), called through ctypes with raw
device pointers on `torch.cuda.current_stream()`; torch only owns memory and streams. There is no CPU path:
CPU tensors raise.
"""
import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn



def _load_lib_module():
    """Load ../_lib.py (the ctypes binding of libgsr.so) by path, once, whichever way this module was imported:
    as a sub-package of the b200 package or as top-level `diff_gaussian_rasterization` from a sys.path entry."""
    import importlib.util
    import os
    import sys

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
ALLOC_FN, GsrGaussians, GsrOutputs, GsrParamGrads = _lib.ALLOC_FN, _lib.GsrGaussians, _lib.GsrOutputs, _lib.GsrParamGrads
GsrPixelGrads, GsrState, GsrStateExport, GsrView = _lib.GsrPixelGrads, _lib.GsrState, _lib.GsrStateExport, _lib.GsrView

PACKET_WORDS = _lib.GSR_PACKET_WORDS
NUM_CHANNELS = 3  # cuda_rasterizer/config.h:15
NUM_CLASS = 2     # cuda_rasterizer/config.h:16 (segment channels rendered when `segments` is absent; with `segments` given the
                  # class count is its second dimension, any value up to 64 -- a compile-time constant in the reference)


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def _ptr(t):
    """Device pointer of a tensor; empty tensors are the reference's "not provided" sentinel (-> NULL)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t, device, what):
    """contiguous fp32 on the compute device (the reference calls .contiguous().data<float>()), starting on a 16-byte boundary:
    the kernels load quaternions as float4 and segments as float2, so a contiguous VIEW at an odd offset of a larger buffer
    (which the reference's scalar loads accept) is copied once instead of faulting."""
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        raise RuntimeError("%s must be on %s (got %s); libgsr has no CPU path" % (what, device, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (got %s)" % (what, t.dtype))
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


class NumRendered(int):
    """The forward's num_rendered (a plain int for every consumer, as in the reference) that also remembers how many Gaussians
    THAT forward found visible -- what sizes the packet buffer of that view's backward, whatever ran on the thread in between."""
    num_visible = 0

    def __new__(cls, rendered, visible):
        obj = super().__new__(cls, rendered)
        obj.num_visible = int(visible)
        return obj


class _Alloc:
    """gsr_alloc_fn: hands libgsr torch-owned byte buffers (the reference's resizeFunctional, rasterize_points.cu:27-33)."""

    def __init__(self, device):
        self.device = device
        self.bufs = [None, None, None]
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, user, which, nbytes):
        try:
            t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
            self.bufs[which] = t
            return t.data_ptr()
        except Exception:  # report as allocation failure through the C ABI
            return 0

    def release(self):
        """Break the self -> CFUNCTYPE -> bound method -> self reference cycle and drop the buffer references."""
        self.cb = None
        self.bufs = None


def _view_struct(rs, M, num_class, keep):
    bg = _prep(rs.bg, keep["device"], "bg")
    vm = _prep(rs.viewmatrix, keep["device"], "viewmatrix")
    pm = _prep(rs.projmatrix, keep["device"], "projmatrix")
    cp = _prep(rs.campos, keep["device"], "campos")
    keep["view_tensors"] = (bg, vm, pm, cp)
    return GsrView(int(rs.image_width), int(rs.image_height), float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier),
                   int(rs.sh_degree), int(M), int(num_class), int(bool(rs.prefiltered)), int(bool(rs.debug)),
                   _ptr(bg), _ptr(vm), _ptr(pm), _ptr(cp))


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                        subset=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings, subset)


def _subset(subset, device):
    """int32 contiguous index list on the compute device, or None."""
    if subset is None:
        return None
    if subset.device != device:
        raise RuntimeError("subset must be on %s; libgsr has no CPU path" % (device,))
    return subset.to(torch.int32).contiguous()


def _forward_native(means3D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp, rs, sh_rest=None, raw_params=False,
                    subset=None):
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-125) over gsr_forward. Returns the reference's 9-tuple
    (num_rendered, color, depth, segment, alpha, radii, geomBuffer, binningBuffer, imgBuffer).

    raw_params=True (fused activations, GsrGaussians.raw_params): the tensors are the model's RAW parameters -- opacity /
    segment logits, log-scales, un-normalised quaternions, sh = _features_dc [P,1,3] and sh_rest = _features_rest [P,M-1,3].

    subset (int32 CUDA tensor, strictly ascending row numbers): render only those Gaussians without materialising masked copies
    (GsrGaussians.subset); radii then has one entry per list element."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor; libgsr has no CPU path")
    L = _lib.lib()
    device = means3D.device
    P, H, W = means3D.size(0), int(rs.image_height), int(rs.image_width)
    num_class = segments.size(1) if (segments is not None and segments.numel() > 0) else NUM_CLASS
    with torch.cuda.device(device):
        opts = dict(dtype=torch.float32, device=device)
        if P == 0:  # the core is skipped; images stay zero (background NOT applied), rasterize_points.cu:87
            z = lambda c: torch.zeros((c, H, W), **opts)
            e = lambda: torch.empty(0, dtype=torch.uint8, device=device)
            return 0, z(NUM_CHANNELS), z(1), z(num_class), z(1), torch.zeros(0, dtype=torch.int32, device=device), e(), e(), e()
        keep = {"device": device}
        M = sh.size(1) if (sh is not None and sh.numel() > 0) else 0
        t_rest = _prep(sh_rest, device, "sh_rest") if raw_params else None
        if raw_params and M > 0:
            if M != 1:
                raise RuntimeError("raw_params: sh must be _features_dc with shape (P, 1, 3)")
            M += t_rest.size(1) if t_rest is not None else 0
        view = _view_struct(rs, M, num_class, keep)
        t_means = _prep(means3D, device, "means3D")
        t_sh, t_col = _prep(sh, device, "sh"), _prep(colors_precomp, device, "colors_precomp")
        t_seg, t_op = _prep(segments, device, "segments"), _prep(opacities, device, "opacities")
        t_sc, t_rot = _prep(scales, device, "scales"), _prep(rotations, device, "rotations")
        t_cov = _prep(cov3Ds_precomp, device, "cov3Ds_precomp")
        t_sub = _subset(subset, device)
        count = P if t_sub is None else int(t_sub.numel())
        gin = GsrGaussians(P, _ptr(t_means), _ptr(t_sh), _ptr(t_col), _ptr(t_seg), _ptr(t_op), _ptr(t_sc), _ptr(t_rot), _ptr(t_cov),
                           _ptr(t_rest), int(bool(raw_params)), _ptr(t_sub), count if t_sub is not None else 0)
        color = torch.empty((NUM_CHANNELS, H, W), **opts)
        segment = torch.empty((num_class, H, W), **opts)
        depth = torch.empty((1, H, W), **opts)
        alpha = torch.empty((1, H, W), **opts)
        radii = torch.empty(count, dtype=torch.int32, device=device)
        if count == 0:  # an empty index list: like P == 0, the core is skipped and the images stay zero
            eb = lambda: torch.empty(0, dtype=torch.uint8, device=device)
            return 0, color.zero_(), depth.zero_(), segment.zero_(), alpha.zero_(), radii, eb(), eb(), eb()
        out = GsrOutputs(color.data_ptr(), segment.data_ptr(), depth.data_ptr(), alpha.data_ptr(), radii.data_ptr())
        alloc = _Alloc(device)
        R = ctypes.c_int32(0)
        stream = torch.cuda.current_stream(device).cuda_stream
        try:
            rc = L.gsr_forward(ctypes.byref(view), ctypes.byref(gin), ctypes.byref(out), alloc.cb, None, ctypes.byref(R), stream)
            bufs = alloc.bufs
        finally:
            alloc.release()  # the ctypes callback <-> bound method cycle would otherwise pin ~1 GB of state until a GC pass
        _lib.check(rc, "gsr_forward")
        e = lambda t: t if t is not None else torch.empty(0, dtype=torch.uint8, device=device)
        return NumRendered(R.value, L.gsr_last_num_visible()), color, depth, segment, alpha, radii, e(bufs[0]), e(bufs[1]), e(bufs[2])


def _forward_parts_native(parts, rs):
    """Sub-scene fusion without concatenation (GsrGaussians.parts; the viewer's _merge_scenes, visualizer.py:196-226, concatenates
    every attribute array of the sub-scenes instead). `parts`: list of dicts with the classic inputs of one sub-scene each --
    means3D, opacities, and shs | colors_precomp, scales + rotations | cov3D_precomp, optionally segments -- all providing the same
    members. Renders them as ONE scene (part 0's Gaussians first) and returns the 9-tuple of _forward_native; radii covers the
    fused scene. Render-only: there is no backward for this entry."""
    if not parts or len(parts) > _lib.GSR_MAX_PARTS:
        raise RuntimeError("forward_parts needs 1..%d sub-scenes" % _lib.GSR_MAX_PARTS)
    L = _lib.lib()
    device = parts[0]["means3D"].device
    if not parts[0]["means3D"].is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor; libgsr has no CPU path")
    H, W = int(rs.image_height), int(rs.image_width)
    keys = ["means3D", "shs", "colors_precomp", "segments", "opacities", "scales", "rotations", "cov3D_precomp"]
    have0 = [k for k in keys if parts[0].get(k) is not None and parts[0][k].numel() > 0]
    with torch.cuda.device(device):
        arr = (GsrGaussians * len(parts))()
        keep, total = [], 0
        for i, part in enumerate(parts):
            if [k for k in keys if part.get(k) is not None and part[k].numel() > 0] != have0:
                raise RuntimeError("sub-scene %d provides a different set of inputs than sub-scene 0" % i)
            if part["means3D"].dim() != 2 or part["means3D"].size(1) != 3 or part["means3D"].size(0) == 0:
                raise RuntimeError("means3D must have dimensions (num_points > 0, 3)")
            t = {k: _prep(part.get(k), device, k) for k in keys}
            keep.append(t)
            n = t["means3D"].size(0)
            arr[i] = GsrGaussians(n, _ptr(t["means3D"]), _ptr(t["shs"]), _ptr(t["colors_precomp"]), _ptr(t["segments"]), _ptr(t["opacities"]),
                                  _ptr(t["scales"]), _ptr(t["rotations"]), _ptr(t["cov3D_precomp"]), None, 0, None, 0, None, 0)
            total += n
        first = keep[0]
        M = first["shs"].size(1) if first["shs"] is not None else 0
        num_class = first["segments"].size(1) if first["segments"] is not None else NUM_CLASS
        k2 = {"device": device}
        view = _view_struct(rs, M, num_class, k2)
        gin = GsrGaussians(total, None, None, None, None, None, None, None, None, None, 0, None, 0, arr, len(parts))
        opts = dict(dtype=torch.float32, device=device)
        color, segment = torch.empty((NUM_CHANNELS, H, W), **opts), torch.empty((num_class, H, W), **opts)
        depth, alpha = torch.empty((1, H, W), **opts), torch.empty((1, H, W), **opts)
        radii = torch.empty(total, dtype=torch.int32, device=device)
        out = GsrOutputs(color.data_ptr(), segment.data_ptr(), depth.data_ptr(), alpha.data_ptr(), radii.data_ptr())
        alloc = _Alloc(device)
        R = ctypes.c_int32(0)
        try:
            rc = L.gsr_forward(ctypes.byref(view), ctypes.byref(gin), ctypes.byref(out), alloc.cb, None, ctypes.byref(R),
                               torch.cuda.current_stream(device).cuda_stream)
            bufs = alloc.bufs
        finally:
            alloc.release()
        _lib.check(rc, "gsr_forward")
        e = lambda t: t if t is not None else torch.empty(0, dtype=torch.uint8, device=device)
        return NumRendered(R.value, L.gsr_last_num_visible()), color, depth, segment, alpha, radii, e(bufs[0]), e(bufs[1]), e(bufs[2])


def count_work(P, W, H, geomBuffer, binningBuffer, imgBuffer, num_rendered):
    """gsr_count_work: algorithmic work of a rendered frame counted from its saved state (measurement support):
    {"E": entries evaluated front to back, "Cc": contributing entries, "E_b": entries the backward re-traverses}."""
    L = _lib.lib()
    device = geomBuffer.device
    with torch.cuda.device(device):
        cnt = torch.zeros(3, dtype=torch.int64, device=device)
        st = GsrState(_ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), int(num_rendered))
        rc = L.gsr_count_work(int(P), int(W), int(H), ctypes.byref(st), cnt.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "gsr_count_work")
        E, Cc, Eb = (int(v) for v in cnt.tolist())
    return {"E": E, "Cc": Cc, "E_b": Eb}


def microbench(device=None):
    """gsr_microbench on the current (or given) device: achievable FFMA / FFMA2 / MUFU.EX2 / SHFL / red.global rates (dict)."""
    L = _lib.lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    with torch.cuda.device(device):
        r = _lib.GsrMicrobench()
        _lib.check(L.gsr_microbench(ctypes.byref(r), torch.cuda.current_stream(device).cuda_stream), "gsr_microbench")
    return {"ffma_tflops": round(r.ffma_tflops, 3), "ffma2_tflops": round(r.ffma2_tflops, 3), "ex2_gops": round(r.ex2_gops, 2),
            "shfl_gops": round(r.shfl_gops, 2), "red_gops": round(r.red_gops, 3), "sm_count": int(r.sm_count),
            "sm_clock_mhz_nominal": round(r.sm_clock_mhz_nominal, 1),
            "what": "dependent-chain-free FFMA (3-register form), packed FFMA2, MUFU.EX2, SHFL (lane-ops/s) and 12-lane red.global.add.f32 "
                    "into 48-byte records of an L2-resident table (float atomics/s), CUDA events, 64 warps/SM"}


def _backward_native(rs, means3D, radii, colors_precomp, segments, scales, rotations, cov3Ds_precomp, grad_color, grad_segment, grad_depth,
                     grad_alpha, sh, geomBuffer, num_rendered, binningBuffer, imgBuffer, alpha, needs=None, out=None, accumulate=False,
                     sh_rest=None, raw_params=False, opacities=None, subset=None):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:127-221) over gsr_backward. Returns a dict of dense
    gradients (zeros for invisible Gaussians). `needs` optionally names the gradients to produce.

    raw_params=True: inputs are the raw parameters (see _forward_native; `opacities` = the logits is then required) and the
    gradients are w.r.t. them; "sh" is [P,1,3] (features_dc) and "sh_rest" [P,M-1,3] (features_rest).

    `out` (dict name -> preallocated contiguous fp32 tensor, e.g. views of one flat buffer) makes the kernels write there
    instead of fresh tensors; with `accumulate=True` the rows of visible Gaussians are ADDED to `out` and nothing else is
    touched (multi-view accumulation without a zero-fill or a torch add per view)."""
    L = _lib.lib()
    device = means3D.device
    P, H, W = means3D.size(0), int(rs.image_height), int(rs.image_width)
    M = sh.size(1) if (sh is not None and sh.numel() > 0) else 0
    Mrest = sh_rest.size(1) if (raw_params and M > 0 and sh_rest is not None and sh_rest.numel() > 0) else 0
    M += Mrest
    num_class = segments.size(1) if (segments is not None and segments.numel() > 0) else NUM_CLASS
    names = ["means3D", "means2D", "sh", "colors_precomp", "segments", "opacities", "scales", "rotations", "cov3Ds_precomp", "sh_rest"]
    want = {n: True for n in names} if needs is None else {n: bool(needs.get(n, n == "sh_rest" and needs.get("sh", False))) for n in names}
    have = {"sh": M > 0, "sh_rest": Mrest > 0, "colors_precomp": colors_precomp is not None and colors_precomp.numel() > 0,
            "segments": segments is not None and segments.numel() > 0,
            "scales": scales is not None and scales.numel() > 0, "rotations": rotations is not None and rotations.numel() > 0,
            "cov3Ds_precomp": cov3Ds_precomp is not None and cov3Ds_precomp.numel() > 0}
    shapes = {"means3D": (P, 3), "means2D": (P, 3), "sh": (P, 1, 3) if raw_params else (P, M, 3), "colors_precomp": (P, 3),
              "segments": (P, num_class), "opacities": (P, 1), "scales": (P, 3), "rotations": (P, 4), "cov3Ds_precomp": (P, 6),
              "sh_rest": (P, Mrest, 3)}
    with torch.cuda.device(device):
        opts = dict(dtype=torch.float32, device=device)
        grads = {}
        for n in names:
            if out is not None:
                t = out.get(n)
                if t is not None and (tuple(t.shape) != shapes[n] or t.dtype != torch.float32 or not t.is_contiguous() or t.device != device):
                    raise RuntimeError("out[%r] must be a contiguous float32 %s tensor on %s" % (n, shapes[n], device))
                grads[n] = t if have.get(n, True) else None
            elif want[n] and have.get(n, True):
                grads[n] = torch.zeros(shapes[n], **opts) if P == 0 else torch.empty(shapes[n], **opts)
            else:
                grads[n] = None
        if accumulate and out is None:
            raise RuntimeError("accumulate=True needs `out`")
        if P == 0:
            return grads
        keep = {"device": device}
        view = _view_struct(rs, M, num_class, keep)
        t_means = _prep(means3D, device, "means3D")
        t_sh, t_col = _prep(sh, device, "sh"), _prep(colors_precomp, device, "colors_precomp")
        t_seg = _prep(segments, device, "segments")
        t_sc, t_rot = _prep(scales, device, "scales"), _prep(rotations, device, "rotations")
        t_cov = _prep(cov3Ds_precomp, device, "cov3Ds_precomp")
        # classic: opacities are not needed, they live in the saved geometry state (as in the reference, conic_opacity.w);
        # raw_params: the sigmoid's backward needs the logits
        t_op = _prep(opacities, device, "opacities") if raw_params else None
        if raw_params and t_op is None:
            raise RuntimeError("raw_params backward needs the raw opacities")
        t_rest = _prep(sh_rest, device, "sh_rest") if raw_params else None
        t_sub = _subset(subset, device)
        count = P if t_sub is None else int(t_sub.numel())
        gin = GsrGaussians(P, _ptr(t_means), _ptr(t_sh), _ptr(t_col), _ptr(t_seg), t_op.data_ptr() if raw_params else t_means.data_ptr(),
                           _ptr(t_sc), _ptr(t_rot), _ptr(t_cov), _ptr(t_rest), int(bool(raw_params)), _ptr(t_sub),
                           count if t_sub is not None else 0)
        if count == 0:  # nothing was rendered: every gradient row is zero
            for t in grads.values():
                if t is not None and not accumulate:
                    t.zero_()
            return grads
        g_col = _prep(grad_color, device, "grad_color")
        if g_col is None:
            g_col = torch.zeros((NUM_CHANNELS, H, W), **opts)
        g_seg, g_dep, g_alp = _prep(grad_segment, device, "grad_segment"), _prep(grad_depth, device, "grad_depth"), _prep(grad_alpha, device, "grad_alpha")
        pix = GsrPixelGrads(g_col.data_ptr(), _ptr(g_seg), _ptr(g_dep), _ptr(g_alp))
        pg = GsrParamGrads(_ptr(grads["means3D"]), _ptr(grads["means2D"]), _ptr(grads["sh"]), _ptr(grads["colors_precomp"]),
                           _ptr(grads["segments"]), _ptr(grads["opacities"]), _ptr(grads["scales"]), _ptr(grads["rotations"]),
                           _ptr(grads["cov3Ds_precomp"]), int(bool(accumulate)), _ptr(grads["sh_rest"]))
        state = GsrState(_ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), int(num_rendered), int(getattr(num_rendered, "num_visible", 0)))
        nscratch = L.gsr_backward_scratch_bytes_n(count, int(num_class))
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=device)
        t_radii = radii.contiguous()
        t_alpha = _prep(alpha, device, "alpha")
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = L.gsr_backward(ctypes.byref(view), ctypes.byref(gin), t_radii.data_ptr(), ctypes.byref(state), t_alpha.data_ptr(),
                            ctypes.byref(pix), ctypes.byref(pg), scratch.data_ptr(), nscratch, stream)
        _lib.check(rc, "gsr_backward")
        return grads


def last_num_visible():
    """Visible Gaussians of the most recent forward on this thread (host value; the forward already synchronised for it)."""
    return int(_lib.lib().gsr_last_num_visible())


def _backward_packets_native(rs, means3D, radii, segments, scales, rotations, grad_color, grad_segment, grad_depth, grad_alpha, sh,
                             geomBuffer, num_rendered, binningBuffer, imgBuffer, alpha, capacity, means2D_grad=None, raw=None, raw_params=None):
    """gsr_backward_packets: the backward of one view as compact per-visible-Gaussian packets (16 words each, see
    include/gsr.h) instead of dense gradient rows. Returns (blob, count int32[1]): blob is ONE int32 tensor of
    packet_index_words(P) + capacity * 16 words -- the view's visibility index followed by the packets -- i.e. the all-gather
    payload of the view (see packet_blob_views). With raw=(packets_ptr, index_ptr) (device addresses, e.g. inside a
    gsr_peer_alloc buffer; room for `capacity` packets) the view is written there instead and blob is None.
    raw_params = {"sh_rest": _features_rest, "opacities": logits}: fused activations, packets carry raw-parameter gradients."""
    L = _lib.lib()
    device = means3D.device
    P, H, W = means3D.size(0), int(rs.image_height), int(rs.image_width)
    M = sh.size(1) + (raw_params["sh_rest"].size(1) if raw_params else 0)
    num_class = segments.size(1) if (segments is not None and segments.numel() > 0) else NUM_CLASS
    with torch.cuda.device(device):
        opts = dict(dtype=torch.float32, device=device)
        keep = {"device": device}
        view = _view_struct(rs, M, num_class, keep)
        t_means, t_sh, t_seg = _prep(means3D, device, "means3D"), _prep(sh, device, "sh"), _prep(segments, device, "segments")
        t_sc, t_rot = _prep(scales, device, "scales"), _prep(rotations, device, "rotations")
        t_rest = _prep(raw_params["sh_rest"], device, "sh_rest") if raw_params else None
        t_op = _prep(raw_params["opacities"], device, "opacities") if raw_params else t_means
        gin = GsrGaussians(P, _ptr(t_means), _ptr(t_sh), None, _ptr(t_seg), t_op.data_ptr(), _ptr(t_sc), _ptr(t_rot), None, _ptr(t_rest),
                           int(raw_params is not None))
        g_col = _prep(grad_color, device, "grad_color")
        if g_col is None:
            g_col = torch.zeros((NUM_CHANNELS, H, W), **opts)
        g_seg, g_dep, g_alp = _prep(grad_segment, device, "grad_segment"), _prep(grad_depth, device, "grad_depth"), _prep(grad_alpha, device, "grad_alpha")
        pix = GsrPixelGrads(g_col.data_ptr(), _ptr(g_seg), _ptr(g_dep), _ptr(g_alp))
        state = GsrState(_ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), int(num_rendered), int(getattr(num_rendered, "num_visible", 0)))
        nscratch = L.gsr_backward_scratch_bytes(P)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=device)
        cap = max(int(capacity), 1)
        count = torch.zeros(1, dtype=torch.int32, device=device)
        if raw is None:
            nidx = int(L.gsr_packet_index_words(P))
            blob = torch.empty(packet_blob_words(P, cap), dtype=torch.int32, device=device)
            pk_ptr, idx_ptr = blob.data_ptr() + 4 * nidx, blob.data_ptr()
        else:
            blob, (pk_ptr, idx_ptr) = None, raw
        t_radii, t_alpha = radii.contiguous(), _prep(alpha, device, "alpha")
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = L.gsr_backward_packets(ctypes.byref(view), ctypes.byref(gin), t_radii.data_ptr(), ctypes.byref(state), t_alpha.data_ptr(),
                                    ctypes.byref(pix), pk_ptr, cap, count.data_ptr(), _ptr(means2D_grad), idx_ptr, scratch.data_ptr(),
                                    nscratch, stream)
        _lib.check(rc, "gsr_backward_packets")
        return blob, count


def packet_index_words(P):
    """gsr_packet_index_words: words reserved at the start of a view blob for the visibility index."""
    return (2 * ((P + 31) // 32) + 31) // 32 * 32


def packet_blob_words(P, capacity):
    """Words of a view blob with room for `capacity` packets (a multiple of 32, so that stacked blobs stay 128-byte aligned)."""
    return packet_index_words(P) + (int(capacity) * _lib.GSR_PACKET_WORDS + 31) // 32 * 32


def packet_blob_capacity(blob, P):
    return (blob.numel() - packet_index_words(P)) // _lib.GSR_PACKET_WORDS


def packet_blob_views(blob, P):
    """(packets int32[capacity, 16], visible-bit words int32[W], first-packet-index words int32[W]) of a view blob (the index
    is stored as pairs {bits, ~first}; `first` is only meaningful where bits != 0)."""
    W = (P + 31) // 32
    n = packet_index_words(P)
    idx = blob[:2 * W].view(W, 2)
    cap = packet_blob_capacity(blob, P)
    return blob[n:n + cap * _lib.GSR_PACKET_WORDS].view(cap, _lib.GSR_PACKET_WORDS), idx[:, 0], ~idx[:, 1]


def gather_packets(means3D, campos_all, sh_degree, sh_coeffs, blobs, out, num_class=NUM_CLASS):
    """gsr_gather_packets: sum the packets of all views (blobs int32[num_views, blob_words], campos_all f32[num_views, 3])
    into the dense gradient tensors of `out` (native names), writing EVERY row (zeros where no view saw the Gaussian)."""
    L = _lib.lib()
    device = means3D.device
    P = means3D.size(0)
    assert blobs.dim() == 2 and blobs.is_contiguous() and blobs.dtype == torch.int32
    nv = int(blobs.size(0))
    cap = packet_blob_capacity(blobs[0], P)
    with torch.cuda.device(device):
        g = lambda n: _ptr(out.get(n))
        pg = GsrParamGrads(g("means3D"), None, g("sh"), None, g("segments"), g("opacities"), g("scales"), g("rotations"), None, 0, g("sh_rest"))
        cp = _prep(campos_all, device, "campos")
        rc = L.gsr_gather_packets(P, int(sh_degree), int(sh_coeffs), int(num_class), means3D.data_ptr(), nv, cp.data_ptr(),
                                  blobs.data_ptr(), int(blobs.size(1)), cap, ctypes.byref(pg),
                                  torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "gsr_gather_packets")


def gather_packets_v(means3D, campos_all, sh_degree, sh_coeffs, view_ptrs, packet_off_words, index_off_words, capacity, out,
                     num_class=NUM_CLASS):
    """gsr_gather_packets_v: the same pass over one blob POINTER per view (device addresses, local or peer memory)."""
    L = _lib.lib()
    device = means3D.device
    P = means3D.size(0)
    nv = len(view_ptrs)
    with torch.cuda.device(device):
        g = lambda n: _ptr(out.get(n))
        pg = GsrParamGrads(g("means3D"), None, g("sh"), None, g("segments"), g("opacities"), g("scales"), g("rotations"), None, 0, g("sh_rest"))
        cp = _prep(campos_all, device, "campos")
        arr = (ctypes.c_void_p * nv)(*[int(p) for p in view_ptrs])
        rc = L.gsr_gather_packets_v(P, int(sh_degree), int(sh_coeffs), int(num_class), means3D.data_ptr(), nv, cp.data_ptr(), arr,
                                    int(packet_off_words), int(index_off_words), int(capacity), ctypes.byref(pg),
                                    torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "gsr_gather_packets_v")


def peer_alloc(nbytes, device):
    """gsr_peer_alloc on `device`: (device address, 64-byte handle as bytes). GSR_PEER_DISABLE=1 makes it fail (to exercise
    the callers' fallback to the NCCL exchange on nodes without CUDA IPC)."""
    import os

    if os.environ.get("GSR_PEER_DISABLE"):
        raise RuntimeError("peer buffers disabled by GSR_PEER_DISABLE")
    L = _lib.lib()
    with torch.cuda.device(device):
        ptr = ctypes.c_void_p()
        h = ctypes.create_string_buffer(_lib.GSR_PEER_HANDLE_BYTES)
        _lib.check(L.gsr_peer_alloc(int(nbytes), ctypes.byref(ptr), h), "gsr_peer_alloc")
        return int(ptr.value), bytes(h.raw)


def peer_open(handle, device):
    L = _lib.lib()
    with torch.cuda.device(device):
        ptr = ctypes.c_void_p()
        h = ctypes.create_string_buffer(bytes(handle), _lib.GSR_PEER_HANDLE_BYTES)
        _lib.check(L.gsr_peer_open(h, ctypes.byref(ptr)), "gsr_peer_open")
        return int(ptr.value)


def peer_copy(dst, src, nbytes, device, stream=None):
    """gsr_peer_copy: stream-ordered copy between device addresses of which either may be a peer's mapping (copy engines)."""
    with torch.cuda.device(device):
        st = (stream if stream is not None else torch.cuda.current_stream(device)).cuda_stream
        _lib.check(_lib.lib().gsr_peer_copy(ctypes.c_void_p(int(dst)), ctypes.c_void_p(int(src)), int(nbytes), st), "gsr_peer_copy")


def peer_close(ptr, device):
    with torch.cuda.device(device):
        _lib.check(_lib.lib().gsr_peer_close(ctypes.c_void_p(ptr)), "gsr_peer_close")


def peer_free(ptr, device):
    with torch.cuda.device(device):
        _lib.check(_lib.lib().gsr_peer_free(ctypes.c_void_p(ptr)), "gsr_peer_free")


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                subset=None):
        args = (means3D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp, raster_settings)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args[:-1] + tuple(raster_settings))  # copy before they can be corrupted
            try:
                num_rendered, color, depth, segment, alpha, radii, geomBuffer, binningBuffer, imgBuffer = _forward_native(*args, subset=subset)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, segment, alpha, radii, geomBuffer, binningBuffer, imgBuffer = _forward_native(*args, subset=subset)

        ctx.subset = subset
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.set_materialize_grads(False)  # unused outputs arrive as None -> NULL (= zeros) instead of zero-filled tensors
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(colors_precomp, segments, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
                              imgBuffer, alpha)
        return color, radii, depth, alpha, segment

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_segment):
        num_rendered = ctx.num_rendered
        rs = ctx.raster_settings
        colors_precomp, segments, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer, alpha = ctx.saved_tensors
        nig = ctx.needs_input_grad
        needs = {"means3D": nig[0], "means2D": nig[1], "sh": nig[2], "colors_precomp": nig[3], "segments": nig[4], "opacities": nig[5],
                 "scales": nig[6], "rotations": nig[7], "cov3Ds_precomp": nig[8]}
        args = (rs, means3D, radii, colors_precomp, segments, scales, rotations, cov3Ds_precomp, grad_color, grad_segment, grad_depth,
                grad_alpha, sh, geomBuffer, num_rendered, binningBuffer, imgBuffer, alpha)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(tuple(rs) + args[1:])
            try:
                g = _backward_native(*args, needs=needs, subset=ctx.subset)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            g = _backward_native(*args, needs=needs, subset=ctx.subset)
        return (g["means3D"], g["means2D"], g["sh"], g["colors_precomp"], g["segments"], g["opacities"], g["scales"], g["rotations"],
                g["cov3Ds_precomp"], None, None)


class _RasterizeGaussiansRaw(torch.autograd.Function):
    """Fused-activation entry: takes the model's RAW parameters (scene/gaussian_model.py:50-56) and runs
    get_opacity / get_segment / get_scaling / get_rotation / get_features (:100-124) inside the preprocess kernels, forward
    and backward; the gradients returned are those autograd would deliver to the raw parameters through the classic API."""

    @staticmethod
    def forward(ctx, xyz, means2D, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions, raster_settings):
        e = torch.empty(0)
        num_rendered, color, depth, segment, alpha, radii, geomBuffer, binningBuffer, imgBuffer = _forward_native(
            xyz, features_dc, e, segment_logits, opacity_logits, log_scales, quaternions, e, raster_settings, sh_rest=features_rest,
            raw_params=True)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(xyz, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions, radii, geomBuffer,
                              binningBuffer, imgBuffer, alpha)
        return color, radii, depth, alpha, segment

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_segment):
        rs = ctx.raster_settings
        xyz, f_dc, f_rest, seg, op, sc, rot, radii, geomBuffer, binningBuffer, imgBuffer, alpha = ctx.saved_tensors
        nig = ctx.needs_input_grad
        needs = {"means3D": nig[0], "means2D": nig[1], "sh": nig[2] or nig[3], "sh_rest": nig[2] or nig[3], "segments": nig[4],
                 "opacities": nig[5], "scales": nig[6], "rotations": nig[7]}
        e = torch.empty(0)
        g = _backward_native(rs, xyz, radii, e, seg, sc, rot, e, grad_color, grad_segment, grad_depth, grad_alpha, f_dc, geomBuffer,
                             ctx.num_rendered, binningBuffer, imgBuffer, alpha, needs=needs, sh_rest=f_rest, raw_params=True, opacities=op)
        return (g["means3D"], g["means2D"], g["sh"], g["sh_rest"], g["segments"], g["opacities"], g["scales"], g["rotations"], None)


def rasterize_gaussians_raw(xyz, means2D, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions,
                            raster_settings):
    return _RasterizeGaussiansRaw.apply(xyz, means2D, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions,
                                        raster_settings)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def mark_visible(positions, viewmatrix, projmatrix):
    """_C.mark_visible (rasterize_points.cu:223-242): bool[P], true where view-space z > 0.2."""
    L = _lib.lib()
    if not positions.is_cuda:
        raise RuntimeError("positions must be a CUDA tensor; libgsr has no CPU path")
    device = positions.device
    P = positions.size(0)
    with torch.cuda.device(device):
        present = torch.zeros(P, dtype=torch.bool, device=device)
        if P:
            pos = _prep(positions, device, "positions")
            vm, pm = _prep(viewmatrix, device, "viewmatrix"), _prep(projmatrix, device, "projmatrix")
            rc = L.gsr_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(), present.data_ptr(),
                                    torch.cuda.current_stream(device).cuda_stream)
            _lib.check(rc, "gsr_mark_visible")
        return present


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            rs = self.raster_settings
            visible = mark_visible(positions, rs.viewmatrix, rs.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, segments=None, scales=None, rotations=None,
                cov3D_precomp=None, subset=None):
        """As the reference (diff_gaussian_rasterization/__init__.py:194-235). `subset` (optional, not in the reference): an int32
        CUDA tensor of strictly ascending Gaussian indices -- only those are rendered, radii has one entry per list element and
        the gradients keep the full number of rows; replaces rendering masked copies (gaussian_renderer/__init__.py:239-268)."""
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        # absent optionals travel as empty tensors, the reference's "not provided" sentinel (__init__.py:208-221)
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if segments is None:
            segments = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, segments, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings, subset)

    def forward_parts(self, parts):
        """Render several resident sub-scenes as ONE scene without concatenating their tensors (the viewer's scene fusion,
        visualizer.py:196-226 / render.py:36). `parts`: list of dicts {means3D, opacities, shs | colors_precomp, scales + rotations |
        cov3D_precomp, segments}. Returns (color, radii, depth, alpha, segment) like forward(); radii covers the fused scene in
        part order. Render-only (no gradients)."""
        with torch.no_grad():
            R, color, depth, segment, alpha, radii, _, _, _ = _forward_parts_native(parts, self.raster_settings)
        return color, radii, depth, alpha, segment

    def forward_raw(self, xyz, means2D, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions):
        """Opt-in fused-activation entry: pass pc._xyz, pc._features_dc, pc._features_rest, pc._segment, pc._opacity,
        pc._scaling, pc._rotation instead of the get_* properties; same outputs as forward()."""
        return rasterize_gaussians_raw(xyz, means2D, features_dc, features_rest, segment_logits, opacity_logits, log_scales, quaternions,
                                       self.raster_settings)


def export_state(P, W, H, geomBuffer, binningBuffer, imgBuffer, num_rendered):
    """Test support: unpack the opaque forward state into the reference's Gaussian-id-indexed arrays
    (gsr_export_state). Returns a dict of tensors."""
    L = _lib.lib()
    device = geomBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    R = int(num_rendered)
    with torch.cuda.device(device):
        o = {
            "depths": torch.empty(P, dtype=torch.float32, device=device),
            "means2D": torch.empty((P, 2), dtype=torch.float32, device=device),
            "conic_opacity": torch.empty((P, 4), dtype=torch.float32, device=device),
            "rgb": torch.empty((P, 3), dtype=torch.float32, device=device),
            "clamped": torch.empty((P, 3), dtype=torch.uint8, device=device),
            "tiles_touched": torch.empty(P, dtype=torch.int32, device=device),
            "point_keys": torch.zeros(max(R, 1), dtype=torch.int64, device=device),
            "point_list": torch.zeros(max(R, 1), dtype=torch.int32, device=device),
            "ranges": torch.empty((T, 2), dtype=torch.int32, device=device),
            "n_contrib": torch.empty(H * W, dtype=torch.int32, device=device),
        }
        ex = GsrStateExport(*[o[k].data_ptr() for k in ["depths", "means2D", "conic_opacity", "rgb", "clamped", "tiles_touched", "point_keys",
                                                       "point_list", "ranges", "n_contrib"]])
        st = GsrState(_ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer), R)
        rc = L.gsr_export_state(P, W, H, ctypes.byref(st), ctypes.byref(ex), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "gsr_export_state")
        o["point_keys"] = o["point_keys"][:R]
        o["point_list"] = o["point_list"][:R]
        return o
