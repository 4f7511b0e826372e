"""Shared test plumbing: package loader, scene -> device, our path, oracle A (reference CUDA, oracle/_ref),
oracle B (CPU C restatement) and the reference-state parsers of SURVEY.md Appendix B."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "3d_gaussian_magic_change-segment_3dgs_b200"


def pkg():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module(PKG_NAME)


def synthetic():
    pkg()
    return importlib.import_module(PKG_NAME + ".synthetic")


def cpu_oracle():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import cpu_oracle as O

    return O


_ref = {}


def ref_dgr():
    """The reference's own pybind module compiled by oracle/build_ref.py (None if not built)."""
    if "dgr" not in _ref:
        d = os.path.join(ROOT, "oracle", "_ref")
        if d not in sys.path:
            sys.path.insert(0, d)
        try:
            _ref["dgr"] = importlib.import_module("ref_dgr_C")
        except Exception as e:  # pragma: no cover
            print("oracle A unavailable:", e)
            _ref["dgr"] = None
    return _ref["dgr"]


def ref_knn():
    if "knn" not in _ref:
        d = os.path.join(ROOT, "oracle", "_ref")
        if d not in sys.path:
            sys.path.insert(0, d)
        try:
            _ref["knn"] = importlib.import_module("ref_knn_C")
        except Exception as e:  # pragma: no cover
            print("oracle A (knn) unavailable:", e)
            _ref["knn"] = None
    return _ref["knn"]


def to_dev(d, device="cuda"):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def settings(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False, prefiltered=False, device="cuda"):
    P = pkg()
    return P.GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg.to(device),
        scale_modifier=scale_modifier, viewmatrix=cam["viewmatrix"].to(device), projmatrix=cam["projmatrix"].to(device),
        sh_degree=sh_degree, campos=cam["campos"].to(device), prefiltered=prefiltered, debug=debug)


EMPTY = lambda: torch.Tensor([])


def run_ours(gs, rs, ug=None, colors_precomp=None, cov3D_precomp=None, use_sh=True, use_scale_rot=True, export=True):
    """Forward (+ backward when ug is given) through the drop-in module's native entry points.
    gs: dict of CUDA tensors. Returns dict with outputs, exported state and grads."""
    P = pkg()
    D = P.diff_gaussian_rasterization
    sh = gs["shs"] if use_sh else EMPTY()
    col = colors_precomp if colors_precomp is not None else EMPTY()
    seg = gs["segments"] if gs.get("segments") is not None else EMPTY()
    sc = gs["scales"] if use_scale_rot else EMPTY()
    rot = gs["rotations"] if use_scale_rot else EMPTY()
    cov = cov3D_precomp if cov3D_precomp is not None else EMPTY()
    R, color, depth, segment, alpha, radii, geom, binb, img = D._forward_native(gs["means3D"], sh, col, seg, gs["opacities"], sc, rot, cov, rs)
    out = dict(num_rendered=R, color=color, depth=depth, segment=segment, alpha=alpha, radii=radii, geom=geom, binning=binb, img=img)
    if export and gs["means3D"].shape[0] > 0:
        out["state"] = D.export_state(gs["means3D"].shape[0], rs.image_width, rs.image_height, geom, binb, img, R)
    if ug is not None:
        g = D._backward_native(rs, gs["means3D"], radii, col, seg, sc, rot, cov, ug["color"], ug.get("segment"), ug.get("depth"),
                               ug.get("alpha"), sh, geom, R, binb, img, alpha)
        out["grads"] = g
    return out


def _align(o, a=128):
    return (o + a - 1) // a * a


def parse_ref_state(P, W, H, R, geom, binb, img):
    """Appendix B: carve the reference's geomBuffer / binningBuffer / imgBuffer (uint8 CUDA tensors)."""
    N = W * H
    st = {}
    o = 0

    def take(buf, off, dtype, count, shape=None):
        itemsize = torch.tensor([], dtype=dtype).element_size()
        off = _align(off)
        t = buf[off:off + count * itemsize].view(dtype)
        return (t.view(shape) if shape else t), off + count * itemsize

    st["depths"], o = take(geom, o, torch.float32, P)
    st["clamped"], o = take(geom, o, torch.uint8, 3 * P, (P, 3))
    st["internal_radii"], o = take(geom, o, torch.int32, P)
    st["means2D"], o = take(geom, o, torch.float32, 2 * P, (P, 2))
    st["cov3D"], o = take(geom, o, torch.float32, 6 * P, (P, 6))
    st["conic_opacity"], o = take(geom, o, torch.float32, 4 * P, (P, 4))
    st["rgb"], o = take(geom, o, torch.float32, 3 * P, (P, 3))
    st["tiles_touched"], o = take(geom, o, torch.int32, P)
    o = 0
    st["point_list"], o = take(binb, o, torch.int32, R)
    st["point_list_unsorted"], o = take(binb, o, torch.int32, R)
    st["point_keys"], o = take(binb, o, torch.int64, R)
    st["point_keys_unsorted"], o = take(binb, o, torch.int64, R)
    o = 0
    st["n_contrib"], o = take(img, o, torch.int32, N)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    rng, o = take(img, o, torch.int32, 2 * N, (N, 2))
    st["ranges"] = rng[:T]
    return st


def run_ref(gs, rs, ug=None, colors_precomp=None, cov3D_precomp=None, use_sh=True, use_scale_rot=True):
    """Same call through the reference's own pybind entry points (ext.cpp:15-19), positional argument order of
    diff_gaussian_rasterization/__init__.py:63-84 and :114-139."""
    C = ref_dgr()
    dev = gs["means3D"].device
    e = lambda: torch.empty(0, device="cpu")
    sh = gs["shs"] if use_sh else e()
    col = colors_precomp if colors_precomp is not None else e()
    seg = gs["segments"]
    sc = gs["scales"] if use_scale_rot else e()
    rot = gs["rotations"] if use_scale_rot else e()
    cov = cov3D_precomp if cov3D_precomp is not None else e()
    R, color, depth, segment, alpha, radii, geom, binb, img = C.rasterize_gaussians(
        rs.bg, gs["means3D"], col, seg, gs["opacities"], sc, rot, rs.scale_modifier, cov, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
        rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
    out = dict(num_rendered=R, color=color, depth=depth, segment=segment, alpha=alpha, radii=radii, geom=geom, binning=binb, img=img)
    P = gs["means3D"].shape[0]
    if P > 0:
        out["state"] = parse_ref_state(P, rs.image_width, rs.image_height, R, geom, binb, img)
    if ug is not None:
        z = lambda c: torch.zeros((c, rs.image_height, rs.image_width), device=dev)
        gseg = ug.get("segment") if ug.get("segment") is not None else z(2)
        gdep = ug.get("depth") if ug.get("depth") is not None else z(1)
        galp = ug.get("alpha") if ug.get("alpha") is not None else z(1)
        (g_m2d, g_col, g_op, g_m3d, g_cov, g_sh, g_sc, g_rot, g_seg) = C.rasterize_gaussians_backward(
            rs.bg, gs["means3D"], radii, col, seg, sc, rot, rs.scale_modifier, cov, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            ug["color"], gseg, gdep, galp, sh, rs.sh_degree, rs.campos, geom, R, binb, img, alpha, rs.debug)
        out["grads"] = dict(means3D=g_m3d, means2D=g_m2d, sh=g_sh, colors_precomp=g_col, segments=g_seg, opacities=g_op, scales=g_sc,
                            rotations=g_rot, cov3Ds_precomp=g_cov)
    return out


def run_cpu_oracle(gs, cam, bg, ug=None, sh_degree=3, scale_modifier=1.0, colors_precomp=None, cov3D_precomp=None, use_sh=True,
                   use_scale_rot=True):
    O = cpu_oracle()
    n = lambda t: None if t is None else t.detach().cpu().numpy()
    st = O.forward(n(gs["means3D"]), n(gs["opacities"]), cam["W"], cam["H"], cam["tanfovx"], cam["tanfovy"], n(cam["viewmatrix"]),
                   n(cam["projmatrix"]), n(cam["campos"]), n(bg), shs=n(gs["shs"]) if use_sh else None,
                   colors_precomp=n(colors_precomp), segments=n(gs.get("segments")), scales=n(gs["scales"]) if use_scale_rot else None,
                   rotations=n(gs["rotations"]) if use_scale_rot else None, cov3D_precomp=n(cov3D_precomp), sh_degree=sh_degree,
                   scale_modifier=scale_modifier)
    if ug is not None:
        st["grads"] = O.backward(st, n(ug["color"]), n(ug.get("depth")), n(ug.get("alpha")), n(ug.get("segment")))
    return st


def rel_linf(a, b):
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b)).double()
    a, b = a.reshape(-1), b.reshape(-1)
    denom = float(b.abs().max()) if b.numel() else 0.0
    if denom == 0.0:
        return float(a.abs().max()) if a.numel() else 0.0
    return float((a - b).abs().max()) / denom
