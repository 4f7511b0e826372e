"""ctypes/numpy front end of oracle B (oracle/gsr_oracle.c, the plain-C CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of gsr_oracle.c. Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by the product package.

`forward()` mirrors CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-344) stage by stage and returns
every intermediate the reference keeps in its geometry/binning/image state, so tests can compare stage-wise.
`backward()` mirrors Rasterizer::backward (rasterizer_impl.cu:348-458) and returns the nine gradients of
RasterizeGaussiansBackwardCUDA (rasterize_points.cu:127-221).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "gsr_oracle.c")
_OUT = os.path.join(HERE, "_build", "libgsr_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off oracle/gsr_oracle.c -> oracle/_build/libgsr_oracle.so"""
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c11", _SRC, "-o", _OUT, "-lm"]
    subprocess.check_call(cmd)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_inclusive_sum.restype = ctypes.c_uint32
        _lib.orc_higher_msb.restype = ctypes.c_uint32
        _lib.orc_preprocess.restype = ctypes.c_int
    return _lib


def _p(a):
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i(v):
    return ctypes.c_int(int(v))


def _f(v):
    return ctypes.c_float(float(v))


def forward(means3D, opacities, W, H, tanfovx, tanfovy, viewmatrix, projmatrix, campos, bg, shs=None, colors_precomp=None,
            segments=None, scales=None, rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0,
            prefiltered=False, row_stride=1, row_offset=0):
    """All array arguments are numpy (any float dtype; converted to fp32). viewmatrix/projmatrix are the 16 floats
    exactly as the reference receives them (transposed, i.e. column-major for the kernels)."""
    L = lib()
    means3D = _f32(means3D).reshape(-1, 3)
    P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    shs, colors_precomp, segments = _f32(shs), _f32(colors_precomp), _f32(segments)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    view, proj = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16)
    campos, bg = _f32(campos).reshape(3), _f32(bg).reshape(3)
    M = 0 if shs is None else shs.shape[1]
    S = 0 if segments is None else segments.shape[1]
    N = W * H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    st = dict(P=P, W=W, H=H, M=M, S=S, D=sh_degree)
    st["radii"] = np.zeros(P, np.int32)
    st["means2D"] = np.zeros((P, 2), np.float32)
    st["depths"] = np.zeros(P, np.float32)
    st["cov3D"] = np.zeros((P, 6), np.float32)
    st["rgb"] = np.zeros((P, 3), np.float32)
    st["conic_opacity"] = np.zeros((P, 4), np.float32)
    st["clamped"] = np.zeros((P, 3), np.uint8)
    st["tiles_touched"] = np.zeros(P, np.uint32)
    color = np.zeros((3, H, W), np.float32)
    segment = np.zeros((S, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    alpha = np.zeros((1, H, W), np.float32)
    st.update(color=color, segment=segment, depth=depth, alpha=alpha)
    if P == 0:  # rasterize_points.cu:87 -- the core is skipped, images stay zero
        st.update(num_rendered=0, point_offsets=np.zeros(0, np.uint32), keys_unsorted=np.zeros(0, np.uint64),
                  point_list_unsorted=np.zeros(0, np.uint32), keys=np.zeros(0, np.uint64), point_list=np.zeros(0, np.uint32),
                  ranges=np.zeros((T, 2), np.uint32), n_contrib=np.zeros(N, np.uint32))
        return st
    vis = L.orc_preprocess(_i(P), _i(sh_degree), _i(M), _p(means3D), _p(scales), _f(scale_modifier), _p(rotations), _p(opacities),
                           _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(view), _p(proj), _p(campos), _i(W), _i(H),
                           _f(tanfovx), _f(tanfovy), _i(prefiltered), _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]),
                           _p(st["cov3D"]), _p(st["rgb"]), _p(st["conic_opacity"]), _p(st["clamped"]), _p(st["tiles_touched"]))
    if vis < 0:
        raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
    st["num_visible"] = vis
    st["point_offsets"] = np.zeros(P, np.uint32)
    R = int(L.orc_inclusive_sum(_i(P), _p(st["tiles_touched"]), _p(st["point_offsets"])))
    st["num_rendered"] = R
    ku, vu = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    L.orc_duplicate_with_keys(_i(P), _p(st["means2D"]), _p(st["depths"]), _p(st["point_offsets"]), _p(st["radii"]), _i(W), _i(H),
                              _p(ku), _p(vu))
    bit = int(L.orc_higher_msb(ctypes.c_uint32(T)))
    ks, vs = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    L.orc_sort_pairs(ctypes.c_size_t(R), _p(ku), _p(ks), _p(vu), _p(vs), _i(32 + bit))
    ranges = np.zeros((T, 2), np.uint32)
    L.orc_identify_tile_ranges(ctypes.c_size_t(R), _p(ks), _i(T), _p(ranges))
    st.update(keys_unsorted=ku, point_list_unsorted=vu, keys=ks, point_list=vs, ranges=ranges, sort_bits=32 + bit)
    feats = colors_precomp if colors_precomp is not None else st["rgb"]
    n_contrib = np.zeros(N, np.uint32)
    L.orc_render_forward(_i(W), _i(H), _i(S), _p(ranges), _p(vs), _p(st["means2D"]), _p(feats), _p(segments), _p(st["depths"]),
                         _p(st["conic_opacity"]), _p(bg), _p(color), _p(segment), _p(depth), _p(alpha), _p(n_contrib),
                         _i(row_stride), _i(row_offset))
    st["n_contrib"] = n_contrib
    st["_rows"] = (row_stride, row_offset)
    st["_inputs"] = dict(means3D=means3D, opacities=opacities, shs=shs, colors_precomp=colors_precomp, segments=segments,
                         scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, view=view, proj=proj, campos=campos,
                         bg=bg, tanfovx=tanfovx, tanfovy=tanfovy, scale_modifier=scale_modifier)
    return st


def backward(st, grad_color, grad_depth=None, grad_alpha=None, grad_segment=None):
    """Returns dict of the reference's nine gradients (+ the intermediate per-Gaussian sums) for the state `st`
    produced by forward(). Missing upstream grads are zeros, as autograd would materialise them."""
    L = lib()
    P, W, H, M, S, D = st["P"], st["W"], st["H"], st["M"], st["S"], st["D"]
    inp = st["_inputs"] if P else None
    out = dict(
        grad_means3D=np.zeros((P, 3), np.float32), grad_means2D=np.zeros((P, 3), np.float32),
        grad_sh=np.zeros((P, M, 3), np.float32), grad_colors_precomp=np.zeros((P, 3), np.float32),
        grad_segments=np.zeros((P, S), np.float32), grad_opacities=np.zeros((P, 1), np.float32),
        grad_scales=np.zeros((P, 3), np.float32), grad_rotations=np.zeros((P, 4), np.float32),
        grad_cov3Ds_precomp=np.zeros((P, 6), np.float32))
    if P == 0:
        return out
    gc = _f32(grad_color).reshape(3, H, W)
    gd = np.zeros((1, H, W), np.float32) if grad_depth is None else _f32(grad_depth).reshape(1, H, W)
    ga = np.zeros((1, H, W), np.float32) if grad_alpha is None else _f32(grad_alpha).reshape(1, H, W)
    gs = np.zeros((S, H, W), np.float32) if grad_segment is None else _f32(grad_segment).reshape(S, H, W)
    feats = inp["colors_precomp"] if inp["colors_precomp"] is not None else st["rgb"]
    dmean2D, dconic = np.zeros((P, 2), np.float64), np.zeros((P, 3), np.float64)
    dopacity, dcolors = np.zeros(P, np.float64), np.zeros((P, 3), np.float64)
    dsegments, ddepths = np.zeros((P, max(S, 1)), np.float64), np.zeros(P, np.float64)
    L.orc_render_backward(_i(W), _i(H), _i(S), _p(st["ranges"]), _p(st["point_list"]), _p(inp["bg"]), _p(st["means2D"]),
                          _p(st["conic_opacity"]), _p(feats), _p(inp["segments"]), _p(st["depths"]), _p(st["alpha"]),
                          _p(st["n_contrib"]), _p(gc), _p(gs), _p(gd), _p(ga), _p(dmean2D), _p(dconic), _p(dopacity), _p(dcolors),
                          _p(dsegments), _p(ddepths), _i(st.get("_rows", (1, 0))[0]), _i(st.get("_rows", (1, 0))[1]))
    cov3Ds = inp["cov3D_precomp"] if inp["cov3D_precomp"] is not None else st["cov3D"]
    m2, cn = dmean2D.astype(np.float32), dconic.astype(np.float32)
    dc, dd = dcolors.astype(np.float32), ddepths.astype(np.float32)
    have_sh, have_sc = inp["shs"] is not None, inp["scales"] is not None
    L.orc_preprocess_backward(_i(P), _i(D), _i(M), _p(inp["means3D"]), _p(st["radii"]), _p(inp["shs"]), _p(st["clamped"]),
                              _p(inp["scales"]), _p(inp["rotations"]), _f(inp["scale_modifier"]), _p(cov3Ds), _p(inp["view"]),
                              _p(inp["proj"]), _p(inp["campos"]), _i(W), _i(H), _f(inp["tanfovx"]), _f(inp["tanfovy"]), _p(m2),
                              _p(cn), _p(dc), _p(dd), _p(out["grad_means3D"]), _p(out["grad_cov3Ds_precomp"]),
                              _p(out["grad_sh"]) if have_sh else _p(None), _p(out["grad_scales"]) if have_sc else _p(None),
                              _p(out["grad_rotations"]) if have_sc else _p(None))
    out["grad_means2D"][:, :2] = m2
    out["grad_colors_precomp"][:] = dc
    out["grad_opacities"][:, 0] = dopacity.astype(np.float32)
    if S:
        out["grad_segments"][:] = dsegments[:, :S].astype(np.float32)
    out["_dL_dconic"] = cn
    out["_dL_ddepths"] = dd
    return out


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    means3D = _f32(means3D).reshape(-1, 3)
    out = np.zeros(means3D.shape[0], np.uint8)
    L.orc_mark_visible(_i(means3D.shape[0]), _p(means3D), _p(_f32(viewmatrix).reshape(16)), _p(_f32(projmatrix).reshape(16)), _p(out))
    return out.astype(bool)


def knn_dist2(points):
    L = lib()
    points = _f32(points).reshape(-1, 3)
    out = np.zeros(points.shape[0], np.float32)
    L.orc_knn_dist2(_i(points.shape[0]), _p(points), _p(out))
    return out


if __name__ == "__main__":
    print(build(force=True))
