"""PLY checkpoints in the reference's layout (scene/gaussian_model.py:186-233 save_ply, :262-309 load_ply) without
the `plyfile` dependency: one `vertex` element of float32 properties

    x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*(M-1)-1)  opacity  segment_0..(C-1)  scale_0..2  rot_0..3

written as `binary_little_endian 1.0` (what plyfile writes for the reference). The SH blocks are stored channel-major, as the
reference does (`_features_*.transpose(1, 2).flatten(start_dim=1)`). Values are the RAW parameters (logits, log-scales,
un-normalised quaternions) -- exactly what optim.FlatParameters holds -- so a file written by the reference loads straight into
the native trainer and vice versa. Host-side code (numpy); no CUDA involved."""
import os

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def attribute_names(sh_coeffs, num_class):
    """construct_list_of_attributes (scene/gaussian_model.py:186-205)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += ["f_dc_%d" % i for i in range(3)]
    names += ["f_rest_%d" % i for i in range(3 * (sh_coeffs - 1))]
    names += ["opacity"]
    names += ["segment_%d" % i for i in range(num_class)]
    names += ["scale_%d" % i for i in range(3)]
    names += ["rot_%d" % i for i in range(4)]
    return names


def save_ply(path, tensors):
    """tensors: dict with means3D [P,3], features_dc [P,1,3], features_rest [P,M-1,3], opacities [P,1], segments [P,C],
    scales [P,3], rotations [P,4] (e.g. `FlatParameters.views`)."""
    t = {k: v.detach().cpu().float() for k, v in tensors.items()}
    P = t["means3D"].shape[0]
    xyz = t["means3D"].numpy()
    f_dc = t["features_dc"].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    f_rest = t["features_rest"].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    cols = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, t["opacities"].reshape(P, 1).numpy(), t["segments"].numpy(),
                           t["scales"].numpy(), t["rotations"].numpy()), axis=1).astype("<f4")
    names = attribute_names(t["features_rest"].shape[1] + 1, t["segments"].shape[1])
    assert cols.shape[1] == len(names)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % n for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(cols).tobytes())


def read_vertices(path):
    """The `vertex` element of a PLY file as a numpy structured array (binary little/big endian or ascii; scalar properties)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s is not a PLY file" % path)
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("unexpected end of PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError("list properties are not supported")
                cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError("the first element must be `vertex`")
        el = elements[0]
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=el["count"], ndmin=2)
            out = np.empty(el["count"], dtype=[(n, t) for n, t in el["props"]])
            for i, (n, _) in enumerate(el["props"]):
                out[n] = rows[:, i]
            return out
        order = "<" if fmt == "binary_little_endian" else ">"
        dtype = np.dtype([(n, order + t) for n, t in el["props"]])
        return np.frombuffer(f.read(dtype.itemsize * el["count"]), dtype=dtype, count=el["count"])


def load_ply(path, max_sh_degree=3, num_class=2, device="cpu"):
    """load_ply (scene/gaussian_model.py:262-309): dict of RAW parameter tensors in the reference's shapes."""
    v = read_vertices(path)
    names = v.dtype.names
    P = v.shape[0]
    col = lambda n: np.asarray(v[n], dtype=np.float32)
    xyz = np.stack((col("x"), col("y"), col("z")), axis=1)
    opacities = col("opacity")[..., np.newaxis]
    segments = np.stack([col("segment_%d" % i) for i in range(num_class)], axis=1)
    features_dc = np.zeros((P, 3, 1), np.float32)
    for c in range(3):
        features_dc[:, c, 0] = col("f_dc_%d" % c)
    extra = sorted([n for n in names if n.startswith("f_rest_")], key=lambda x: int(x.split("_")[-1]))
    if len(extra) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise ValueError("expected %d f_rest_* properties for SH degree %d, found %d" % (3 * (max_sh_degree + 1) ** 2 - 3, max_sh_degree, len(extra)))
    features_extra = np.stack([col(n) for n in extra], axis=1).reshape(P, 3, (max_sh_degree + 1) ** 2 - 1) if extra else np.zeros((P, 3, 0), np.float32)
    scales = np.stack([col(n) for n in sorted([n for n in names if n.startswith("scale_")], key=lambda x: int(x.split("_")[-1]))], axis=1)
    rots = np.stack([col(n) for n in sorted([n for n in names if n.startswith("rot")], key=lambda x: int(x.split("_")[-1]))], axis=1)
    t = lambda a: torch.tensor(a, dtype=torch.float, device=device)
    return {"means3D": t(xyz), "features_dc": t(features_dc).transpose(1, 2).contiguous(), "features_rest": t(features_extra).transpose(1, 2).contiguous(),
            "opacities": t(opacities), "segments": t(segments), "scales": t(scales), "rotations": t(rots)}
