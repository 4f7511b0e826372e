"""Seeded synthetic scenes and cameras of the BASELINE.json shapes (SURVEY.md section 8d recipe).

Used by tests/, bench.py and __graft_entry__.smoke(); no datasets or checkpoints exist offline. Everything is
generated on the CPU with a seeded torch.Generator and moved to the GPU by the caller, so every implementation
(libgsr, the reference CUDA build, the CPU oracle) sees bit-identical inputs.

The camera maths restates the reference's getWorld2View2 / getProjectionMatrix (utils/graphics_utils.py:38-74)
and Camera.__init__ (scene/cameras.py:58-61): matrices are handed to the rasterizer TRANSPOSED.
"""
import math

import numpy as np
import torch

CONFIGS = {
    # name: (P, W, H, seed)
    "cfg1": (100_000, 800, 800, 0),      # plumbing / parity
    "cfg2": (3_000_000, 1297, 840, 1),   # Mip-NeRF360 shape
    "cfg3": (6_000_000, 1920, 1080, 2),  # headline: 1080p, 6M Gaussians, depth
    "cfg4": (6_000_000, 1920, 1080, 3),  # multi-view (8 views/step)
    "cfg5": (10_000_000, 3840, 2160, 4), # viewer fusion (4 merged sub-scenes), forward only
}


def make_camera(W, H, fovx_deg=60.0, yaw_deg=0.0, radius=5.0, znear=0.01, zfar=100.0):
    """Camera on a circle of `radius` around the cube centre, looking at it; yaw 0 = the recipe's R=I, t=(0,0,5)."""
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx / 2.0) * H / W)
    th = math.radians(yaw_deg)
    # camera-to-world rotation (the reference's R) and world-to-camera translation t = -R^T C
    R = np.array([[math.cos(th), 0.0, math.sin(th)], [0.0, 1.0, 0.0], [-math.sin(th), 0.0, math.cos(th)]], dtype=np.float64)
    C = np.array([-radius * math.sin(th), 0.0, -radius * math.cos(th)], dtype=np.float64)
    t = -R.T @ C
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    w2c = np.float32(Rt)  # getWorld2View2 with translate=0, scale=1
    tanx, tany = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right = tany * znear, tanx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    world_view = torch.tensor(w2c).transpose(0, 1).contiguous()
    proj = Pm.transpose(0, 1)
    full = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = world_view.inverse()[3, :3].contiguous()
    return dict(W=W, H=H, tanfovx=tanx, tanfovy=tany, viewmatrix=world_view, projmatrix=full, campos=campos, fovx=fovx, fovy=fovy)


def make_gaussians(P, seed, scale_P=None, num_class=2):
    """Post-activation Gaussian parameters, the rasterizer's inputs."""
    g = torch.Generator().manual_seed(seed)
    mu = math.log(0.20 * 12.0 / float(scale_P or P) ** (1.0 / 3.0))
    means3D = (torch.rand(P, 3, generator=g) * 12.0 - 6.0)
    scales = torch.exp(torch.randn(P, 3, generator=g) * 0.5 + mu)
    rotations = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=1)
    opacities = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0)
    shs = torch.randn(P, 16, 3, generator=g) * 0.15
    shs[:, 0, :] = torch.randn(P, 3, generator=g)
    segments = torch.sigmoid(torch.randn(P, num_class, generator=g)) if num_class else None
    return dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities, shs=shs, segments=segments)


def make_scene(name_or_P, W=None, H=None, seed=None, yaw_deg=0.0):
    """A named BASELINE config ("cfg1".."cfg5") or an ad-hoc (P, W, H, seed)."""
    if isinstance(name_or_P, str):
        P, W, H, seed = CONFIGS[name_or_P]
        if name_or_P == "cfg5":  # four sub-scenes of 2.5M, seeds 4..7, concatenated like visualizer._merge_scenes
            parts = [make_gaussians(P // 4, seed + i, scale_P=P) for i in range(4)]
            gs = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
        else:
            gs = make_gaussians(P, seed)
    else:
        gs = make_gaussians(int(name_or_P), 0 if seed is None else seed)
    cam = make_camera(W, H, yaw_deg=yaw_deg)
    return gs, cam


def upstream_grads(W, H, seed, with_depth=True, with_segment=False, with_alpha=False, num_class=2):
    """Seeded dL/d(outputs) for rasterizer-only fwd+bwd timing and parity."""
    g = torch.Generator().manual_seed(1000 + seed)
    N = W * H
    out = {"color": torch.randn(3, H, W, generator=g) / (3 * N)}
    out["depth"] = torch.randn(1, H, W, generator=g) / N if with_depth else None
    out["segment"] = torch.randn(num_class, H, W, generator=g) / (num_class * N) if with_segment else None
    out["alpha"] = torch.randn(1, H, W, generator=g) / N if with_alpha else None
    return out
