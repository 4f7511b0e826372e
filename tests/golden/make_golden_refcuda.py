"""Generate ref_cuda_small.npz from the reference's OWN CUDA rasterizer (oracle/_ref/ref_dgr_C.so, built by
oracle/build_ref.py) on a B200. Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden_refcuda.py gpurun_out/ref_cuda_small.npz'

then copy the file to tests/golden/. The inputs are regenerated from the seed by synthetic.make_scene.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

P, W, Hh, seed = 3000, 160, 112, 3
syn = H.synthetic()
gs, cam = syn.make_scene(P, W, Hh, seed=seed)
ug = syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True)
bg = torch.tensor([0.2, 0.3, 0.4])
rs = H.settings(cam, bg)
ref = H.run_ref(H.to_dev(gs), rs, H.to_dev(ug))
torch.cuda.synchronize()
c = lambda t: t.detach().cpu().numpy()
out = dict(P=P, W=W, H=Hh, seed=seed, bg=bg.numpy(), num_rendered=ref["num_rendered"], radii=c(ref["radii"]),
           point_list=c(ref["state"]["point_list"]), ranges=c(ref["state"]["ranges"]), n_contrib=c(ref["state"]["n_contrib"]),
           color=c(ref["color"]), depth=c(ref["depth"]), alpha=c(ref["alpha"]), segment=c(ref["segment"]))
names = {"means3D": "grad_means3D", "means2D": "grad_means2D", "sh": "grad_sh", "segments": "grad_segments", "opacities": "grad_opacities",
         "scales": "grad_scales", "rotations": "grad_rotations"}
for k, o in names.items():
    out[o] = c(ref["grads"][k])
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_cuda_small.npz")
os.makedirs(os.path.dirname(dst), exist_ok=True)
np.savez_compressed(dst, **out)
print("wrote", dst, "R =", ref["num_rendered"])
