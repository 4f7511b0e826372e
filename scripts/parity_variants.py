"""Gradient parity of the current compositing-backward variant against the reference CUDA build at a BASELINE workload:

    GSR_BWD_VARIANT=7 python scripts/parity_variants.py cfg2

Prints, per gradient tensor, the tensor-wise relative L-inf error (the north star's <= 1e-4 bar) of ours vs the reference, and of a
SECOND reference run vs the first (the reference's own atomic-order noise), plus where the largest difference sits."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS[wl]
gs, cam = syn.make_scene(wl)
gs = H.to_dev(gs)
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True))
rs = H.settings(cam, torch.tensor([0.2, 0.1, 0.3]))
ours = H.run_ours(gs, rs, ug, export=False)["grads"]
ref = H.run_ref(gs, rs, ug)["grads"]
ref2 = H.run_ref(gs, rs, ug)["grads"]
out = {"variant": os.environ.get("GSR_BWD_VARIANT", "default"), "workload": wl}
for k in ["means3D", "means2D", "sh", "segments", "opacities", "scales", "rotations"]:
    a, b, c = ours[k], ref[k].reshape(ours[k].shape), ref2[k].reshape(ours[k].shape)
    d = (a - b).abs()
    i = int(d.argmax())
    out[k] = {"ours_vs_ref": H.rel_linf(a, b), "ref_vs_ref": H.rel_linf(c, b), "max_abs_ref": float(b.abs().max()),
              "at": {"ours": float(a.reshape(-1)[i]), "ref": float(b.reshape(-1)[i]), "ref2": float(c.reshape(-1)[i])}}
print(json.dumps(out))
