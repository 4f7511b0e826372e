"""A/B of the compositing-backward variants at a BASELINE workload (one process per variant: the variant is read once):

    GSR_BWD_VARIANT=2 python scripts/ab_bwd.py save [cfg3]     # 2 pixels/lane scalar kernel (round 1): saves its gradients
    GSR_BWD_VARIANT=5 python scripts/ab_bwd.py cmp  [cfg3]     # packed fp32x2 kernel: stage times + max relative difference

Prints the per-stage device times (gsr_set_profiling, CUDA events inside libgsr) median of 21 steps."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "cmp"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
Pk = H.pkg()
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS[wl]
gs, cam = syn.make_scene(wl)
gs = H.to_dev(gs)
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True, with_segment=True, with_alpha=True))
rs = H.settings(cam, torch.tensor([0.05, 0.1, 0.15]))
L = Pk._lib.lib()
out = H.run_ours(gs, rs, ug, export=False)
torch.cuda.synchronize()
L.gsr_set_profiling(1)
samples = {}
for _ in range(21):
    o = H.run_ours(gs, rs, ug, export=False)
    torch.cuda.synchronize()
    for k, v in Pk._lib.stage_times().items():
        samples.setdefault(k, []).append(v)
acc = {k: sorted(v)[len(v) // 2] for k, v in samples.items()}  # median: the first stage's event also sees allocator hiccups
L.gsr_set_profiling(0)
tag = os.environ.get("GSR_BWD_VARIANT", "default") + "/" + os.environ.get("GSR_FILL_STREAM", "side")
print(json.dumps({"variant": tag, "workload": wl, "stages_ms": {k: round(v, 4) for k, v in acc.items()}}))
path = os.path.join(ROOT, "gpurun_out", "ab_bwd_%s.pt" % wl)
if mode == "save":
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({k: v.cpu() for k, v in out["grads"].items() if v is not None}, path)
elif os.path.exists(path):
    ref = torch.load(path)
    print(json.dumps({"rel_linf_vs_saved": {k: H.rel_linf(out["grads"][k], ref[k]) for k in ref}}))
