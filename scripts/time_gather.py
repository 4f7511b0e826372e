"""Times the gather-form packet rebuild (gsr_gather_packets) (round 1 also timed a per-view read-modify-write form, since removed: 2.07 ms vs 0.75 ms for 8 views)
for NV views of cfg3 on ONE GPU (no NCCL): python scripts/time_gather.py [NV]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

NV = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Pk = H.pkg()
mv = importlib.import_module(H.PKG_NAME + ".multiview")
D = Pk.diff_gaussian_rasterization
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS["cfg3"]
gs, _ = syn.make_scene("cfg3")
gs = H.to_dev(gs)
ug = H.to_dev(syn.upstream_grads(W, Hh, seed, with_depth=True))
bg = torch.tensor([0.0, 0.0, 0.0])
e = torch.empty(0)
sets, campos = [], []
for v in range(NV):
    cam = syn.make_camera(W, Hh, yaw_deg=45.0 * v)
    rs = H.settings(cam, bg)
    fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
    sets.append(mv.native_view_backward_packets(D, gs, rs, fwd, ug))
    campos.append(cam["campos"].cuda())
    del fwd
flat = mv.FlatGradients(P, "cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_g = timeit(lambda: mv.exchange_packets(D, None, flat, gs, sets, [campos], 3, world=1))
st = {}
mv.exchange_packets(D, None, flat, gs, sets, [campos], 3, world=1, state=st)
sets2 = []
for v in range(NV):
    cam = syn.make_camera(W, Hh, yaw_deg=45.0 * v)
    rs = H.settings(cam, bg)
    fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
    sets2.append(mv.native_view_backward_packets(D, gs, rs, fwd, ug, capacity=st["cap"]))
    del fwd
send = torch.stack([s[0] for s in sets2])
cps = torch.stack(campos)
t_k = timeit(lambda: D.gather_packets(gs["means3D"], cps, 3, 16, send, flat.backward_out()))
print("views=%d visible=%s gather_exchange=%.3f ms gather_kernel_only=%.3f ms" % (NV, [s[2] for s in sets], t_g, t_k))
