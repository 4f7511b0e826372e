"""torchrun worker of tests/test_multigpu_gpu.py: N ranks, one view each; the packet exchange must equal the dense all-reduce
of the per-rank dense gradients (<= 1e-6 relative) and leave bitwise-identical buffers on every rank."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
Pk = H.pkg()
mv = importlib.import_module(H.PKG_NAME + ".multiview")
D = Pk.diff_gaussian_rasterization
syn = H.synthetic()
P, W, Hh = 60_000, 400, 304
gs, _ = syn.make_scene(P, W, Hh, seed=81)
gs = {k: v.to(dev) for k, v in gs.items()}
cams = [syn.make_camera(W, Hh, yaw_deg=45.0 * r) for r in range(world)]
ug = {k: (v.to(dev) if v is not None else None) for k, v in syn.upstream_grads(W, Hh, 81, with_depth=True, with_segment=True, with_alpha=True).items()}
bg = torch.tensor([0.1, 0.2, 0.3])
rs = H.settings(cams[rank], bg, device=dev)
e = torch.empty(0)
fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
# dense path: flat buffer + all-reduce
dense = mv.FlatGradients(P, dev)
mv.native_view_backward(D, gs, rs, fwd, ug, dense, first=True)
dense.allreduce(dist)
# packet path
flat = mv.FlatGradients(P, dev)
flat.buffer.fill_(float(rank + 1))  # the gather pass overwrites every row
sets = [mv.native_view_backward_packets(D, gs, rs, fwd, ug)]
campos = [[c["campos"].to(dev)] for c in cams]
mv.exchange_packets(D, dist, flat, gs, sets, campos, 3, world)
torch.cuda.synchronize()
err = float((flat.packed() - dense.packed()).abs().max()) / float(dense.packed().abs().max())
assert err <= 2e-5, err  # two backward runs: fp32 atomic-order noise
# replicas bitwise identical
ref = flat.packed()
dist.broadcast(ref, 0)
assert torch.equal(ref, flat.packed()), "replicas differ"
# peer-memory path, both forms (3 steps each: both buffers): push = the copy engines write every blob into the peers' receive slots
# and the gather kernel reads local memory; pull = the gather kernel reads the peers' buffers over NVLink
for peer_mode in ("push", "pull"):
    px = mv.PeerPacketExchange(D, dist, P, 1, rank, world, dev, mode=peer_mode)
    for step in range(3):
        peer = mv.FlatGradients(P, dev)
        peer.buffer.fill_(-3.0)
        cnt = px.view_backward(gs, rs, fwd, ug, 0)
        px.exchange(peer, gs, campos, 3)
        torch.cuda.synchronize()
        assert int(cnt) == sets[0][2]
        err_p = float((peer.packed() - dense.packed()).abs().max()) / float(dense.packed().abs().max())
        assert err_p <= 2e-5, (peer_mode, step, err_p)
        ref = peer.packed()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, peer.packed()), "peer replicas differ (%s)" % peer_mode
    px.close()
# raw-parameter layout end to end: flat parameters -> fused-activation forward -> packets into peer buffers -> gather into the
# split flat gradient buffer -> fused Adam; replicas must stay bitwise identical after the optimiser step
optim = importlib.import_module(H.PKG_NAME + ".optim")
eps_ = 1e-6
logit = lambda p: torch.log(p.clamp(eps_, 1 - eps_) / (1 - p.clamp(eps_, 1 - eps_)))
params = optim.FlatParameters.from_tensors({"means3D": gs["means3D"], "features_dc": gs["shs"][:, :1].contiguous(),
                                            "features_rest": gs["shs"][:, 1:].contiguous(), "segments": logit(gs["segments"]),
                                            "opacities": logit(gs["opacities"]), "scales": torch.log(gs["scales"]), "rotations": gs["rotations"] * 1.5})
rgrads = mv.FlatGradients(P, dev, split_sh=True)
opt = optim.FusedAdam(params, rgrads, {"xyz": 1e-4, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "segment": 0.01, "scaling": 5e-3, "rotation": 1e-3})
px2 = mv.PeerPacketExchange(D, dist, P, 1, rank, world, dev)
# dense raw-parameter reference: each rank's own dense gradient, all-reduced
fwd_r = mv.native_view_forward(D, params.views, rs)
dense_r = mv.FlatGradients(P, dev, split_sh=True)
mv.native_view_backward(D, params.views, rs, fwd_r, ug, dense_r, first=True)
dense_r.allreduce(dist)
px2.view_backward(params.views, rs, fwd_r, ug, 0)
px2.exchange(rgrads, params.views, campos, 3)
torch.cuda.synchronize()
err_r = float((rgrads.packed() - dense_r.packed()).abs().max()) / float(dense_r.packed().abs().max())
assert err_r <= 2e-5, err_r
before = params.buffer.clone()
opt.step()
torch.cuda.synchronize()
assert not torch.equal(before, params.buffer)
ref = params.buffer.clone()
dist.broadcast(ref, 0)
assert torch.equal(ref, params.buffer), "parameter replicas differ after the fused Adam step"
px2.close()
# the native training driver on all ranks: 2 x world views per step, densification at step 3, opacity reset at step 4; replicas
# must hold bitwise-identical parameters and Adam moments after every step
trainer = importlib.import_module(H.PKG_NAME + ".trainer")
o = trainer.OptimizationParams(densify_from_iter=1, densification_interval=3, opacity_reset_interval=4, densify_until_iter=100)
init = {k: v.clone() for k, v in params.views.items()}
tr = trainer.NativeTrainer(D, init, o, cameras_extent=5.0, bg=bg, dist=dist, rank=rank, world=world)
tr.active_sh_degree = 3
B = 2 * world
tcams = [dict(syn.make_camera(W, Hh, yaw_deg=20.0 * i), image_width=W, image_height=Hh) for i in range(B)]
gg = torch.Generator().manual_seed(3)
tgts = [torch.rand(3, Hh, W, generator=gg).to(dev) for _ in range(B)]
sizes = []
for it in range(5):
    loss = tr.train_step(tcams, tgts)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss))
    sizes.append(tr.P)
    for buf in (tr.params.packed(), tr.opt.exp_avg, tr.opt.exp_avg_sq):
        ref = buf.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, buf), "trainer replicas differ at step %d" % (it + 1)
assert len(set(sizes)) > 1, sizes  # the model was rebuilt by the densification
tr.close()
# a rank that cannot set up peer buffers makes EVERY rank raise PeerUnavailable (no hang, no half-open state)
if rank == world - 1:
    os.environ["GSR_PEER_DISABLE"] = "1"
try:
    mv.PeerPacketExchange(D, dist, P, 1, rank, world, dev)
    raise AssertionError("expected PeerUnavailable")
except mv.PeerUnavailable:
    pass
os.environ.pop("GSR_PEER_DISABLE", None)
if rank == 0:
    print("multigpu ok: world=%d rel_err=%.3e" % (world, err))
dist.destroy_process_group()
