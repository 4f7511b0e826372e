"""torchrun --nproc-per-node 2 scripts/time_peer_gather.py : cost of the gather kernel when its 8 view blobs are all LOCAL, all
in the PEER GPU's memory (NVLink loads), or half and half. cfg3, 4 views per rank."""
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
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
Pk = H.pkg()
mv = importlib.import_module(H.PKG_NAME + ".multiview")
D = Pk.diff_gaussian_rasterization
syn = H.synthetic()
P, W, Hh, seed = syn.CONFIGS["cfg3"]
gs, _ = syn.make_scene("cfg3")
gs = {k: v.to(dev) for k, v in gs.items()}
ug = {k: (v.to(dev) if v is not None else None) for k, v in syn.upstream_grads(W, Hh, seed, with_depth=True).items()}
bg = torch.tensor([0.0, 0.0, 0.0])
e = torch.empty(0)
NV = 4
px = mv.PeerPacketExchange(D, dist, P, NV, rank, world, dev, mode="pull")  # this script times the kernel's own NVLink loads
campos = [[syn.make_camera(W, Hh, yaw_deg=45.0 * (r * NV + v))["campos"].to(dev) for v in range(NV)] for r in range(world)]
for v in range(NV):
    cam = syn.make_camera(W, Hh, yaw_deg=45.0 * (rank * NV + v))
    rs = H.settings(cam, bg, device=dev)
    fwd = D._forward_native(gs["means3D"], gs["shs"], e, gs["segments"], gs["opacities"], gs["scales"], gs["rotations"], e, rs)
    px.view_backward(gs, rs, fwd, ug, v)
    del fwd
torch.cuda.synchronize()
dist.barrier()
flat = mv.FlatGradients(P, dev)
local = [px.blob_ptr(px.peers[0][rank], v) for v in range(NV)]
remote = [px.blob_ptr(px.peers[0][1 - rank], v) for v in range(NV)]
cp_local = torch.stack(campos[rank] + campos[rank])
cp_remote = torch.stack(campos[1 - rank] + campos[1 - rank])
cp_mix = torch.stack(campos[rank] + campos[1 - rank])


def timeit(ptrs, cp, n=10):
    f = lambda: D.gather_packets_v(gs["means3D"], cp, 3, 16, ptrs, px.packet_off, px.index_off, px.capacity, flat.backward_out())
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_l = timeit(local + local, cp_local)
t_r = timeit(remote + remote, cp_remote)
t_m = timeit(local + remote, cp_mix)
t_r4 = timeit(remote, torch.stack(campos[1 - rank]))
t_l4 = timeit(local, torch.stack(campos[rank]))
print("rank %d: 8 views local %.3f ms | 8 views remote %.3f ms | 4 local + 4 remote %.3f ms | 4 local %.3f | 4 remote %.3f" % (rank, t_l, t_r, t_m, t_l4, t_r4),
      flush=True)
px.close()
dist.destroy_process_group()
