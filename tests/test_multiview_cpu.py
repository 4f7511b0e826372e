"""world_size-2 `gloo` test of the multi-view data-parallel step's host logic (view sharding, gradient all-reduce,
densification statistics) with an injected toy differentiable renderer: the sharded result must equal sequential
single-process accumulation over the same views."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _toy_render(leaves, cam):
    # smooth function of every leaf and of the camera; `viewspace_points` mimics gaussian_renderer.render()
    vp = torch.zeros_like(leaves["means3D"], requires_grad=True)
    w = (leaves["means3D"] + vp) @ cam["m"]
    img = torch.tanh(w).sum(1) * leaves["opacities"][:, 0] + leaves["scales"].prod(1)
    radii = (leaves["means3D"][:, 0] * cam["s"] > 0).to(torch.int32) * (1 + cam["k"])
    return {"render": img, "viewspace_points": vp, "radii": radii}


def _loss(out, cam, vi):
    return (out["render"] * (out["radii"] > 0)).pow(2).sum() * (1.0 + 0.1 * vi)


def _make(P=257, B=5):
    g = torch.Generator().manual_seed(0)
    leaves = {"means3D": torch.randn(P, 3, generator=g), "opacities": torch.rand(P, 1, generator=g), "scales": torch.rand(P, 3, generator=g) + 0.5}
    cams = [{"m": torch.randn(3, 3, generator=g), "s": (-1.0) ** k, "k": k} for k in range(B)]
    return leaves, cams


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, H.ROOT)
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    leaves, cams = _make()
    leaves = {k: v.clone().requires_grad_(True) for k, v in leaves.items()}
    stats = mv.DensificationStats(leaves["means3D"].shape[0], "cpu")
    total = mv.multiview_step(leaves, cams, _toy_render, _loss, rank=rank, world=world, dist=dist, stats=stats)
    # numpy arrays travel through the queue BY VALUE; torch tensors would be shared through a file-descriptor socket of this
    # process, which may already be gone when the parent reads the queue (FileNotFoundError, seen 1 run in 3)
    q.put((rank, float(total), {k: v.grad.numpy().copy() for k, v in leaves.items()}, stats.xyz_gradient_accum.numpy().copy(),
           stats.denom.numpy().copy(), stats.max_radii2D.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_multiview_step_world2_matches_sequential():
    H.pkg()
    mv = importlib.import_module(H.PKG_NAME + ".multiview")
    assert mv.shard_views(5, 0, 2) == [0, 2, 4] and mv.shard_views(5, 1, 2) == [1, 3] and mv.shard_views(1, 1, 2) == []
    # sequential reference: one process, all views
    leaves, cams = _make()
    seq = {k: v.clone().requires_grad_(True) for k, v in leaves.items()}
    sstats = mv.DensificationStats(seq["means3D"].shape[0], "cpu")
    stotal = mv.multiview_step(seq, cams, _toy_render, _loss, stats=sstats)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket

    with socket.socket() as sk:  # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, p.exitcode
    results = [(r, t, {k: torch.from_numpy(v) for k, v in g.items()}, torch.from_numpy(a), torch.from_numpy(d), torch.from_numpy(m))
               for r, t, g, a, d, m in results]
    for rank, total, grads, acc, den, mr in results:
        assert abs(total - float(stotal)) <= 1e-4 * abs(float(stotal))
        for k in grads:
            assert torch.allclose(grads[k], seq[k].grad, rtol=1e-4, atol=1e-6), k
        assert torch.allclose(acc, sstats.xyz_gradient_accum, rtol=1e-4, atol=1e-6)
        assert torch.equal(den, sstats.denom) and torch.equal(mr, sstats.max_radii2D)
    # both ranks hold bitwise-identical reduced gradients (replicas stay in lock-step)
    for k in results[0][2]:
        assert torch.equal(results[0][2][k], results[1][2][k])
