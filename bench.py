#!/usr/bin/env python
"""bench.py -- train-step it/s (rasterizer forward+backward, 1080p, 6M Gaussians, depth) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of views: for each of the rank's views, rasterize_gaussians forward
(colour + depth + alpha + segment) and backward (all parameter gradients, dense). N = 1: through the drop-in PyTorch API and
autograd. N > 1 (multi-view data parallelism; parameters replicated, one view per rank per step): the
gradients of all ranks' views are summed into one flat buffer on every rank -- `--grad-exchange peer` (default): each view's
backward writes 68-byte packets of its visible Gaussians into peer-visible memory and ONE kernel per rank pulls all ranks'
packets over NVLink while summing them (multiview.PeerPacketExchange); `packets`: the same packets through one NCCL
all-gather; `dense`: one NCCL all-reduce of the flat buffer. Workload = BASELINE.json configs[2] ("cfg3": 6M Gaussians SH3,
1920x1080, depth render + depth gradient), the configuration the headline metric is quoted on; synthetic seeded scene
(synthetic.py).

Printed JSON (rank 0, one line): `value` = views/s with everything resident in HBM (CUDA events, max over ranks);
`e2e` = the same step driven from HOST buffers: camera matrices + ground-truth image and depth copied H2D from pinned
memory every view, L1 colour + depth loss in torch, loss scalar read back D2H; `roofline` = dominant kernel against the
measured HBM peak; `cpu_baseline` = the C oracle (oracle/gsr_oracle.c, OpenMP) on a bounded sample of the same workload.

`--impl reference` times the reference's own CUDA rasterizer (oracle/_ref/ref_dgr_C.so, compiled from the unmodified
sources by oracle/build_ref.py) through the same loops -- the reference ships no CPU path, so its "own implementation of
the path" is this CUDA build; when the .so or a GPU is missing it falls back to the CPU oracle port.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "3d_gaussian_magic_change-segment_3dgs_b200"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ rasterizers
def load_ours():
    pkg = importlib.import_module(PKG)
    return pkg


class RefRasterize(torch.autograd.Function):
    """The reference's _RasterizeGaussians (diff_gaussian_rasterization/__init__.py:46-166) restated over the reference's
    OWN compiled extension (oracle/_ref/ref_dgr_C.so): same positional argument order, same saved tensors, same outputs."""

    @staticmethod
    def forward(ctx, C, means3D, means2D, sh, colors_precomp, segments, opacities, scales, rotations, cov3Ds_precomp, rs):
        args = (rs.bg, means3D, colors_precomp, segments, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        num_rendered, color, depth, segment, alpha, radii, geomBuffer, binningBuffer, imgBuffer = C.rasterize_gaussians(*args)
        ctx.C, ctx.rs, ctx.num_rendered = C, rs, num_rendered
        ctx.save_for_backward(colors_precomp, segments, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
                              alpha)
        return color, radii, depth, alpha, segment

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_segment):
        rs = ctx.rs
        colors_precomp, segments, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer, alpha = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, segments, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, grad_color, grad_segment, grad_depth, grad_alpha, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered,
                binningBuffer, imgBuffer, alpha, rs.debug)
        (g_m2d, g_col, g_op, g_m3d, g_cov, g_sh, g_sc, g_rot, g_seg) = ctx.C.rasterize_gaussians_backward(*args)
        return None, g_m3d, g_m2d, g_sh, None, g_seg, g_op, g_sc, g_rot, None, None


def make_rasterize_fn(impl, pkg):
    """Returns f(leaves, means2D, rs) -> (color, radii, depth, alpha, segment) for the chosen implementation."""
    if impl == "ours":
        def f(lv, means2D, rs):
            return pkg.GaussianRasterizer(rs)(means3D=lv["means3D"], means2D=means2D, opacities=lv["opacities"], shs=lv["shs"],
                                              segments=lv["segments"], scales=lv["scales"], rotations=lv["rotations"])
        return f
    import helpers as H

    C = H.ref_dgr()
    if C is None:
        return None
    e = torch.empty(0)

    def f(lv, means2D, rs):
        return RefRasterize.apply(C, lv["means3D"], means2D, lv["shs"], e, lv["segments"], lv["opacities"], lv["scales"], lv["rotations"], e, rs)
    return f


# ------------------------------------------------------------------------------------------------ workload
LEAVES = ["means3D", "shs", "segments", "opacities", "scales", "rotations"]  # 3+48+2+1+3+4 = 61 floats per Gaussian


def build_workload(pkg, syn, name, device, views, view_base, seed_override=None):
    P, W, Hh, seed = syn.CONFIGS[name]
    gs, _ = syn.make_scene(name)
    cams = [syn.make_camera(W, Hh, yaw_deg=45.0 * (view_base + v)) for v in range(views)]
    ug = syn.upstream_grads(W, Hh, seed, with_depth=True)
    return dict(P=P, W=W, H=Hh, seed=seed, gs={k: v.to(device) for k, v in gs.items()}, cams=cams,
                ug={k: (v.to(device) if v is not None else None) for k, v in ug.items()})


def settings_for(pkg, cam, bg, device):
    return pkg.GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg, scale_modifier=1.0,
        viewmatrix=cam["viewmatrix_dev"], projmatrix=cam["projmatrix_dev"], sh_degree=3, campos=cam["campos_dev"], prefiltered=False, debug=False)


def _claim_stdout():
    """NCCL / torchrun print banners to fd 1; the contract is ONE JSON line on stdout. Park the real stdout on a private fd
    and point fd 1 at stderr for the rest of the run."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--views-per-rank", type=int, default=1)
    ap.add_argument("--grad-exchange", default="peer", choices=["peer", "packets", "dense"],
                    help="N>1: peer = gather kernel pulls every rank's 68-B gradient packets over NVLink peer memory (default); "
                         "packets = NCCL all-gather of the packets, then the gather kernel; dense = all-reduce of the flat buffer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-profile", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # the first ~8 steps grow the caching allocator's pools (cudaMalloc of the GB-sized state/gradient buffers): always run
    # at least that many untimed steps before the W warm-ups the caller asked for are considered done
    args.extra_warmup = max(0, 8 - args.warmup) if os.environ.get("GSR_BENCH_MIN_WARMUP", "1") == "1" else 0

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    have_gpu = torch.cuda.is_available()

    if args.impl == "reference" and world > 1 and rank != 0:
        return 0  # the reference is single-GPU: rank 0 alone runs and prints it

    syn = importlib.import_module(PKG + ".synthetic")
    if args.impl == "reference" and (not have_gpu or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_dgr_C.so"))):
        return reference_cpu_port(args, syn, out)

    if not have_gpu:
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    # clocks are sampled for the whole run; starting nvidia-smi here keeps its (slow, driver-locking) start-up out of the
    # timed regions
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("GSR_BENCH_NO_SAMPLER"):
        sampler.start()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pkg = load_ours()  # also provides GaussianRasterizationSettings for the reference arm
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=device)
    nranks = world if args.impl == "ours" else 1

    rasterize = make_rasterize_fn(args.impl, pkg)
    if rasterize is None:
        return reference_cpu_port(args, syn, out)

    V = args.views_per_rank
    wl = build_workload(pkg, syn, args.workload, device, V, rank * V)
    P, W, Hh = wl["P"], wl["W"], wl["H"]
    N = W * Hh
    bg = torch.zeros(3, device=device)
    for cam in wl["cams"]:
        for k in ["viewmatrix", "projmatrix", "campos"]:
            cam[k + "_pin"] = cam[k].clone().pin_memory()
            cam[k + "_dev"] = cam[k].to(device)
    leaves = {k: wl["gs"][k].clone().requires_grad_(True) for k in LEAVES}
    ug = wl["ug"]

    # host-side "dataset" for the e2e leg: ground-truth image and (inverse, normalised) depth per view, pinned
    g = torch.Generator().manual_seed(77 + rank)
    gt_img = [torch.rand(3, Hh, W, generator=g).pin_memory() for _ in range(V)]
    gt_dep = [torch.rand(1, Hh, W, generator=g).pin_memory() for _ in range(V)]
    h2d_bytes = V * (gt_img[0].numel() + gt_dep[0].numel() + 16 + 16 + 3) * 4
    d2h_bytes = V * 4

    def zero_grads():
        for v in leaves.values():
            v.grad = None

    e2e_sets = []
    xstate = {}  # sticky blob capacity of the packet exchange

    def allreduce_grads():
        if dist is None:
            return
        if use_peer:
            px.exchange(flat, leaves, all_campos, 3)
            return
        if use_packets and e2e_sets:
            mv.exchange_packets(Dmod, dist, flat, leaves, e2e_sets, all_campos, 3, nranks, state=xstate)
            e2e_sets.clear()
            return
        if flat is not None:
            flat.allreduce(dist)
            return
        for k in LEAVES:
            dist.all_reduce(leaves[k].grad)

    mv = importlib.import_module(PKG + ".multiview")
    use_flat = args.impl == "ours" and (nranks > 1 or V > 1)
    flat = mv.FlatGradients(P, device) if use_flat else None
    Dmod = pkg.diff_gaussian_rasterization
    empty = torch.empty(0)

    use_packets = use_flat and dist is not None and args.grad_exchange == "packets"
    use_peer = use_flat and dist is not None and args.grad_exchange == "peer"
    px = None
    if use_peer:
        try:
            px = mv.PeerPacketExchange(Dmod, dist, P, V, rank, nranks, device)
        except mv.PeerUnavailable as ex:  # raised on every rank together: fall back to the NCCL all-gather of the same packets
            if rank == 0:
                print("bench: %s -- falling back to --grad-exchange packets" % ex, file=sys.stderr, flush=True)
            use_peer, use_packets = False, True
    all_campos = None
    if use_packets or use_peer:  # every rank knows every camera of the step
        all_campos = [[syn.make_camera(W, Hh, yaw_deg=45.0 * (r * V + v))["campos"].to(device) for v in range(V)] for r in range(nranks)]

    def step_device_flat():
        """multi-view / multi-GPU step. dense: every view's backward adds into ONE flat gradient buffer, one NCCL all-reduce.
        packets: every view's backward emits 68-B packets of its visible Gaussians + an id map, NCCL all-gathers, then one
        gather pass that sums all views and writes every dense row once."""
        sets = []
        for v in range(V):
            rs = settings_for(pkg, wl["cams"][v], bg, device)
            with torch.no_grad():
                fwd = Dmod._forward_native(leaves["means3D"], leaves["shs"], empty, leaves["segments"], leaves["opacities"], leaves["scales"],
                                           leaves["rotations"], empty, rs)
                if use_peer:
                    px.view_backward(leaves, rs, fwd, ug, v)
                elif use_packets:
                    sets.append(mv.native_view_backward_packets(Dmod, leaves, rs, fwd, ug, capacity=xstate.get("cap", 0)))
                else:
                    mv.native_view_backward(Dmod, leaves, rs, fwd, ug, flat, first=(v == 0))
        if use_peer:
            px.exchange(flat, leaves, all_campos, 3)
        elif use_packets:
            mv.exchange_packets(Dmod, dist, flat, leaves, sets, all_campos, 3, nranks, state=xstate)
        elif dist is not None:
            flat.allreduce(dist)

    def step_device():
        if use_flat:
            return step_device_flat()
        zero_grads()
        for v in range(V):
            rs = settings_for(pkg, wl["cams"][v], bg, device)
            means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
            color, radii, depth, alpha, segment = rasterize(leaves, means2D, rs)
            torch.autograd.backward([color, depth], [ug["color"], ug["depth"]])
        allreduce_grads()

    loss_host = [0.0]
    copy_stream = torch.cuda.Stream(device=device)

    def stage_view(v):
        """H2D of view v's inputs (camera + ground truth) from pinned memory on the copy stream, double-buffered like a
        prefetching data loader; the compute stream waits on the event before using them."""
        with torch.cuda.stream(copy_stream):
            cam = wl["cams"][v]
            mats = [cam[k + "_pin"].to(device, non_blocking=True) for k in ("viewmatrix", "projmatrix", "campos")]
            gi = gt_img[v].to(device, non_blocking=True)
            gd = gt_dep[v].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return mats, gi, gd, ev

    pending = {}

    def step_e2e():
        zero_grads()
        total = None
        for v in range(V):
            cam = wl["cams"][v]
            if v not in pending:
                pending[v] = stage_view(v)
            mats, gi, gd, ev = pending.pop(v)
            torch.cuda.current_stream().wait_event(ev)
            for t in mats + [gi, gd]:
                t.record_stream(torch.cuda.current_stream())
            nv = (v + 1) % V
            pending[nv] = stage_view(nv)  # next view's (next step's) inputs travel while this view computes
            rs = settings_for(pkg, dict(cam, viewmatrix_dev=mats[0], projmatrix_dev=mats[1], campos_dev=mats[2]), bg, device)
            if use_flat:
                with torch.no_grad():
                    fwd = Dmod._forward_native(leaves["means3D"], leaves["shs"], empty, leaves["segments"], leaves["opacities"],
                                               leaves["scales"], leaves["rotations"], empty, rs)
                color, depth = fwd[1].requires_grad_(True), fwd[2].requires_grad_(True)
                dn = depth / (depth.max() + 1e-5)
                loss = (color - gi).abs().mean() + 0.1 * (dn - gd).abs().mean()
                loss.backward()  # pixel gradients only; the rasterizer backward runs natively into the flat buffer
                with torch.no_grad():
                    pg = {"color": color.grad, "depth": depth.grad}
                    if use_peer:
                        px.view_backward(leaves, rs, fwd, pg, v)
                    elif use_packets:
                        e2e_sets.append(mv.native_view_backward_packets(Dmod, leaves, rs, fwd, pg, capacity=xstate.get("cap", 0)))
                    else:
                        mv.native_view_backward(Dmod, leaves, rs, fwd, pg, flat, first=(v == 0))
            else:
                means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
                color, radii, depth, alpha, segment = rasterize(leaves, means2D, rs)
                dn = depth / (depth.max() + 1e-5)  # gaussian_renderer/__init__.py:375
                loss = (color - gi).abs().mean() + 0.1 * (dn - gd).abs().mean()  # train.py:111-121 shape (L1 + depth term)
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        allreduce_grads()
        loss_host[0] = float(total.item())  # D2H read of the step's result

    step_stats = {}

    def timed(fn, steps, warmup):
        import gc

        for _ in range(warmup + args.extra_warmup):
            fn()
        # a full (generation-2) Python GC pass walks every object torch/numpy created at import time (30-40 ms, seen as a
        # single slow step); collect now and freeze the survivors so no such pass can land inside the timed region
        gc.collect()
        gc.freeze()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ms0 = torch.cuda.memory_stats(device)
        host_t = []
        t0 = time.time()
        evs[0].record()
        for i in range(steps):
            h0 = time.perf_counter()
            fn()
            host_t.append((time.perf_counter() - h0) * 1e3)
            evs[i + 1].record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms1 = torch.cuda.memory_stats(device)
        ms = evs[0].elapsed_time(evs[steps])  # the K steps, bracketed
        if os.environ.get("GSR_BENCH_DEBUG_STEPS"):
            print("STEPS", fn.__name__, [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(steps)], file=sys.stderr, flush=True)
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        step_stats[fn.__name__] = {"median_ms": round(per[len(per) // 2], 4), "min_ms": round(per[0], 4), "max_ms": round(per[-1], 4),
                                   "host_max_ms": round(max(host_t), 3), "host_median_ms": round(sorted(host_t)[len(host_t) // 2], 3),
                                   "cudaMalloc_calls": int(ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0)),
                                   "cudaFree_calls": int(ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0))}
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, t0, t1

    if os.environ.get("GSR_BENCH_DEBUG") and args.impl == "ours":
        Ld = pkg._lib.lib()
        Ld.gsr_set_profiling(1)
        for name, fn in [("device", step_device), ("e2e", step_e2e), ("device", step_device), ("e2e", step_e2e)]:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.time()
            fn()
            torch.cuda.synchronize()
            print("DEBUG", name, "wall_ms=%.3f" % ((time.time() - t0) * 1e3), pkg._lib.stage_times(), file=sys.stderr, flush=True)
        Ld.gsr_set_profiling(0)

    L = pkg._lib.lib()
    launches0 = int(L.gsr_launch_count())
    ms_dev, t0, t1 = timed(step_device, args.steps, args.warmup)
    launches_timed = (int(L.gsr_launch_count()) - launches0) * args.steps // (args.steps + args.warmup + args.extra_warmup)
    clocks = sampler.summary(t0, t1) if rank == 0 else None
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    if rank == 0:
        sampler.stop()

    comm_ms = None
    if dist is not None:
        def comm_only():
            if use_peer:
                px.exchange(flat, leaves, all_campos, 3)
            elif use_packets:
                mv.exchange_packets(Dmod, dist, flat, leaves, comm_sets, all_campos, 3, nranks, state=xstate)
            else:
                flat.allreduce(dist)
        comm_sets = []
        if use_packets:
            with torch.no_grad():
                for v in range(V):
                    rs = settings_for(pkg, wl["cams"][v], bg, device)
                    fwd = Dmod._forward_native(leaves["means3D"], leaves["shs"], empty, leaves["segments"], leaves["opacities"],
                                               leaves["scales"], leaves["rotations"], empty, rs)
                    comm_sets.append(mv.native_view_backward_packets(Dmod, leaves, rs, fwd, ug, capacity=xstate.get("cap", 0)))
        for _ in range(2):
            comm_only()
        torch.cuda.synchronize()
        dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            comm_only()
        c1.record()
        torch.cuda.synchronize()
        comm_ms = c0.elapsed_time(c1) / 5

    # forward-only ms/frame (the second half of BASELINE.json's metric), same workload, rank-local
    def fwd_only():
        with torch.no_grad():
            rs = settings_for(pkg, wl["cams"][0], bg, device)
            means2D = torch.zeros_like(leaves["means3D"])
            return rasterize(leaves, means2D, rs)

    for _ in range(5):
        fwd_only()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    nf = max(10, min(50, args.steps))
    for _ in range(nf):
        fwd_only()
    f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / nf

    views_total = V * nranks * args.steps
    value = views_total / (ms_dev / 1e3)
    e2e_value = views_total / (ms_e2e / 1e3)

    # ---- realised workload statistics (one forward, untimed) ----
    with torch.no_grad():
        rs = settings_for(pkg, wl["cams"][0], bg, device)
        D = pkg.diff_gaussian_rasterization
        e = torch.empty(0)
        R, color, depth, segment, alpha, radii, geom, binb, img = D._forward_native(leaves["means3D"], leaves["shs"], e, leaves["segments"],
                                                                                   leaves["opacities"], leaves["scales"], leaves["rotations"], e, rs)
        Vn = int((radii > 0).sum())
        Pn = int(pkg.mark_visible(leaves["means3D"], rs.viewmatrix, rs.projmatrix).sum())
        T = ((W + 15) // 16) * ((Hh + 15) // 16)
        del color, depth, segment, alpha, geom, binb, img
    stats = dict(P=P, Pn=Pn, V=Vn, R=int(R), N=N, T=T)
    # algorithmic bytes
    B = {
        "preprocess_fwd": 20 * P + 52 * Pn + 239 * Vn,
        "binning": 8 * P + (8 * P + 12 * Vn + 12 * R) + 24 * R + (8 * R + 8 * T),
        "render_fwd": 28 * R + 24 * Vn + 32 * N,
        "render_bwd": 52 * R + 36 * N + 48 * Vn,
        "preprocess_bwd": (56 + 36) * Vn + (303 + 232) * Vn + 312 * (P - Vn),
    }
    bytes_step = sum(B.values())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    roofline = None
    stages = None
    if args.impl == "ours" and not args.no_stage_profile:  # every rank runs it (ranks stay in step); rank 0 reports
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
        L.gsr_set_profiling(1)
        acc = {}
        nprof = max(3, min(10, args.steps))
        for _ in range(nprof):
            zero_grads()
            rs = settings_for(pkg, wl["cams"][0], bg, device)
            means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
            color, radii, depth, alpha, segment = rasterize(leaves, means2D, rs)
            torch.autograd.backward([color, depth], [ug["color"], ug["depth"]])
            torch.cuda.synchronize()
            for k, v in pkg._lib.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / nprof
        L.gsr_set_profiling(0)
        if os.environ.get("GSR_BENCH_DEBUG"):
            print("DEBUG rank", rank, "stages", {k: round(v, 4) for k, v in acc.items()}, file=sys.stderr, flush=True)
        group = {"preprocess_fwd": ["preprocess_fwd"], "binning": ["depth_sort", "emit", "tile_sort", "tile_ranges"], "render_fwd": ["render_fwd"],
                 "render_bwd": ["render_bwd"], "preprocess_bwd": ["preprocess_bwd"]}
        stages = {}
        for gname, members in group.items():
            ms = sum(acc.get(m, 0.0) for m in members)
            stages[gname] = {"ms": round(ms, 4), "alg_bytes": B[gname], "GBps": round(B[gname] / (ms * 1e-3) / 1e9, 1) if ms > 0 else None,
                             "frac_hbm": round(B[gname] / (ms * 1e-3) / 1e9 / hbm_peak, 4) if ms > 0 else None}
        dom = max(stages, key=lambda k: stages[k]["ms"])
        ach = stages[dom]["GBps"]
        traffic = None  # dram bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
        try:
            kern = json.load(open(os.path.join(ROOT, "profiles", "r01_kernels.json")))
            for kname, kv in kern.items():
                if dom in kname and "_kernel" in kname:  # render_bwd -> render_bwd2_kernel<2> / render_bwd_kernel<2>
                    traffic = kv.get("dram_traffic_bytes")
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                    "traffic": traffic, "peak_source": peak_src, "launch_ms": stages[dom]["ms"], "alg_bytes_per_launch": B[dom],
                    "note": "compositing is FP32-issue / shared-memory / atomic bound, not HBM bound (no stage is a dense contraction); "
                            "the HBM fraction is reported because the contract asks for it, see DESIGN.md"}
    elif args.impl == "reference":
        ach = bytes_step / (ms_dev / args.steps / V * 1e-3) / 1e9
        roofline = {"kernel": "whole step (reference CUDA)", "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(ach / hbm_peak, 4), "traffic": None, "peak_source": peak_src}

    if px is not None:
        px.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if not args.no_cpu_baseline and nranks == 1:
        cpu_baseline = run_cpu_baseline(syn, args.workload)
    if args.impl == "reference":
        cpu_baseline = {"value": round(value, 4), "unit": "it/s", "cores": 0, "kind": "reference",
                        "sample": "full %s on the GPU: the reference's own implementation of this path is CUDA (oracle/_ref/ref_dgr_C.so), "
                                  "it ships no CPU path" % args.workload}

    line = {
        "metric": "train-step it/s (fwd+bwd, 1080p, 6M gaussians)" if args.workload in ("cfg3", "cfg4") else "train-step it/s (fwd+bwd)",
        "value": round(value, 4), "unit": "it/s", "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %d Gaussians SH3, %dx%d, rasterize_gaussians fwd (colour+depth+alpha+segment) + bwd (dL/dcolour, dL/ddepth), "
                               "%d view(s)/rank/step%s" % (args.workload, P, W, Hh, V,
                                                           (", gradient exchange = peer memory: every rank's gather kernel pulls all ranks' 68-B packets of the visible Gaussians over "
                                                            "NVLink while summing them into the flat buffer (one stream-ordered barrier, no all-gather)" if use_peer else
                                                            ", gradient exchange = one NCCL all-gather of per-view blobs (68-B packets of the visible Gaussians + "
                                                            "visibility index), then one gather pass into the flat buffer" if use_packets else
                                                            ", gradients accumulated in one flat buffer (61 floats/Gaussian), one NCCL all-reduce")
                                                           if nranks > 1 else ""),
                   "views_per_rank": V, "l2": "inputs (%.2f GB of parameters) are larger than the 126 MB L2" % (61 * 4 * P / 1e9), "stats": stats,
                   "alg_bytes_per_step": bytes_step},
        "e2e": {"value": round(e2e_value, 4), "unit": "it/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": round(ms_e2e / args.steps, 4), "loss": loss_host[0]},
        "gpu_launches": launches_timed if args.impl == "ours" else 0,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_step": {"alg_bytes": bytes_step, "achieved": round(bytes_step / (ms_dev / args.steps / V * 1e-3) / 1e9, 1), "peak": hbm_peak,
                          "unit": "GB/s", "frac": round(bytes_step / (ms_dev / args.steps / V * 1e-3) / 1e9 / hbm_peak, 4)},
        "cpu_baseline": cpu_baseline,
        "step_ms": step_stats,
        "fwd_ms_per_frame": round(fwd_ms, 4),
    }
    if comm_ms is not None:
        gbytes = 61 * 4 * P / 1e9
        if use_peer:
            line["collective"] = {"op": "4-byte ncclAllReduce as stream-ordered barrier + ONE gather kernel over %d views reading peer blobs "
                                        "over NVLink (68 B per visible Gaussian + 2 index words per 32 Gaussians per view)" % (nranks * V),
                                  "bytes_pulled_per_rank": int((68 * stats["V"] + P // 4) * V * (nranks - 1)), "ms": round(comm_ms, 3),
                                  "dense_allreduce_bytes": int(61 * 4 * P)}
        elif use_packets:
            line["collective"] = {"op": "count all-gather + ONE ncclAllGather of view blobs (68 B per visible Gaussian + 2 index words per 32 "
                                        "Gaussians) + ONE gather pass over %d views that writes every dense row once" % (nranks * V),
                                  "bytes_sent_per_rank": int((68 * xstate.get("cap", stats["V"]) + P // 4) * V), "ms": round(comm_ms, 3),
                                  "dense_allreduce_bytes": int(61 * 4 * P)}
        else:
            line["collective"] = {"op": "1 x ncclAllReduce(sum, fp32) of the flat gradient buffer", "bytes": int(61 * 4 * P),
                                  "ms": round(comm_ms, 3), "algbw_GBps": round(gbytes / (comm_ms * 1e-3), 1),
                                  "busbw_GBps": round(gbytes / (comm_ms * 1e-3) * 2 * (nranks - 1) / nranks, 1)}
    if stages is not None:
        line["stages"] = stages
    if args.impl == "reference":
        line["impl"] = "reference"
    out.write(json.dumps(line) + "\n")
    out.flush()
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ CPU legs
def run_cpu_baseline(syn, workload, row_stride=1):
    """Oracle B (C + OpenMP, all host cores) on a bounded sample of the SAME workload: full per-Gaussian stages and binning,
    compositing forward+backward on every `row_stride`-th tile row; the compositing time is scaled by the sampled
    fraction of tile instances. Reported baseline, not a target."""
    from oracle import cpu_oracle as O

    P, W, Hh, seed = syn.CONFIGS[workload]
    gs, cam = syn.make_scene(workload)
    ug = syn.upstream_grads(W, Hh, seed, with_depth=True)
    n = lambda t: t.numpy()
    O.lib()
    gy = (Hh + 15) // 16
    row_stride = min(row_stride, gy)
    reps, t_f, t_b = 0, 0.0, 0.0
    while reps < 1 or (t_f + t_b < 10.0 and reps < 5):  # about 10 s of CPU work, averaged
        t0 = time.time()
        st = O.forward(n(gs["means3D"]), n(gs["opacities"]), W, Hh, cam["tanfovx"], cam["tanfovy"], n(cam["viewmatrix"]), n(cam["projmatrix"]),
                       n(cam["campos"]), np.zeros(3, np.float32), shs=n(gs["shs"]), segments=n(gs["segments"]), scales=n(gs["scales"]),
                       rotations=n(gs["rotations"]), row_stride=row_stride, row_offset=row_stride // 2)
        t1 = time.time()
        O.backward(st, n(ug["color"]), n(ug["depth"]))
        t_f += t1 - t0
        t_b += time.time() - t1
        reps += 1
    t0, t1, t2 = 0.0, t_f / reps, (t_f + t_b) / reps
    total_cpu_s = t_f + t_b
    # fraction of tile instances in the sampled rows
    gx = (W + 15) // 16
    rng = st["ranges"].astype(np.int64)
    lens = (rng[:, 1] - rng[:, 0]).reshape(gy, gx).sum(1)
    frac = float(lens[row_stride // 2::row_stride].sum()) / max(1.0, float(lens.sum()))
    # split: per-Gaussian + binning parts run in full, compositing parts are sampled. Time them separately by a second,
    # compositing-free estimate: t_full_parts = total - t_sampled_compositing is not separable without extra timers, so
    # re-run the sampled compositing alone to measure it.
    L = O.lib()
    tc0 = time.time()
    nc = np.zeros(W * Hh, np.uint32)
    col, seg, dep, alp = np.zeros((3, Hh, W), np.float32), np.zeros((2, Hh, W), np.float32), np.zeros((1, Hh, W), np.float32), np.zeros((1, Hh, W), np.float32)
    L.orc_render_forward(O._i(W), O._i(Hh), O._i(2), O._p(st["ranges"]), O._p(st["point_list"]), O._p(st["means2D"]), O._p(st["rgb"]),
                         O._p(st["_inputs"]["segments"]), O._p(st["depths"]), O._p(st["conic_opacity"]), O._p(np.zeros(3, np.float32)), O._p(col),
                         O._p(seg), O._p(dep), O._p(alp), O._p(nc), O._i(row_stride), O._i(row_stride // 2))
    tc1 = time.time()
    t_render_fwd_s = tc1 - tc0
    t_fwd_full_parts = max(0.0, (t1 - t0) - t_render_fwd_s)
    # backward: compositing dominates the sampled part; preprocess backward runs in full. Estimate the split with the same ratio
    # measured on the forward would be wrong, so time preprocess_backward-only by calling backward on a state with empty ranges.
    st_empty = dict(st)
    st_empty["ranges"] = np.zeros_like(st["ranges"])
    tb0 = time.time()
    O.backward(st_empty, n(ug["color"]), n(ug["depth"]))
    tb1 = time.time()
    t_bwd_full_parts = tb1 - tb0
    t_render_bwd_s = max(0.0, (t2 - t1) - t_bwd_full_parts)
    est = t_fwd_full_parts + t_bwd_full_parts + (t_render_fwd_s + t_render_bwd_s) / max(frac, 1e-9)
    return {"value": round(1.0 / est, 5), "unit": "it/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%s scene; per-Gaussian stages, key emit, sort and ranges in full; compositing fwd+bwd on every %dth tile row "
                      "(%.1f%% of tile instances), scaled; %d repetition(s), %.1f s of CPU work measured, %.1f s per step"
                      % (workload, row_stride, 100 * frac, reps, total_cpu_s, est),
            "measured_s": round(total_cpu_s, 2)}


def reference_cpu_port(args, syn, out):
    """Fallback of `--impl reference` when the reference CUDA build (oracle/_ref) or a GPU is unavailable: the C port."""
    cb = run_cpu_baseline(syn, args.workload)
    line = {"metric": "train-step it/s (fwd+bwd, 1080p, 6M gaussians)", "value": cb["value"], "unit": "it/s", "n_gpus": 0, "steps": 1, "warmup": 0,
            "ms_per_step": round(1e3 / cb["value"], 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload}, "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
