"""Multi-view data-parallel step: the only way this path shards.

Gaussian parameters are REPLICATED on every rank; the B views of a step are split `views[rank::world]`; each rank
rasterizes its own cameras (forward + backward through the drop-in API, each view's loss pre-scaled by 1/B) and the
parameter gradients are summed over ranks with NCCL (`torch.distributed`, one process per GPU). Densification
statistics are reduced the same way, so every rank can apply the identical densify/prune decision:

    xyz_gradient_accum += || dL/d(mean2D)[:, :2] ||   per view, visible Gaussians only   (scene/gaussian_model.py:523-526)
    denom              += visible                                                         (same)
    max_radii2D         = max(max_radii2D, radii)      visible Gaussians only             (train.py:172)

The per-view norm is taken BEFORE summing across views, which is what sequential single-view training accumulates.
The reference itself is single-view / single-GPU; this driver is new behaviour layered on the unchanged rasterizer API.
Nothing here depends on CUDA: with the `gloo` backend and an injected render function it runs on CPU (tests/).
"""
import torch

LEAVES = ["means3D", "shs", "segments", "opacities", "scales", "rotations"]  # 3+48+2+1+3+4 = 61 floats per Gaussian


def sh_coeffs_of(leaves):
    """SH coefficients per Gaussian of a leaf dict in either layout (shs, or features_dc + features_rest)."""
    if "shs" in leaves:
        return leaves["shs"].size(1)
    return 1 + leaves["features_rest"].size(1)


def _sh_and_raw(leaves):
    """(sh tensor, raw_params dict or None) of a leaf dict: the classic layout (activated values, "shs") or the raw-parameter
    layout of optim.FlatParameters ("features_dc" + "features_rest", logits / log-scales / un-normalised quaternions)."""
    if "shs" in leaves:
        return leaves["shs"], None
    return leaves["features_dc"], {"sh_rest": leaves["features_rest"], "opacities": leaves["opacities"]}


def native_view_forward(D, leaves, rs):
    """gsr_forward for a leaf dict in either layout (see _sh_and_raw); returns D._forward_native's 9-tuple."""
    e = torch.empty(0)
    sh, raw = _sh_and_raw(leaves)
    return D._forward_native(leaves["means3D"], sh, e, leaves["segments"], leaves["opacities"], leaves["scales"], leaves["rotations"], e, rs,
                             sh_rest=raw["sh_rest"] if raw else None, raw_params=raw is not None)


def shard_views(num_views, rank, world):
    """Indices of the views this rank renders."""
    return list(range(rank, num_views, world))


class DensificationStats:
    def __init__(self, P, device):
        self.xyz_gradient_accum = torch.zeros(P, 1, device=device)
        self.denom = torch.zeros(P, 1, device=device)
        self.max_radii2D = torch.zeros(P, device=device)

    def add_view(self, viewspace_grad, radii):
        vis = radii > 0
        self.max_radii2D[vis] = torch.max(self.max_radii2D[vis], radii[vis].to(self.max_radii2D.dtype))
        self.xyz_gradient_accum[vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
        self.denom[vis] += 1

    def allreduce(self, dist, group=None):
        dist.all_reduce(self.xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)


def multiview_step(leaves, cams, render_fn, loss_fn, rank=0, world=1, dist=None, group=None, stats=None):
    """One batched step.

    leaves    dict name -> leaf tensor (requires_grad) of the replicated Gaussians (any subset of LEAVES)
    cams      list of the step's B cameras (same list on every rank)
    render_fn (leaves, cam) -> dict with at least "render", "viewspace_points", "radii" (gaussian_renderer.render()'s dict)
    loss_fn   (render dict, cam, view_index) -> scalar loss of that view
    Returns the summed loss over ALL views (all-reduced). Gradients end up in leaf.grad, identical on all ranks.
    """
    B = len(cams)
    for t in leaves.values():
        t.grad = None
    total = None
    for vi in shard_views(B, rank, world):
        out = render_fn(leaves, cams[vi])
        loss = loss_fn(out, cams[vi], vi) / B
        loss.backward()
        if stats is not None:
            vp = out["viewspace_points"]
            # the densification criterion uses the gradient of the UNscaled per-view loss
            stats.add_view(vp.grad * B if vp.grad is not None else torch.zeros_like(vp), out["radii"])
        total = loss.detach() if total is None else total + loss.detach()
    first = next(iter(leaves.values()))
    if total is None:
        total = torch.zeros((), device=first.device)
    if dist is not None and world > 1:
        for name in leaves:
            t = leaves[name]
            if t.grad is None:  # a rank with no views still takes part in the collective
                t.grad = torch.zeros_like(t)
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        if stats is not None:
            stats.allreduce(dist, group)
    return total


# ---------------------------------------------------------------------------------------------------------------------
# Flat gradient buffer + in-kernel accumulation: the rasterizer's backward adds each view's
# gradient rows straight into views of ONE flat fp32 buffer, which is then summed over ranks with a single NCCL all-reduce.
# ---------------------------------------------------------------------------------------------------------------------
class FlatGradients:
    """One flat fp32 buffer, tensor-major: means3D | shs | segments | opacities | scales | rotations (61 floats per Gaussian),
    every block starting on a 256-byte boundary (the kernels load quaternions as float4 and segments as float2, so a block
    must not start at an odd multiple of 4 bytes when P is odd). `views[name]` are contiguous views shaped like the rasterizer
    inputs; they can be installed as the `.grad` of the leaves. `offsets()` gives (start, count) of every block."""
    ALIGN = 64  # floats

    def __init__(self, P, device, sh_coeffs=16, num_class=2, split_sh=False):
        """split_sh=True: the raw-parameter layout (fused activations): features_dc | features_rest instead of shs."""
        sh = {"features_dc": (P, 1, 3), "features_rest": (P, sh_coeffs - 1, 3)} if split_sh else {"shs": (P, sh_coeffs, 3)}
        self.split_sh = split_sh
        self.shapes = {"means3D": (P, 3), **sh, "segments": (P, num_class), "opacities": (P, 1), "scales": (P, 3), "rotations": (P, 4)}
        self._offsets, off = {}, 0
        for name, shape in self.shapes.items():
            cnt = int(torch.Size(shape).numel())
            self._offsets[name] = (off, cnt)
            off = (off + cnt + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.buffer = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = {name: self.buffer[o:o + c].view(self.shapes[name]) for name, (o, c) in self._offsets.items()}

    def offsets(self):
        """dict block name -> (start, count) in floats."""
        return dict(self._offsets)

    def packed(self):
        """The blocks without the alignment gaps, concatenated (a copy; for comparisons and checkpoints)."""
        return torch.cat([v.reshape(-1) for v in self.views.values()])

    def backward_out(self, means2D_grad=None):
        """dict for `_backward_native(out=...)` (the native names differ from the leaf names for SH)."""
        v = self.views
        return {"means3D": v["means3D"], "means2D": means2D_grad, "sh": v["features_dc"] if self.split_sh else v["shs"],
                "sh_rest": v["features_rest"] if self.split_sh else None, "colors_precomp": None, "segments": v["segments"],
                "opacities": v["opacities"], "scales": v["scales"], "rotations": v["rotations"], "cov3Ds_precomp": None}

    def install(self, leaves):
        for name, t in leaves.items():
            t.grad = self.views[name]

    def allreduce(self, dist, group=None):
        dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)


def native_view_backward(D, leaves, rs, fwd, upstream, flat, first, means2D_grad=None):
    """Backward of one view into the flat buffer: overwrite (zero-fill + visible rows) for the step's first local view,
    accumulate (visible rows only) for the following ones. `D` is the drop-in diff_gaussian_rasterization module, `fwd`
    the 9-tuple returned by D._forward_native, `upstream` a dict with dL/d{color, depth, alpha, segment} (missing = zeros)."""
    R, color, depth, segment, alpha, radii, geom, binb, img = fwd
    e = torch.empty(0)
    if means2D_grad is not None and not first:
        means2D_grad.zero_()  # per-view screen-space gradient (densification statistics need each view's own norm)
    sh, raw = _sh_and_raw(leaves)
    return D._backward_native(rs, leaves["means3D"], radii, e, leaves["segments"], leaves["scales"], leaves["rotations"], e,
                              upstream.get("color"), upstream.get("segment"), upstream.get("depth"), upstream.get("alpha"), sh,
                              geom, R, binb, img, alpha, out=flat.backward_out(means2D_grad), accumulate=not first,
                              sh_rest=raw["sh_rest"] if raw else None, raw_params=raw is not None, opacities=raw["opacities"] if raw else None)


# ---------------------------------------------------------------------------------------------------------------------
# Sparse gradient exchange: each view's gradient is non-zero only on that view's visible Gaussians (~20%), and its
# 48-float SH part is rank one. Ranks all-gather 68-byte packets (id + 16 floats) instead of all-reducing 244-byte dense rows,
# and every rank rebuilds + sums the rows locally in the same (rank, view) order, so replicas stay bitwise identical.
# Traffic per rank: (N-1) x 64 B x V_visible instead of ~2 x 244 B x P.
# ---------------------------------------------------------------------------------------------------------------------
def native_view_backward_packets(D, leaves, rs, fwd, upstream, means2D_grad=None, capacity=0):
    """Backward of one view as packets. Returns (blob, count int32[1], V): blob = the view's all-gather payload (packets +
    visibility index, see D.packet_blob_views) with room for max(capacity, V) packets. Passing the exchange's sticky
    capacity (`state["cap"]` of exchange_packets) lets the blob be all-gathered in place, without a repacking copy."""
    R, color, depth, segment, alpha, radii, geom, binb, img = fwd
    # V of the forward this backward belongs to (carried by its num_rendered), not of whatever forward ran last on the thread
    nvis = int(R.num_visible) if hasattr(R, "num_visible") else int((radii > 0).sum())
    sh, raw = _sh_and_raw(leaves)
    blob, count = D._backward_packets_native(rs, leaves["means3D"], radii, leaves["segments"], leaves["scales"], leaves["rotations"],
                                             upstream.get("color"), upstream.get("segment"), upstream.get("depth"), upstream.get("alpha"),
                                             sh, geom, R, binb, img, alpha, capacity=max(nvis, int(capacity)),
                                             means2D_grad=means2D_grad, raw_params=raw)
    return blob, count, nvis


def exchange_packets(D, dist, flat, leaves, local_sets, all_campos, sh_degree, world, group=None, state=None):
    """local_sets: [(blob, count, V)] of this rank's views, in view order; all_campos[r][v]: camera centre (device [3]) of
    view v of rank r (every rank knows every camera). Fills flat.buffer with the SUM over all ranks' views: ONE all-gather of
    the view blobs, then ONE gather pass (gsr_gather_packets) that writes every dense row once -- no zero fill.

    state (optional dict, kept by the caller across steps) holds the sticky blob capacity "cap": when every view of the step
    fits and was produced with that capacity, the blobs are all-gathered as they are; otherwise they are repacked to the
    step's maximum and the capacity grows (5% slack) for the following steps."""
    device = flat.buffer.device
    nv = len(local_sets)
    M = sh_coeffs_of(leaves)
    P = leaves["means3D"].size(0)
    counts_local = torch.tensor([s[2] for s in local_sets], dtype=torch.int32, device=device)
    if world > 1:
        counts_all = torch.empty(world * nv, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(counts_all, counts_local, group=group)
        need = max(int(counts_all.max().item()), 1)  # the only host read of the exchange
    else:  # one rank (several views per step on one GPU): the counts are host values already, nothing is read back
        counts_all = counts_local
        need = max(max(int(s[2]) for s in local_sets), 1)
    if world == 1:  # local gather: one pointer per view blob, no stacking copy, no common capacity needed
        campos = torch.stack([all_campos[0][v] for v in range(nv)]).contiguous()
        cap = max(D.packet_blob_capacity(s[0], P) for s in local_sets)
        if state is not None:
            state["cap"] = max(int(state.get("cap", 0)), min(P, (int(need * 1.05) + 1023) // 1024 * 1024))
        D.gather_packets_v(leaves["means3D"], campos, sh_degree, M, [s[0].data_ptr() for s in local_sets], D.packet_index_words(P), 0, cap,
                           flat.backward_out())
        return counts_all
    sticky = int(state.get("cap", 0)) if state is not None else 0
    if sticky >= need:
        cap = sticky
    else:
        cap = need
        if state is not None:
            state["cap"] = min(P, (int(need * 1.05) + 1023) // 1024 * 1024)
    nidx = D.packet_index_words(P)
    words = D.packet_blob_words(P, cap)
    if all(s[0].numel() == words for s in local_sets):
        send = local_sets[0][0].view(1, words) if nv == 1 else torch.stack([s[0] for s in local_sets])
    else:  # repack to the common capacity
        send = torch.empty((nv, words), dtype=torch.int32, device=device)
        for v, (blob, _, n) in enumerate(local_sets):
            send[v, :nidx + n * D.PACKET_WORDS].copy_(blob[:nidx + n * D.PACKET_WORDS])
    if world > 1:
        recv = torch.empty((world * nv, words), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    else:
        recv = send
    campos = torch.stack([all_campos[r][v] for r in range(world) for v in range(nv)]).contiguous()
    D.gather_packets(leaves["means3D"], campos, sh_degree, M, recv, flat.backward_out())
    return counts_all


# ---------------------------------------------------------------------------------------------------------------------
# Peer-memory exchange (NVLink / NVSwitch): every rank owns a buffer that the other ranks have mapped (gsr_peer_alloc /
# gsr_peer_open). Two forms, one stream-ordered barrier each, no all-gather, no host read:
#   pull  (default) a rank writes its views' blobs into ITS OWN buffer and the gather kernel of every rank reads all ranks' packets
#         straight over NVLink while it sums them -- the transfer IS the kernel's loads;
#   push  every rank's buffer has a receive slot per (rank, view); as soon as a view's backward has produced its blob, the copy
#         engines write it into that slot of every peer (one side stream per peer: with several views per rank the transfer of
#         view k travels while view k + 1 renders), and the gather kernels read LOCAL memory only.
# Measured on 8 B200 (cfg4, one view per rank, 551 MB in and out per GPU and step): pull 2.0 ms for barrier + gather = 276 GB/s
# inbound, the same with and without the software-pipelined kernel (not a per-warp latency limit); push 3.3 ms for copies +
# barrier + local gather (0.64 ms) = ~205 GB/s for the peer copies, one stream or seven. NCCL's own all-reduce of the dense
# buffer moves ~420 GB/s per GPU and direction on the same box, so peer traffic, not the kernels, bounds the exchange here;
# pull stays the default (GSR_PEER_MODE=push or mode="push" selects the other; both are parity-tested on 2 GPUs).
# Blobs are double-buffered: a rank may write step s+1 while a slower rank still reads step s (the barrier of step s+1 orders
# the reuse at s+2).
# ---------------------------------------------------------------------------------------------------------------------
class PeerUnavailable(RuntimeError):
    """Raised on EVERY rank when peer-visible buffers could not be set up on some rank (no CUDA IPC / no peer access)."""


class PeerPacketExchange:
    def __init__(self, D, dist, P, views_per_rank, rank, world, device, capacity=None, group=None, mode=None):
        """Collective: all ranks construct it together. Either every rank succeeds or every rank raises PeerUnavailable (the
        outcome is agreed with an all-reduce), so callers can fall back to exchange_packets (NCCL all-gather) consistently.
        mode: "pull" (default) or "push" (see above; GSR_PEER_MODE overrides the default)."""
        import os

        self.D, self.dist, self.P, self.nv, self.rank, self.world, self.device, self.group = D, dist, P, views_per_rank, rank, world, device, group
        self.mode = mode or os.environ.get("GSR_PEER_MODE", "pull")
        if self.mode not in ("push", "pull"):
            raise ValueError("PeerPacketExchange mode must be 'push' or 'pull'")
        self.capacity = int(capacity) if capacity else P  # packets per view; P always fits
        self.index_off = 0
        self.packet_off = D.packet_index_words(P)  # 128-byte aligned
        self.blob_words = D.packet_blob_words(P, self.capacity)
        self.slots = self.nv * (world if self.mode == "push" else 1)  # push: a receive slot per (rank, view)
        # one side stream per peer, so that the copies to different peers can run on different copy engines
        self._push_stream = [torch.cuda.Stream(device=device) for _ in range(world)] if (self.mode == "push" and world > 1) else None
        self._push_done = None
        self._nvis = {}
        self.pushed_bytes = 0
        nbytes = 4 * self.blob_words * self.slots
        self.local, self.peers, handles = [], [[], []], []
        self._opened = False
        err = None
        try:
            for b in range(2):
                ptr, h = D.peer_alloc(nbytes, device)
                self.local.append(ptr)
                handles.append(h)
        except Exception as ex:  # reported collectively below
            err = ex
            handles = [bytes(64), bytes(64)]
        mine = torch.tensor(list(handles[0] + handles[1]) + [0 if err else 1], dtype=torch.uint8, device=device)
        everyone = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
        if world > 1:
            dist.all_gather_into_tensor(everyone, mine, group=group)
        else:
            everyone.copy_(mine)
        everyone = everyone.cpu().view(world, -1)
        if bool((everyone[:, -1] == 1).all()):
            try:
                for b in range(2):
                    for r in range(world):
                        h = bytes(everyone[r, 64 * b:64 * (b + 1)].tolist())
                        self.peers[b].append(self.local[b] if r == rank else D.peer_open(h, device))
            except Exception as ex:
                err = ex
        elif err is None:
            err = RuntimeError("another rank could not allocate its peer buffer")
        ok = torch.tensor([0.0 if err else 1.0], device=device)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) < 1.0:
            self._release()
            raise PeerUnavailable("peer-memory exchange unavailable on this node: %s" % (err or "failed on another rank"))
        self.parity = 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self._opened = True

    def _release(self):
        for b in range(2):
            for r, ptr in enumerate(self.peers[b]):
                if r != self.rank:
                    try:
                        self.D.peer_close(ptr, self.device)
                    except Exception:
                        pass
        self.peers = [[], []]
        for ptr in self.local:
            try:
                self.D.peer_free(ptr, self.device)
            except Exception:
                pass
        self.local = []

    def blob_ptr(self, base, v):
        return base + 4 * self.blob_words * v

    def slot(self, r, v):
        """Index of the blob of view v of rank r inside a rank's buffer."""
        return r * self.nv + v if self.mode == "push" else v

    def view_backward(self, leaves, rs, fwd, upstream, v, means2D_grad=None):
        """Backward of this rank's view v of the step, written as a blob into the current peer-visible buffer (push: and from
        there into every peer's receive slot, by the copy engines on a side stream)."""
        R, color, depth, segment, alpha, radii, geom, binb, img = fwd
        base = self.blob_ptr(self.local[self.parity], self.slot(self.rank, v))
        sh, raw = _sh_and_raw(leaves)
        _, count = self.D._backward_packets_native(rs, leaves["means3D"], radii, leaves["segments"], leaves["scales"], leaves["rotations"],
                                                   upstream.get("color"), upstream.get("segment"), upstream.get("depth"), upstream.get("alpha"),
                                                   sh, geom, R, binb, img, alpha, capacity=self.capacity, means2D_grad=means2D_grad,
                                                   raw=(base + 4 * self.packet_off, base + 4 * self.index_off), raw_params=raw)
        if self._push_stream is not None:
            nvis = int(R.num_visible) if hasattr(R, "num_visible") else self.capacity  # V is a host value of this view's forward
            self._nvis[v] = nvis
            self.push_view(v, nvis)
        return count

    def push_view(self, v, nvis):
        """Push form: copy the used prefix of this rank's blob of view v -- the visibility index, then its `nvis` packets -- into
        the same slot of every peer's current buffer, on the side stream, ordered after what the current stream has enqueued."""
        nbytes = 4 * (self.packet_off + min(int(nvis), self.capacity) * self.D.PACKET_WORDS)
        off = 4 * self.blob_words * self.slot(self.rank, v)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        if self._push_done is None:
            self._push_done = []
        for r in range(self.world):
            if r != self.rank:
                st = self._push_stream[r]
                st.wait_event(ready)
                self.D.peer_copy(self.peers[self.parity][r] + off, self.local[self.parity] + off, nbytes, self.device, st)
                ev = torch.cuda.Event()
                ev.record(st)
                self._push_done.append(ev)
        self.pushed_bytes = nbytes * (self.world - 1)

    def repush(self):
        """Re-issue the pushes of the step's views (measurement support: bench.py times push + barrier + gather in isolation)."""
        if self._push_stream is not None:
            for v, n in sorted(self._nvis.items()):
                self.push_view(v, n)

    def exchange(self, flat, leaves, all_campos, sh_degree):
        """Fills flat.buffer with the sum over all ranks' views of the step (all ranks call this after their view_backward
        calls). all_campos[r][v]: camera centre (device [3]) of view v of rank r."""
        if self._push_done:  # this rank's blobs have landed in every peer's slots
            cur = torch.cuda.current_stream(self.device)
            for ev in self._push_done:
                cur.wait_event(ev)
        self._push_done = None
        if self.world > 1:
            self.dist.all_reduce(self._flag, group=self.group)  # stream-ordered barrier: every rank's blobs are complete
        if self.mode == "push":
            ptrs = [self.blob_ptr(self.local[self.parity], self.slot(r, v)) for r in range(self.world) for v in range(self.nv)]
        else:
            ptrs = [self.blob_ptr(self.peers[self.parity][r], v) for r in range(self.world) for v in range(self.nv)]
        campos = torch.stack([all_campos[r][v] for r in range(self.world) for v in range(self.nv)]).contiguous()
        self.D.gather_packets_v(leaves["means3D"], campos, sh_degree, sh_coeffs_of(leaves), ptrs, self.packet_off, self.index_off,
                                self.capacity, flat.backward_out())
        self.parity ^= 1

    def close(self):
        """Collective. Unmap the peers' buffers, then (after a barrier) free this rank's."""
        if not getattr(self, "_opened", False):
            return
        self._opened = False
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier(group=self.group)
        for b in range(2):
            for r, ptr in enumerate(self.peers[b]):
                if r != self.rank:
                    self.D.peer_close(ptr, self.device)
        self.peers = [[], []]
        if self.world > 1:
            self.dist.barrier(group=self.group)
        for ptr in self.local:
            self.D.peer_free(ptr, self.device)
        self.local = []
