// Tile binning between preprocess and compositing.
//
// Reference: duplicateWithKeys (rasterizer_impl.cu:70-111) emits R 64-bit (tile<<32 | depth) keys in Gaussian-id
// order, CUB sorts them (rasterizer_impl.cu:307-312) and identifyTileRanges (rasterizer_impl.cu:116-138) finds
// the per-tile ranges. Here:
//   1. depth_keys      : (depth bits, slot) of the V visible Gaussians, in Gaussian-id order, + digit totals   [V pairs]
//   2. radix sort      : stable on the 32 depth bits, one look-back kernel per pass                [V pairs, gsr_scan_sort.cu]
//   3. instance_scan   : exclusive scan of tiles_touched in depth order (one look-back kernel) + chunk owners   [V]
//   4. emit            : one CTA per 2048 consecutive INSTANCES, one thread per instance (owner range from the chunk table,
//                        then a shared-memory binary search), so a Gaussian covering thousands of tiles costs the same per
//                        instance as a small one; + digit totals of the tile sort              [R pairs]
//   5. radix sort      : stable on the tile id only (<= 16 bits), one look-back kernel per pass    [R pairs]
//   6. tile_ranges     : boundaries of the sorted tile ids                                      [R]
// 11 launches (round 1: 29). Stability of 2 and 5 gives exactly the reference's (tile, depth bits, Gaussian id) order.
#include "gsr_common.cuh"

namespace gsr
{
namespace
{
// (1) persistent warps stride over the slot-blocks; the visible slots of block b are [b*256, b*256 + blk_count[b]). Besides the
// (depth bits, slot) pairs the kernel accumulates the digit totals of all four passes of the depth sort, so each pass is one kernel.
__global__ void __launch_bounds__(PRE_BLOCK) depth_keys_kernel(GeomState g, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                               uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_h[4][256];
#pragma unroll
    for (int p = 0; p < 4; p++) s_h[p][threadIdx.x] = 0;
    __syncthreads();
    // a WARP per slot-block (~50 visible slots of 256 at 20 % visibility: two trips of 32 lanes instead of one 256-thread trip
    // with 80 % idle lanes)
    const uint32_t lane = threadIdx.x & 31u, nwarps = gridDim.x * (PRE_BLOCK / 32);
    for (uint32_t b = blockIdx.x * (PRE_BLOCK / 32) + (threadIdx.x >> 5); b < g.nblk; b += nwarps) {
        const uint32_t cnt = g.blk_count[b], off = g.blk_offset[b];
        for (uint32_t j = lane; j < cnt; j += 32) {
            const uint32_t slot = b * PRE_BLOCK + j;
            const uint32_t key = g.dkeys[1][slot]; // depth bits, written by the preprocess next to the record (4 contiguous bytes per
                                                   // slot instead of one word of every 48-byte record: 60 MB -> 5 MB of reads at cfg3)
            keys[off + j] = key;
            vals[off + j] = slot;
#pragma unroll
            for (int p = 0; p < 4; p++) atomicAdd(&s_h[p][(key >> (8 * p)) & 0xffu], 1u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const uint32_t c = s_h[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
    }
}

// (3) tiles_touched of the Gaussians in depth order -> exclusive scan, in ONE kernel: CTA tiles of 2048 items handed out by an
// atomic ticket, the running total passed from tile to tile through 64-bit look-back words (flag in the top two bits, like the
// radix passes). While it holds item i's instance range [o_i, o_i + n_i) a thread also records, for every multiple of EMIT_CHUNK
// inside it, that Gaussian i owns that instance: the emit kernel then starts from a table look-up instead of a search.
constexpr int EMIT_CHUNK = 2048;
constexpr int EMIT_PER_THREAD = EMIT_CHUNK / 256; // consecutive instances per thread of the emit kernel
constexpr unsigned long long SB_AGG = 1ull << 62, SB_INC = 2ull << 62, SB_VAL = (1ull << 62) - 1ull;

__global__ void __launch_bounds__(256) instance_scan_kernel(const ushort4* __restrict__ rect, const uint32_t* __restrict__ sorted_slots, uint32_t V,
                                                            uint32_t* __restrict__ soff, uint32_t* __restrict__ chunk_owner,
                                                            unsigned long long* status /*[tiles] + ticket at [tiles]*/, uint32_t tiles)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(status + tiles, 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_ITEMS + threadIdx.x * 8; // thread t owns 8 consecutive items
    uint32_t v[8];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = 0;
        if (base + k < V) {
            const ushort4 r = rect[sorted_slots[base + k]];
            v[k] = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
        }
        sum += v[k];
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) {
        const uint32_t c = s_warp[w];
        if (w < warp) wbase += c;
        total += c;
    }
    if (warp == 0) { // the look-back is done by a whole warp: 32 predecessors' words per round trip (one thread walking them one
                     // dependent load at a time was the longest phase of this kernel)
        volatile unsigned long long* st = status;
        if (lane == 0) st[tile] = (tile == 0 ? SB_INC : SB_AGG) | (unsigned long long)total;
        unsigned long long excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            while (true) {
                const int tt = t - (int)lane;
                unsigned long long w = 2ull << 62; // = SB_INC: entries before tile 0 are never consumed (tile 0 is INCLUSIVE)
                if (tt >= 0) w = st[tt];
                while ((w & ~SB_VAL) == 0ull) w = st[tt];          // every predecessor holds a ticket, so it publishes
                const uint32_t incs = __ballot_sync(0xffffffffu, (w & SB_INC) != 0ull);
                const uint32_t first = incs ? (uint32_t)__ffs(incs) - 1u : 31u;
                unsigned long long val = lane <= first ? (w & SB_VAL) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (incs) break;
                t -= 32;
            }
            if (lane == 0) st[tile] = SB_INC | (excl + total);
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    uint32_t o = (uint32_t)s_prefix + wbase + incl - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < V) {
            soff[base + k] = o;
            const uint32_t e = o + v[k];
            for (uint32_t c = (o + EMIT_CHUNK - 1) / EMIT_CHUNK; c * EMIT_CHUNK < e; c++) chunk_owner[c] = base + k;
            if (base + k == V - 1) soff[V] = e;
            o = e;
        }
    }
}

// (4) One CTA per chunk of EMIT_CHUNK consecutive INSTANCES (not Gaussians): the near-camera Gaussians of a scene cover
// thousands of tiles each, so a per-Gaussian (or per-256-Gaussian) decomposition leaves a few CTAs with nearly all the work.
// The owners of the chunk's first and last instance come from chunk_owner (written by the scan); the CTA stages the owners'
// offsets / rectangles / slots in shared memory and resolves every instance with a shared-memory binary search. It also
// accumulates the digit totals of the tile sort's passes, so that each of them is one kernel.
__global__ void __launch_bounds__(256) emit_kernel(const ushort4* __restrict__ rect, const uint32_t* __restrict__ sorted_slots,
                                                   const uint32_t* __restrict__ soff, const uint32_t* __restrict__ chunk_owner, uint32_t V,
                                                   uint32_t R, int grid_x, int digit_bits, int passes, uint32_t* __restrict__ hist,
                                                   uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals)
{
    __shared__ uint32_t s_off[EMIT_CHUNK + 4];
    __shared__ ushort4 s_rect[EMIT_CHUNK + 2];
    __shared__ uint32_t s_slot[EMIT_CHUNK + 2];
    __shared__ uint32_t s_h[4][256];
#pragma unroll
    for (int p = 0; p < 4; p++) s_h[p][threadIdx.x] = 0;
    const uint32_t begin = blockIdx.x * EMIT_CHUNK;
    const uint32_t end = min(R, begin + EMIT_CHUNK);
    // every visible Gaussian owns >= 1 instance, so at most EMIT_CHUNK owners overlap the chunk (+ 1: j1 may be the next chunk's first owner)
    const uint32_t j0 = chunk_owner[blockIdx.x];
    const uint32_t j1 = (blockIdx.x + 1u) * EMIT_CHUNK < R ? chunk_owner[blockIdx.x + 1] : V - 1;
    const uint32_t cnt = j1 - j0 + 1;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        const uint32_t slot = sorted_slots[j0 + i];
        s_slot[i] = slot;
        s_rect[i] = rect[slot];
        s_off[i] = soff[j0 + i];
    }
    if (threadIdx.x == 0) s_off[cnt] = soff[j0 + cnt];
    __syncthreads();
    const uint32_t mask = (1u << digit_bits) - 1u;
    // A thread resolves EMIT_PER_THREAD CONSECUTIVE instances: one binary search and one division for the first, then it walks
    // (the next instance is the next tile of the same rectangle, or the first tile of the next owner), and its keys / slots leave
    // as 16-byte stores. (One search + one runtime division per instance cost ~170 instructions per instance: ALU pipe 64 % busy.)
    const uint32_t k0 = begin + threadIdx.x * EMIT_PER_THREAD;
    if (k0 < end) {
        uint32_t lo = 0, hi = cnt; // largest j in [0,cnt) with s_off[j] <= k0
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= k0) lo = mid;
            else hi = mid;
        }
        ushort4 r = s_rect[lo];
        uint32_t next_off = s_off[lo + 1];
        uint32_t slot = s_slot[lo];
        const uint32_t local = k0 - s_off[lo];
        const uint32_t w0 = (uint32_t)(r.z - r.x);
        uint32_t ty = r.y + local / w0;
        uint32_t tx = r.x + local % w0;
        uint32_t keys[EMIT_PER_THREAD], vals[EMIT_PER_THREAD];
        uint32_t run_digit = 0xffffffffu, run_count = 0;
#pragma unroll
        for (int e = 0; e < EMIT_PER_THREAD; e++) {
            const uint32_t k = k0 + e;
            if (k < end) {
                while (k >= next_off) { // first tile of the next owner (every owner has >= 1 instance)
                    lo++;
                    r = s_rect[lo];
                    slot = s_slot[lo];
                    next_off = s_off[lo + 1];
                    tx = r.x;
                    ty = r.y;
                }
                const uint32_t key = ty * (uint32_t)grid_x + tx;
                keys[e] = key;
                vals[e] = slot;
                atomicAdd(&s_h[0][key & mask], 1u);
                // digit of the second pass (tile id / 128): a thread's consecutive tiles nearly always share it, and so do its
                // neighbours' -- counted one by one these were 32-way same-address shared-memory atomics; counted per run they are few
                if (passes > 1) {
                    const uint32_t d1 = (key >> digit_bits) & mask;
                    if (d1 != run_digit) {
                        if (run_count) atomicAdd(&s_h[1][run_digit], run_count);
                        run_digit = d1;
                        run_count = 0;
                    }
                    run_count++;
                    for (int p = 2; p < passes; p++) atomicAdd(&s_h[p][(key >> (p * digit_bits)) & mask], 1u);
                }
                if (++tx == r.z) { // next tile of the rectangle, row-major like the reference (rasterizer_impl.cu:96-108)
                    tx = r.x;
                    ty++;
                }
            } else {
                keys[e] = 0;
                vals[e] = 0;
            }
        }
        if (run_count) atomicAdd(&s_h[1][run_digit], run_count);
        if (k0 + EMIT_PER_THREAD <= end) { // out_keys / out_vals are 256-byte aligned and k0 is a multiple of EMIT_PER_THREAD
#pragma unroll
            for (int e = 0; e < EMIT_PER_THREAD; e += 4) {
                *reinterpret_cast<uint4*>(out_keys + k0 + e) = make_uint4(keys[e], keys[e + 1], keys[e + 2], keys[e + 3]);
                *reinterpret_cast<uint4*>(out_vals + k0 + e) = make_uint4(vals[e], vals[e + 1], vals[e + 2], vals[e + 3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < EMIT_PER_THREAD; e++)
                if (k0 + e < end) {
                    out_keys[k0 + e] = keys[e];
                    out_vals[k0 + e] = vals[e];
                }
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; p++) {
        const uint32_t c = s_h[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
    }
}

// (6) boundaries of the sorted tile ids, four keys per thread; ranges was zeroed, so untouched tiles stay (0,0) like the reference.
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ tkeys, uint32_t R, uint2* __restrict__ ranges)
{
    const uint32_t i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= R) return;
    uint32_t k[4];
    if (i0 + 3 < R) {
        const uint4 q = *reinterpret_cast<const uint4*>(tkeys + i0); // the key buffers are 256-byte aligned
        k[0] = q.x; k[1] = q.y; k[2] = q.z; k[3] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) k[j] = i0 + j < R ? tkeys[i0 + j] : 0u;
    }
    uint32_t prev = i0 ? tkeys[i0 - 1] : 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i = i0 + j;
        if (i >= R) break;
        const uint32_t cur = k[j];
        if (i == 0) ranges[cur].x = 0;
        else if (cur != prev) {
            ranges[prev].y = i;
            ranges[cur].x = i;
        }
        if (i == R - 1) ranges[cur].y = R;
        prev = cur;
    }
}
// (7) Optional tile schedule (GSR_TILE_ORDER=1): the compositing kernels run one CTA per tile and the hardware hands CTAs out in
// index order, so tiles can be ordered by descending list length (a counting sort on length / 16, 256 buckets, one CTA): the long
// lists start first and the tail of the grid is made of short ones. Measured at cfg3 / cfg5 on B200 it does not pay -- compositing
// forward 0.4709 vs 0.4712 ms, backward 0.856 vs 0.864 ms, cfg5 forward 1.609 vs 1.625 ms, against 0.012 ms for this kernel on the
// critical path -- so the default is the raster order (with ~27 waves of tiles the tail is already short).
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint2* __restrict__ ranges, uint32_t T, uint32_t* __restrict__ order)
{
    __shared__ uint32_t s_cnt[256];
    __shared__ uint32_t s_off[256];
    if (threadIdx.x < 256) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < T; t += 1024) {
        const uint2 r = ranges[t];
        atomicAdd(&s_cnt[255u - min((r.y - r.x) >> 4, 255u)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 256; b++) {
            s_off[b] = run;
            run += s_cnt[b];
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < T; t += 1024) {
        const uint2 r = ranges[t];
        order[atomicAdd(&s_off[255u - min((r.y - r.x) >> 4, 255u)], 1u)] = t; // any order inside a bucket will do
    }
}
} // namespace

int launch_depth_keys(const GeomState& g, uint32_t* hist, cudaStream_t s)
{
    if (g.nblk == 0) return 0;
    const uint32_t grid = g.nblk < 148u * 8u ? g.nblk : 148u * 8u;
    depth_keys_kernel<<<grid, PRE_BLOCK, 0, s>>>(g, g.dkeys[0], g.dvals[0], hist); count_launches(1);
    return 0;
}

int launch_instance_offsets(const GeomState& g, const BinState& b, uint32_t V, uint32_t R, const uint32_t* sorted_slots, cudaStream_t s)
{
    (void)R;
    if (V == 0) {
        GSR_CUDA(cudaMemsetAsync(b.soff, 0, sizeof(uint32_t), s));
        return 0;
    }
    const uint32_t tiles = (V + SCAN_ITEMS - 1) / SCAN_ITEMS;
    instance_scan_kernel<<<tiles, 256, 0, s>>>(g.rect, sorted_slots, V, b.soff, b.chunk_owner, b.scan_status, tiles); count_launches(1);
    return 0;
}

int launch_emit(const GeomState& g, const BinState& b, uint32_t V, uint32_t R, const uint32_t* sorted_slots, int grid_x, int digit_bits, int passes,
                uint32_t* hist, uint32_t* out_keys, uint32_t* out_vals, cudaStream_t s)
{
    if (V == 0 || R == 0) return 0;
    emit_kernel<<<(R + EMIT_CHUNK - 1) / EMIT_CHUNK, 256, 0, s>>>(g.rect, sorted_slots, b.soff, b.chunk_owner, V, R, grid_x, digit_bits, passes, hist,
                                                                   out_keys, out_vals);
    count_launches(1);
    return 0;
}

int launch_tile_ranges(const uint32_t* sorted_tile_keys, uint32_t R, uint2* ranges, uint32_t* tile_order, uint32_t T, cudaStream_t s)
{
    GSR_CUDA(cudaMemsetAsync(ranges, 0, (size_t)T * sizeof(uint2), s));
    if (R > 0) {
        tile_ranges_kernel<<<(R + 1023) / 1024, 256, 0, s>>>(sorted_tile_keys, R, ranges); count_launches(1);
    }
    if (tile_order) {
        tile_order_kernel<<<1, 1024, 0, s>>>(ranges, T, tile_order); count_launches(1);
    }
    return 0;
}
} // namespace gsr
