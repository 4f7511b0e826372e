// Tile binning between preprocess and compositing.
//
// Reference: duplicateWithKeys (rasterizer_impl.cu:70-111) emits R 64-bit (tile<<32 | depth) keys in Gaussian-id
// order, CUB sorts them (rasterizer_impl.cu:307-312) and identifyTileRanges (rasterizer_impl.cu:116-138) finds
// the per-tile ranges. Here:
//   1. depth_keys      : (depth bits, slot) of the V visible Gaussians, in Gaussian-id order   [V pairs]
//   2. radix sort      : stable on the 32 depth bits                                            [V pairs, gsr_scan_sort.cu]
//   3. instance_offsets: exclusive scan of tiles_touched in depth order                         [V]
//   4. emit            : one CTA per 2048 consecutive INSTANCES, one thread per instance (owner found by a warp 32-ary search
//                        + a shared-memory binary search), so a Gaussian covering thousands of tiles costs the same per
//                        instance as a small one                                                 [R pairs]
//   5. radix sort      : stable on the tile id only (<= 16 bits)                                [R pairs]
//   6. tile_ranges     : boundaries of the sorted tile ids                                      [R]
// Stability of 2 and 5 gives exactly the reference's (tile, depth bits, Gaussian id) order.
#include "gsr_common.cuh"

namespace gsr
{
namespace
{
// (1) one CTA per slot-block; the visible slots of block b are [b*256, b*256 + blk_count[b]).
__global__ void __launch_bounds__(PRE_BLOCK) depth_keys_kernel(GeomState g, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t b = blockIdx.x;
    const uint32_t cnt = g.blk_count[b];
    if (threadIdx.x >= cnt) return;
    const uint32_t slot = b * PRE_BLOCK + threadIdx.x;
    const uint32_t dst = g.blk_offset[b] + threadIdx.x;
    const float depth = g.rec[3 * (size_t)slot + 2].y;
    keys[dst] = __float_as_uint(depth);
    vals[dst] = slot;
}

// (3a) tiles_touched of the Gaussians in depth order
__global__ void __launch_bounds__(256) sorted_tiles_kernel(const ushort4* __restrict__ rect, const uint32_t* __restrict__ sorted_slots, uint32_t V,
                                                           uint32_t* __restrict__ out)
{
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s >= V) return;
    const ushort4 r = rect[sorted_slots[s]];
    out[s] = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
}

// (4) One CTA per chunk of EMIT_CHUNK consecutive INSTANCES (not Gaussians): the near-camera Gaussians of a scene cover
// thousands of tiles each, so a per-Gaussian (or per-256-Gaussian) decomposition leaves a few CTAs with nearly all the work.
// The CTA finds the owners of its first and last instance with a 32-ary warp search of the offset array, stages the owners'
// offsets / rectangles / slots in shared memory and resolves every instance with a shared-memory binary search.
constexpr int EMIT_CHUNK = 2048;

// largest j in [0, n) with a[j] <= key, for non-decreasing a with a[0] <= key; whole warp cooperates (32-ary steps)
__device__ __forceinline__ uint32_t warp_upper_owner(const uint32_t* __restrict__ a, uint32_t n, uint32_t key, uint32_t lane)
{
    uint32_t lo = 0, len = n; // answer in [lo, lo+len)
    while (len > 1) {
        const uint32_t step = (len + 31) / 32;
        const uint32_t idx = lo + lane * step;
        const bool le = idx < lo + len && a[idx] <= key;
        const uint32_t m = __ballot_sync(0xffffffffu, le);
        const uint32_t k = 31 - __clz(m); // lane 0 always satisfies a[lo] <= key
        lo = lo + k * step;
        len = min(step, n - lo);
    }
    return lo;
}

__global__ void __launch_bounds__(256) emit_kernel(const ushort4* __restrict__ rect, const uint32_t* __restrict__ sorted_slots,
                                                   const uint32_t* __restrict__ soff, uint32_t V, uint32_t R, int grid_x,
                                                   uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals)
{
    __shared__ uint32_t s_off[EMIT_CHUNK + 2];
    __shared__ ushort4 s_rect[EMIT_CHUNK + 1];
    __shared__ uint32_t s_slot[EMIT_CHUNK + 1];
    __shared__ uint32_t s_owner[2];
    const uint32_t begin = blockIdx.x * EMIT_CHUNK;
    const uint32_t end = min(R, begin + EMIT_CHUNK);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (warp < 2) {
        const uint32_t j = warp_upper_owner(soff, V, warp == 0 ? begin : end - 1, lane);
        if (lane == 0) s_owner[warp] = j;
    }
    __syncthreads();
    // every visible Gaussian owns >= 1 instance, so at most EMIT_CHUNK owners overlap the chunk
    const uint32_t j0 = s_owner[0], cnt = s_owner[1] - j0 + 1;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        const uint32_t slot = sorted_slots[j0 + i];
        s_slot[i] = slot;
        s_rect[i] = rect[slot];
        s_off[i] = soff[j0 + i];
    }
    if (threadIdx.x == 0) s_off[cnt] = soff[j0 + cnt];
    __syncthreads();
    for (uint32_t k = begin + threadIdx.x; k < end; k += 256) {
        uint32_t lo = 0, hi = cnt; // largest j in [0,cnt) with s_off[j] <= k
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= k) lo = mid;
            else hi = mid;
        }
        const ushort4 r = s_rect[lo];
        const uint32_t local = k - s_off[lo];
        const uint32_t w = (uint32_t)(r.z - r.x);
        const uint32_t ty = r.y + local / w;
        const uint32_t tx = r.x + local % w;
        out_keys[k] = ty * (uint32_t)grid_x + tx;
        out_vals[k] = s_slot[lo];
    }
}

// (6) boundaries of the sorted tile ids; ranges was zeroed, so untouched tiles stay (0,0) like the reference.
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ tkeys, uint32_t R, uint2* __restrict__ ranges)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = tkeys[i];
    if (i == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = tkeys[i - 1];
        if (cur != prev) {
            ranges[prev].y = i;
            ranges[cur].x = i;
        }
    }
    if (i == R - 1) ranges[cur].y = R;
}
} // namespace

int launch_depth_keys(const GeomState& g, cudaStream_t s)
{
    if (g.nblk == 0) return 0;
    depth_keys_kernel<<<g.nblk, PRE_BLOCK, 0, s>>>(g, g.dkeys[0], g.dvals[0]); count_launches(1);
    return 0;
}

int launch_instance_offsets(const GeomState& g, const BinState& b, uint32_t V, const uint32_t* sorted_slots, cudaStream_t s)
{
    if (V == 0) {
        GSR_CUDA(cudaMemsetAsync(b.soff, 0, sizeof(uint32_t), s));
        return 0;
    }
    sorted_tiles_kernel<<<(V + 255) / 256, 256, 0, s>>>(g.rect, sorted_slots, V, b.soff); count_launches(1);
    return exclusive_scan_u32(b.soff, b.soff, V, true, b.scan_part, s);
}

int launch_emit(const GeomState& g, const BinState& b, uint32_t V, uint32_t R, const uint32_t* sorted_slots, int grid_x,
                uint32_t* out_keys, uint32_t* out_vals, cudaStream_t s)
{
    if (V == 0 || R == 0) return 0;
    emit_kernel<<<(R + EMIT_CHUNK - 1) / EMIT_CHUNK, 256, 0, s>>>(g.rect, sorted_slots, b.soff, V, R, grid_x, out_keys, out_vals);
    count_launches(1);
    return 0;
}

int launch_tile_ranges(const uint32_t* sorted_tile_keys, uint32_t R, uint2* ranges, uint32_t T, cudaStream_t s)
{
    GSR_CUDA(cudaMemsetAsync(ranges, 0, (size_t)T * sizeof(uint2), s));
    if (R == 0) return 0;
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(sorted_tile_keys, R, ranges); count_launches(1);
    return 0;
}
} // namespace gsr
