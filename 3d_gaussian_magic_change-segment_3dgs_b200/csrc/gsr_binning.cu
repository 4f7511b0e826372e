// Tile binning between preprocess and compositing.
//
// Reference: duplicateWithKeys (rasterizer_impl.cu:70-111) emits R 64-bit (tile<<32 | depth) keys in Gaussian-id
// order, CUB sorts them (rasterizer_impl.cu:307-312) and identifyTileRanges (rasterizer_impl.cu:116-138) finds
// the per-tile ranges. Here:
//   1. depth_keys      : (depth bits, slot) of the V visible Gaussians, in Gaussian-id order   [V pairs]
//   2. radix sort      : stable on the 32 depth bits                                            [V pairs, gsr_scan_sort.cu]
//   3. instance_offsets: exclusive scan of tiles_touched in depth order                         [V]
//   4. emit            : one thread per INSTANCE (binary search of the owner in shared memory), so a Gaussian
//                        covering thousands of tiles costs the same per instance as a small one   [R pairs]
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

// (4) CTA = 256 consecutive depth-sorted Gaussians; threads stride over the CTA's contiguous instance range.
__global__ void __launch_bounds__(256) emit_kernel(const ushort4* __restrict__ rect, const uint32_t* __restrict__ sorted_slots,
                                                   const uint32_t* __restrict__ soff, uint32_t V, int grid_x, uint32_t* __restrict__ out_keys,
                                                   uint32_t* __restrict__ out_vals)
{
    __shared__ uint32_t s_off[257];
    __shared__ ushort4 s_rect[256];
    __shared__ uint32_t s_slot[256];
    const uint32_t first = blockIdx.x * 256;
    const uint32_t cnt = min(256u, V - first);
    if (threadIdx.x < cnt) {
        const uint32_t slot = sorted_slots[first + threadIdx.x];
        s_slot[threadIdx.x] = slot;
        s_rect[threadIdx.x] = rect[slot];
        s_off[threadIdx.x] = soff[first + threadIdx.x];
    }
    if (threadIdx.x == 0) s_off[cnt] = soff[first + cnt];
    __syncthreads();
    const uint32_t begin = s_off[0], end = s_off[cnt];
    for (uint32_t k = begin + threadIdx.x; k < end; k += 256) {
        // largest j in [0,cnt) with s_off[j] <= k  (s_off is non-decreasing; empty owners cannot occur: tiles >= 1)
        uint32_t lo = 0, hi = cnt;
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

int launch_depth_keys(const GeomState& g, const BinState& b, cudaStream_t s)
{
    if (g.nblk == 0) return 0;
    depth_keys_kernel<<<g.nblk, PRE_BLOCK, 0, s>>>(g, b.dkeys[0], b.dvals[0]); count_launches(1);
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
    emit_kernel<<<(V + 255) / 256, 256, 0, s>>>(g.rect, sorted_slots, b.soff, V, grid_x, out_keys, out_vals); count_launches(1);
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
