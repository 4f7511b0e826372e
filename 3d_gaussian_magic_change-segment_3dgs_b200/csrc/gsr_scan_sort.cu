// Device-wide primitives written for this pipeline: an exclusive u32 scan and a stable LSD radix sort of
// (u32 key, u32 value) pairs. They replace cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:281) and
// cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:307-312; simple_knn.cu:213) of the reference.
//
// Sort design: the reference sorts R 64-bit (tile|depth) keys in 6 passes. Here the visible Gaussians (V << R)
// are depth-sorted once on their 32 depth bits, instances are emitted in that order, and only the tile id
// (<= 16 bits -> 2 passes) is sorted at instance granularity. Both sorts are stable, so the final order equals
// the reference's (tile, depth, Gaussian id) order bit for bit.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "gsr_common.cuh"

namespace cg = cooperative_groups;

namespace gsr
{
namespace
{
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = SCAN_ITEMS / SCAN_THREADS; // 8

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += n;
    }
    return v;
}

// Lanes of the warp whose (<= 8-bit) digit equals this lane's, by eight ballots -- one per digit bit -- instead of match.any: in the
// ncu source view of the look-back pass, match.any and the instruction consuming its result held 52 % of all stall samples (a long,
// unpipelined latency paid 16 times per thread and tile); the ballots are independent of one another and pipeline.
__device__ __forceinline__ uint32_t same_digit_lanes(const uint32_t d, const bool valid)
{
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
        const bool one = (d >> bit) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, one);
        peers &= one ? b : ~b;
    }
    return peers;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) {
        const uint32_t c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ partials)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const uint32_t i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) sum += in[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_warp[w];
        partials[blockIdx.x] = t;
    }
}

// Each CTA sums the partials of the CTAs before it (a few hundred values), then scans its own 2048 items.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* in, uint32_t n, const uint32_t* __restrict__ partials,
                                                                  uint32_t* out, int write_total)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_prefix;
    uint32_t pre = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += SCAN_THREADS) pre += partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = pre;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_warp[w];
        s_prefix = t;
    }
    __syncthreads();
    const uint32_t prefix = s_prefix;

    // thread t owns items [t*8, t*8+8) of the CTA tile
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        sum += v[k];
    }
    uint32_t total;
    uint32_t excl = block_excl_scan_256(sum, s_warp, total) + prefix;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    // the thread that owns item n-1 also writes out[n] = grand total
    if (write_total && n > 0 && base <= n - 1 && n - 1 < base + SCAN_PER_THREAD) out[n] = excl;
}

// ------------------------------------------------------------------------------------------ radix sort
constexpr int RADIX_THREADS = 256;
constexpr int RADIX_PER_THREAD = RADIX_ITEMS / RADIX_THREADS; // 8
constexpr int RADIX_WARP_ITEMS = RADIX_ITEMS / 8;             // 256 consecutive keys per warp

// Per-tile digit histogram, written digit-major: hist[d * ntiles + tile]. With rowsum != nullptr the tile's counts are also
// added to the per-digit totals (used by the fused sort to skip a separate reduction pass).
__device__ __forceinline__ void tile_hist_body(uint32_t tile, const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t mask,
                                               uint32_t ntiles, uint32_t* __restrict__ hist, uint32_t* rowsum)
{
    __shared__ uint32_t s_hist[256];
    __syncthreads(); // previous use of s_hist (tile loop of the fused sort)
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * RADIX_ITEMS;
#pragma unroll
    for (int k = 0; k < RADIX_PER_THREAD; k++) {
        const uint32_t i = base + k * RADIX_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= mask) {
        const uint32_t c = s_hist[threadIdx.x];
        hist[threadIdx.x * ntiles + tile] = c;
        if (rowsum && c) atomicAdd(&rowsum[threadIdx.x], c);
    }
}

// n_dev != nullptr: the element count lives in device memory (grid sized for a capacity, extra CTAs see no keys).
__global__ void __launch_bounds__(RADIX_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ n_dev,
                                                                   int shift, uint32_t mask, uint32_t nblocks, uint32_t* __restrict__ hist)
{
    if (n_dev) n = *n_dev;
    tile_hist_body(blockIdx.x, keys, n, shift, mask, nblocks, hist, nullptr);
}

// Stable scatter. Warp w of the CTA owns keys [w*256, w*256+256) of the CTA tile and walks them in order, 32 at a
// time; ranks inside a 32-key step come from match.any, ranks across steps / warps from shared counters. Pairs are first
// placed at their rank INSIDE the CTA tile in shared memory (digit-major), then written out by consecutive threads, so
// every digit run of the tile is one contiguous, coalesced burst (a direct scatter issues 32 unrelated 4-byte stores per
// warp instruction).
__device__ __forceinline__ void tile_scatter_body(uint32_t tile, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                  uint32_t mask, uint32_t nblocks, const uint32_t* __restrict__ offsets /* scanned hist */)
{
    __shared__ uint32_t s_cnt[8][256];
    __shared__ uint32_t s_gbase[256];
    __shared__ uint32_t s_keys[RADIX_ITEMS];
    __shared__ uint32_t s_vals[RADIX_ITEMS];
    __shared__ uint32_t s_warp[8];
    const uint32_t tile_base = tile * RADIX_ITEMS;
    if (tile_base >= n) return;
    __syncthreads(); // previous tile of the fused sort has left shared memory
    const uint32_t tile_n = min((uint32_t)RADIX_ITEMS, n - tile_base);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 8 * 256; i += RADIX_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    const uint32_t wbase = tile_base + warp * RADIX_WARP_ITEMS;
    uint32_t key[RADIX_PER_THREAD], val[RADIX_PER_THREAD];
#pragma unroll
    for (int k = 0; k < RADIX_PER_THREAD; k++) { // all 16 loads of the thread are in flight together
        const uint32_t i = wbase + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : 0xffffffffu;
        val[k] = i < n ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < RADIX_PER_THREAD; k++) {
        const uint32_t i = wbase + k * 32 + lane;
        if (i < n) atomicAdd(&s_cnt[warp][(key[k] >> shift) & mask], 1u);
    }
    __syncthreads();
    // digit d (thread d): counts of the 8 warps -> exclusive over warps; tile total per digit -> exclusive over digits
    uint32_t tot = 0;
    uint32_t wpre[8];
    if (threadIdx.x <= mask) {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            wpre[w] = tot;
            tot += s_cnt[w][threadIdx.x];
        }
    }
    // block-wide exclusive scan of `tot` over the 256 threads
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t dstart = incl - tot;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++)
        if (w < warp) dstart += s_warp[w];
    if (threadIdx.x <= mask) {
#pragma unroll
        for (int w = 0; w < 8; w++) s_cnt[w][threadIdx.x] = dstart + wpre[w]; // rank base inside the tile
        s_gbase[threadIdx.x] = offsets[threadIdx.x * nblocks + tile] - dstart;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RADIX_PER_THREAD; k++) {
        const uint32_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((key[k] >> shift) & mask) : 0u;
        const uint32_t peers = same_digit_lanes(d, valid);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (valid) pos = s_cnt[warp][d] + rank;
        __syncwarp();
        if (valid && rank == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) {
            s_keys[pos] = key[k];
            s_vals[pos] = val[k];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RADIX_THREADS) {
        const uint32_t kk = s_keys[i];
        const uint32_t g = s_gbase[(kk >> shift) & mask] + i;
        keys_out[g] = kk;
        vals_out[g] = s_vals[i];
    }
}

__global__ void __launch_bounds__(RADIX_THREADS) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                      uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                                                                      const uint32_t* __restrict__ n_dev, int shift, uint32_t mask, uint32_t nblocks,
                                                                      const uint32_t* __restrict__ offsets)
{
    if (n_dev) n = *n_dev;
    tile_scatter_body(blockIdx.x, keys_in, vals_in, keys_out, vals_out, n, shift, mask, nblocks, offsets);
}

// ------------------------------------------------------------------------------------------ fused (cooperative) sort
// All passes of one sort in ONE cooperative launch: per pass  tile histograms (+ per-digit totals by atomics) | grid.sync |
// per-digit row scans | grid.sync | stable scatter | grid.sync.  A 4-pass sort of ~1 M pairs is 16 dependent tiny launches in
// the multi-kernel path (~12 us each, mostly launch/drain latency); here it is one launch and 12 grid barriers.
struct RadixCoopArgs
{
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t n;
    const uint32_t* n_dev;
    int passes, digit_bits;
    uint32_t* hist;   // [bins * ntiles]
    uint32_t* rowsum; // [2][256], zeroed by the host before the launch
};

__global__ void __launch_bounds__(RADIX_THREADS) radix_sort_coop_kernel(const RadixCoopArgs a)
{
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_base;
    cg::grid_group grid = cg::this_grid();
    const uint32_t n = a.n_dev ? min(*a.n_dev, a.n) : a.n;
    const uint32_t ntiles = (n + RADIX_ITEMS - 1) / RADIX_ITEMS;
    const uint32_t bins = 1u << a.digit_bits, mask = bins - 1;
    int cur = 0;
    for (int p = 0; p < a.passes; p++) {
        const int shift = p * a.digit_bits;
        uint32_t* rowsum = a.rowsum + (p & 1) * 256;
        if (blockIdx.x == 0) a.rowsum[((p + 1) & 1) * 256 + threadIdx.x] = 0; // next pass's totals (idle during this pass)
        // ---- tile histograms ----
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) tile_hist_body(t, a.keys[cur], n, shift, mask, ntiles, a.hist, rowsum);
        grid.sync();
        // ---- digit d: exclusive scan of its row, offset by the totals of the smaller digits ----
        for (uint32_t d = blockIdx.x; d < bins; d += gridDim.x) {
            uint32_t part = (threadIdx.x < d) ? rowsum[threadIdx.x] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            __syncthreads();
            if ((threadIdx.x & 31u) == 0) s_scan[threadIdx.x >> 5] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t b = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) b += s_scan[w];
                s_base = b;
            }
            __syncthreads();
            uint32_t running = s_base;
            uint32_t* row = a.hist + (size_t)d * ntiles;
            for (uint32_t c0 = 0; c0 < ntiles; c0 += RADIX_THREADS) {
                const uint32_t i = c0 + threadIdx.x;
                const uint32_t v = i < ntiles ? row[i] : 0u;
                uint32_t total;
                const uint32_t excl = block_excl_scan_256(v, s_scan, total);
                if (i < ntiles) row[i] = running + excl;
                running += total;
            }
        }
        grid.sync();
        // ---- stable scatter ----
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x)
            tile_scatter_body(t, a.keys[cur], a.vals[cur], a.keys[cur ^ 1], a.vals[cur ^ 1], n, shift, mask, ntiles, a.hist);
        grid.sync();
        cur ^= 1;
    }
}

// ------------------------------------------------------------------------------------------ single-kernel passes (decoupled look-back)
// One kernel per radix pass instead of histogram + 2-kernel scan + scatter. The digit totals of ALL passes are known before the
// first pass (accumulated by the kernel that produces the keys, or by digit_hist_all_kernel), so a pass only needs, per CTA tile
// and digit, the number of keys with that digit in the tiles before it. Tile t publishes its own count (flag AGGREGATE) as soon as
// it has counted, walks back over its predecessors' published words until it meets an INCLUSIVE one, and publishes its own
// inclusive prefix. Tiles are handed out by an atomic ticket, so every predecessor of a running tile is running or done and the
// walk cannot dead-lock. Ranking inside the tile is the stable match.any scheme of tile_scatter_body.
constexpr uint32_t LB_AGG = 1u << 30, LB_INC = 2u << 30, LB_VAL = (1u << 30) - 1u;
constexpr int LB_BATCH = 8; // predecessors whose look-back words are requested together
// Keys per thread of the look-back passes (tile = 256 x LB_PER_THREAD keys): larger tiles amortise the look-back walk and the
// fixed per-tile work. GSR_LB_PT=8 selects 2048-key tiles for A/B runs.

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct LookbackPassArgs
{
    const uint32_t* keys_in;
    const uint32_t* vals_in;
    uint32_t* keys_out;
    uint32_t* vals_out;
    uint32_t n;
    const uint32_t* n_dev;
    int shift;
    uint32_t mask;
    const uint32_t* digit_total; // [256] keys per digit of this pass (whole input)
    uint32_t* status;            // [tiles][256], zeroed
    uint32_t* ticket;            // zeroed
};

template <int LB_PER_THREAD>
__global__ void __launch_bounds__(RADIX_THREADS, 4) radix_lookback_pass_kernel(const LookbackPassArgs a)
{
    constexpr int LB_ITEMS = RADIX_THREADS * LB_PER_THREAD;
    __shared__ uint32_t s_cnt[8][256];
    __shared__ uint32_t s_gbase[256];
    __shared__ uint32_t s_keys[LB_ITEMS];
    __shared__ uint32_t s_vals[LB_ITEMS];
    __shared__ uint32_t s_warp[8], s_warp2[8];
    __shared__ uint32_t s_tile;
    const uint32_t n = a.n_dev ? min(*a.n_dev, a.n) : a.n;
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    for (uint32_t i = threadIdx.x; i < 8 * 256; i += RADIX_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * LB_ITEMS;
    if (tile_base >= n) return;
    const uint32_t tile_n = min((uint32_t)LB_ITEMS, n - tile_base);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int shift = a.shift;
    const uint32_t mask = a.mask;

    const uint32_t wbase = tile_base + warp * (LB_ITEMS / 8);
    uint32_t key[LB_PER_THREAD], val[LB_PER_THREAD];
#pragma unroll
    for (int k = 0; k < LB_PER_THREAD; k++) {
        const uint32_t i = wbase + k * 32 + lane;
        key[k] = i < n ? a.keys_in[i] : 0xffffffffu;
        val[k] = i < n ? a.vals_in[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < LB_PER_THREAD; k++) {
        const uint32_t i = wbase + k * 32 + lane;
        if (i < n) atomicAdd(&s_cnt[warp][(key[k] >> shift) & mask], 1u);
    }
    __syncthreads();
    const bool has_digit = threadIdx.x <= mask;
    uint32_t tot = 0, wpre[8];
    if (has_digit) {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            wpre[w] = tot;
            tot += s_cnt[w][threadIdx.x];
        }
        st_volatile_u32(a.status + (size_t)tile * 256 + threadIdx.x, (tile == 0 ? LB_INC : LB_AGG) | tot); // publish early
    }
    // two block-wide exclusive scans over the digits: this tile's counts (position of the digit run inside the tile) and the
    // global digit totals (start of the digit's region in the output)
    const uint32_t gtot = has_digit ? a.digit_total[threadIdx.x] : 0u;
    uint32_t incl = tot, gincl = gtot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o), g = __shfl_up_sync(0xffffffffu, gincl, o);
        if (lane >= (uint32_t)o) {
            incl += v;
            gincl += g;
        }
    }
    if (lane == 31) {
        s_warp[warp] = incl;
        s_warp2[warp] = gincl;
    }
    __syncthreads();
    uint32_t dstart = incl - tot, gstart = gincl - gtot;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++)
        if (w < warp) {
            dstart += s_warp[w];
            gstart += s_warp2[w];
        }
    if (has_digit) {
#pragma unroll
        for (int w = 0; w < 8; w++) s_cnt[w][threadIdx.x] = dstart + wpre[w]; // rank base inside the tile
    }
    __syncthreads();
    // ---- rank the keys inside the tile and park them in shared memory in tile order: everything that does NOT need the
    //      predecessors. The look-back comes after it, so the tiles in front have had this whole phase to publish. ----
#pragma unroll
    for (int k = 0; k < LB_PER_THREAD; k++) {
        const uint32_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((key[k] >> shift) & mask) : 0u;
        const uint32_t peers = same_digit_lanes(d, valid);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (valid) pos = s_cnt[warp][d] + rank;
        __syncwarp();
        if (valid && rank == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) {
            s_keys[pos] = key[k];
            s_vals[pos] = val[k];
        }
    }
    if (has_digit) {
        uint32_t excl = 0;
        if (tile > 0) {
            // Walk back over the predecessors' words LB_BATCH at a time: the loads of a batch are independent, so a walk over k
            // tiles costs ~k / LB_BATCH L2 round trips instead of k.
            int t = (int)tile - 1;
            bool found = false;
            while (!found) {
                uint32_t v[LB_BATCH];
#pragma unroll
                for (int j = 0; j < LB_BATCH; j++) // entries before tile 0 are never consumed: tile 0 always publishes INCLUSIVE
                    v[j] = t - j >= 0 ? ld_volatile_u32(a.status + (size_t)(t - j) * 256 + threadIdx.x) : LB_INC;
#pragma unroll
                for (int j = 0; j < LB_BATCH; j++) {
                    if (!found) {
                        uint32_t w = v[j];
                        while ((w & ~LB_VAL) == 0u) w = ld_volatile_u32(a.status + (size_t)(t - j) * 256 + threadIdx.x); // not published yet
                        excl += w & LB_VAL;
                        found = (w & LB_INC) != 0u;
                    }
                }
                t -= LB_BATCH;
            }
            st_volatile_u32(a.status + (size_t)tile * 256 + threadIdx.x, LB_INC | (excl + tot));
        }
        s_gbase[threadIdx.x] = gstart + excl - dstart;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RADIX_THREADS) {
        const uint32_t kk = s_keys[i];
        const uint32_t g = s_gbase[(kk >> shift) & mask] + i;
        a.keys_out[g] = kk;
        a.vals_out[g] = s_vals[i];
    }
}

// digit totals of every pass in one read of the keys (for callers whose key producer does not accumulate them itself)
__global__ void __launch_bounds__(RADIX_THREADS) digit_hist_all_kernel(const uint32_t* __restrict__ keys, uint32_t n, const uint32_t* n_dev,
                                                                      int passes, int digit_bits, uint32_t* hist /*[passes][256]*/)
{
    __shared__ uint32_t s_h[4][256];
    if (n_dev) n = min(*n_dev, n);
    for (int p = 0; p < passes; p++) s_h[p][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t mask = (1u << digit_bits) - 1u;
    for (uint32_t i = blockIdx.x * RADIX_THREADS + threadIdx.x; i < n; i += gridDim.x * RADIX_THREADS) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&s_h[p][(k >> (p * digit_bits)) & mask], 1u);
    }
    __syncthreads();
    for (int p = 0; p < passes; p++) {
        const uint32_t c = s_h[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
    }
}
} // namespace

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, bool write_total, uint32_t* partials, cudaStream_t s)
{
    if (n == 0) {
        if (write_total) GSR_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), s));
        return 0;
    }
    const uint32_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, n, partials); count_launches(1);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, n, partials, out, write_total ? 1 : 0); count_launches(1);
    return 0;
}

int radix_num_passes(int nbits) { return nbits <= 0 ? 0 : (nbits + 7) / 8; }

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int nbits, uint32_t* hist, size_t hist_words, cudaStream_t s,
                     const uint32_t* n_dev)
{
    const int passes = radix_num_passes(nbits);
    if (n == 0 || passes == 0) return 0;
    const uint32_t nb = (n + RADIX_ITEMS - 1) / RADIX_ITEMS;
    const int digit_bits = (nbits + passes - 1) / passes;
    const uint32_t bins = 1u << digit_bits;
    const size_t hwords = (size_t)bins * nb;
    const size_t pwords = (hwords + SCAN_ITEMS - 1) / SCAN_ITEMS;
    if (hwords + pwords > hist_words) {
        set_error("radix_sort_pairs: histogram workspace too small (%zu > %zu words)", hwords + pwords, hist_words);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    // Fused cooperative sort (one launch for all passes), opt-in with GSR_SORT_COOP=1. Measured on B200 at cfg3 it is NOT faster
    // than the multi-kernel path (binning 0.585 ms vs 0.532 ms: the grid barriers and the per-CTA tile loops cost more than the
    // launch latency they remove, because the sorts are enqueued ahead of the GPU anyway), so the multi-kernel path is the default.
    static const int coop_mode = getenv("GSR_SORT_COOP") ? atoi(getenv("GSR_SORT_COOP")) : 0;
    if (coop_mode && hwords + 512 <= hist_words) {
        int dev = 0, coop = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radix_sort_coop_kernel, RADIX_THREADS, 0);
        if (coop && sms > 0 && per_sm > 0) {
            RadixCoopArgs ca;
            ca.keys[0] = keys[0]; ca.keys[1] = keys[1]; ca.vals[0] = vals[0]; ca.vals[1] = vals[1];
            ca.n = n; ca.n_dev = n_dev; ca.passes = passes; ca.digit_bits = digit_bits;
            ca.hist = hist; ca.rowsum = hist + hwords;
            GSR_CUDA(cudaMemsetAsync(ca.rowsum, 0, 512 * sizeof(uint32_t), s));
            uint32_t want = nb > bins ? nb : bins;
            uint32_t grid = (uint32_t)(sms * per_sm);
            if (grid > want) grid = want;
            void* kargs[] = {(void*)&ca};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)radix_sort_coop_kernel, dim3(grid), dim3(RADIX_THREADS), kargs, 0, s);
            if (e == cudaSuccess) {
                count_launches(1);
                return passes & 1;
            }
            (void)cudaGetLastError(); // fall back to the multi-kernel path
        }
    }
    uint32_t* partials = hist + hwords;
    int cur = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = p * digit_bits;
        const uint32_t mask = bins - 1;
        radix_hist_kernel<<<nb, RADIX_THREADS, 0, s>>>(keys[cur], n, n_dev, shift, mask, nb, hist); count_launches(1);
        int rc = exclusive_scan_u32(hist, hist, (uint32_t)hwords, false, partials, s);
        if (rc) return rc;
        radix_scatter_kernel<<<nb, RADIX_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, n_dev, shift, mask, nb, hist);
        count_launches(1);
        cur ^= 1;
    }
    return cur;
}
} // namespace gsr

namespace gsr
{
size_t radix_lookback_ws_words(uint32_t n_cap, int nbits)
{
    const int passes = radix_num_passes(nbits);
    const size_t tiles = ((size_t)n_cap + RADIX_ITEMS - 1) / RADIX_ITEMS; // sized for the smallest tile the passes may use
    return (size_t)passes * 256 + 32 + (size_t)passes * tiles * 256;
}

int radix_digit_bits(int nbits)
{
    const int passes = radix_num_passes(nbits);
    return passes ? (nbits + passes - 1) / passes : 0;
}

// Stable LSD radix sort, ONE kernel per pass. ws: radix_lookback_ws_words(n, nbits) words laid out as
// [passes][256] digit totals | 32 tickets | [passes][tiles][256] look-back words. With hist_ready the caller has zeroed the
// whole workspace and its key producer has accumulated the digit totals (digit p = bits [p * digit_bits, (p + 1) * digit_bits));
// otherwise the workspace is zeroed and the totals are computed here. n < 2^30 (the look-back words carry 30 bits).
int radix_sort_pairs_lookback(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int nbits, uint32_t* ws, size_t ws_words, bool hist_ready,
                              cudaStream_t s, const uint32_t* n_dev, int keys_per_thread)
{
    const int passes = radix_num_passes(nbits);
    if (n == 0 || passes == 0) return 0;
    if (n >= (1u << 30) || passes > 4) return radix_sort_pairs(keys, vals, n, nbits, ws, ws_words, s, n_dev); // multi-kernel path
    if (radix_lookback_ws_words(n, nbits) > ws_words) {
        set_error("radix_sort_pairs_lookback: workspace too small (%zu > %zu words)", radix_lookback_ws_words(n, nbits), ws_words);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const int digit_bits = radix_digit_bits(nbits);
    static const int pt_env = getenv("GSR_LB_PT") ? atoi(getenv("GSR_LB_PT")) : 0;
    const int pt = pt_env ? pt_env : keys_per_thread;
    const uint32_t items = RADIX_THREADS * (uint32_t)(pt == 8 ? 8 : 16);
    const uint32_t tiles = (n + items - 1) / items;
    const uint32_t tiles_cap = (n + RADIX_ITEMS - 1) / RADIX_ITEMS;
    uint32_t* hist = ws;
    uint32_t* tickets = ws + (size_t)passes * 256;
    uint32_t* status = tickets + 32;
    if (!hist_ready) {
        GSR_CUDA(cudaMemsetAsync(ws, 0, radix_lookback_ws_words(n, nbits) * sizeof(uint32_t), s));
        const uint32_t grid = tiles_cap < 1184u ? tiles_cap : 1184u;
        digit_hist_all_kernel<<<grid, RADIX_THREADS, 0, s>>>(keys[0], n, n_dev, passes, digit_bits, hist); count_launches(1);
    }
    int cur = 0;
    for (int p = 0; p < passes; p++) {
        LookbackPassArgs a;
        a.keys_in = keys[cur]; a.vals_in = vals[cur]; a.keys_out = keys[cur ^ 1]; a.vals_out = vals[cur ^ 1];
        a.n = n; a.n_dev = n_dev; a.shift = p * digit_bits; a.mask = (1u << digit_bits) - 1u;
        a.digit_total = hist + p * 256; a.status = status + (size_t)p * tiles_cap * 256; a.ticket = tickets + p;
        if (pt == 8) radix_lookback_pass_kernel<8><<<tiles, RADIX_THREADS, 0, s>>>(a);
        else radix_lookback_pass_kernel<16><<<tiles, RADIX_THREADS, 0, s>>>(a);
        count_launches(1);
        cur ^= 1;
    }
    return cur;
}
} // namespace gsr
