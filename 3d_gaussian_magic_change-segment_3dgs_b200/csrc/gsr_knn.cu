// simple-knn rebuilt: mean squared distance of every point to its 3 nearest neighbours (distCUDA2).
//
// Reference: SimpleKNN::knn (simple-knn/simple_knn.cu:185-221): AABB reduce (two blocking D2H copies), 30-bit Morton
// codes, CUB sort, boxes of 1024 sorted points, then one THREAD per point scanning every box whose AABB is closer
// than its current 3rd-best -- O(P * P/1024) box tests and 1024-point serial scans per surviving box.
// The result is the exact 3-NN mean, so any exact search gives the same bits as long as each squared distance is
// evaluated with the reference's expression (updateKBest, simple_knn.cu:131-145) and the three best are summed in
// ascending order. Here:
//   * the AABB stays on the device (no host round trip);
//   * the Morton sort uses this library's radix sort; sorted points are gathered into a float4 array once;
//   * a 32-ary AABB hierarchy over 64-point leaves replaces the flat box list;
//   * one WARP owns 32 consecutive sorted queries (all in one leaf) and walks the hierarchy cooperatively: lanes
//     test 32 child boxes at once against the warp's query AABB and its worst current 3rd-best distance, and a
//     surviving leaf is staged to shared memory and broadcast to all 32 queries.
#include <float.h>

#include "gsr_common.cuh"

namespace gsr
{
namespace
{
constexpr int LEAF = 64;
constexpr int FAN = 32;
constexpr int MAX_LEVELS = 8;

struct KnnLevels
{
    int nlevels;               // level 0 = leaves
    uint32_t count[MAX_LEVELS];
    float4* bmin[MAX_LEVELS];
    float4* bmax[MAX_LEVELS];
};

struct KnnWs
{
    float* minmax;      // [8] min xyz, pad, max xyz, pad
    float* partial;     // [6 * 1024]
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;
    size_t hist_words;
    float4* spts;       // sorted points, w = original index bits
    KnnLevels lv;
};

size_t knn_layout(char* base, int P, KnnWs& w)
{
    char* p = base;
    carve(p, w.minmax, (size_t)8);
    carve(p, w.partial, (size_t)6 * 1024);
    for (int i = 0; i < 2; i++) {
        carve(p, w.keys[i], (size_t)P);
        carve(p, w.vals[i], (size_t)P);
    }
    const size_t nb = ((size_t)P + RADIX_ITEMS - 1) / RADIX_ITEMS;
    const size_t h = 256 * (nb + 1);
    w.hist_words = h + (h + SCAN_ITEMS - 1) / SCAN_ITEMS + 64;
    carve(p, w.hist, w.hist_words);
    carve(p, w.spts, (size_t)P);
    uint32_t n = (uint32_t)(((size_t)P + LEAF - 1) / LEAF);
    int l = 0;
    for (;;) {
        w.lv.count[l] = n;
        carve(p, w.lv.bmin[l], (size_t)n);
        carve(p, w.lv.bmax[l], (size_t)n);
        l++;
        if (n <= FAN || l >= MAX_LEVELS) break;
        n = (n + FAN - 1) / FAN;
    }
    w.lv.nlevels = l;
    return (size_t)(p - base) + 256;
}

// ---- AABB of all points; the initial value (0,0,0) takes part in both reductions (simple_knn.cu:191-200) ----
__global__ void __launch_bounds__(256) aabb_partial_kernel(int P, const float* __restrict__ pts, float* __restrict__ partial)
{
    __shared__ float s[6][8];
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float v = pts[3 * (size_t)i + a];
            mn[a] = min(mn[a], v);
            mx[a] = max(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            s[a][threadIdx.x >> 5] = mn[a];
            s[3 + a][threadIdx.x >> 5] = mx[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? min(v, s[threadIdx.x][w]) : max(v, s[threadIdx.x][w]);
        partial[threadIdx.x * 1024 + blockIdx.x] = v;
    }
}
__global__ void aabb_final_kernel(int nparts, const float* __restrict__ partial, float* __restrict__ minmax)
{
    // 6 warps, one per component
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v = 0.f;
    for (int i = lane; i < nparts; i += 32) v = c < 3 ? min(v, partial[c * 1024 + i]) : max(v, partial[c * 1024 + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = c < 3 ? min(v, u) : max(v, u);
    }
    if (lane == 0) minmax[c < 3 ? c : c + 1] = v;
}

__device__ __forceinline__ uint32_t spread_bits_10(uint32_t x) // simple_knn.cu:45-52
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ minmax,
                                                     uint32_t* __restrict__ codes, uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float3 minn = {minmax[0], minmax[1], minmax[2]}, maxx = {minmax[4], minmax[5], minmax[6]};
    const float3 c = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    const uint32_t x = spread_bits_10(((c.x - minn.x) / (maxx.x - minn.x)) * ((1 << 10) - 1));
    const uint32_t y = spread_bits_10(((c.y - minn.y) / (maxx.y - minn.y)) * ((1 << 10) - 1));
    const uint32_t z = spread_bits_10(((c.z - minn.z) / (maxx.z - minn.z)) * ((1 << 10) - 1));
    codes[i] = x | (y << 1) | (z << 2);
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_points_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order,
                                                            float4* __restrict__ spts)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const uint32_t o = order[i];
    spts[i] = {pts[3 * (size_t)o], pts[3 * (size_t)o + 1], pts[3 * (size_t)o + 2], __uint_as_float(o)};
}

// one warp per leaf (64 points, 2 per lane)
__global__ void __launch_bounds__(256) leaf_boxes_kernel(int P, const float4* __restrict__ spts, uint32_t nleaves, float4* __restrict__ bmin,
                                                         float4* __restrict__ bmax)
{
    const uint32_t leaf = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (leaf >= nleaves) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
    for (int k = 0; k < LEAF / 32; k++) {
        const size_t i = (size_t)leaf * LEAF + k * 32 + lane;
        if (i < (size_t)P) {
            const float4 p = spts[i];
            mn[0] = min(mn[0], p.x); mn[1] = min(mn[1], p.y); mn[2] = min(mn[2], p.z);
            mx[0] = max(mx[0], p.x); mx[1] = max(mx[1], p.y); mx[2] = max(mx[2], p.z);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if (lane == 0) {
        bmin[leaf] = {mn[0], mn[1], mn[2], 0.f};
        bmax[leaf] = {mx[0], mx[1], mx[2], 0.f};
    }
}

// one warp per parent node (32 children, one per lane)
__global__ void __launch_bounds__(256) parent_boxes_kernel(uint32_t nchildren, const float4* __restrict__ cmin, const float4* __restrict__ cmax,
                                                           uint32_t nparents, float4* __restrict__ pmin, float4* __restrict__ pmax)
{
    const uint32_t node = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (node >= nparents) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const uint32_t c = node * FAN + lane;
    if (c < nchildren) {
        const float4 a = cmin[c], b = cmax[c];
        mn[0] = a.x; mn[1] = a.y; mn[2] = a.z;
        mx[0] = b.x; mx[1] = b.y; mx[2] = b.z;
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if (lane == 0) {
        pmin[node] = {mn[0], mn[1], mn[2], 0.f};
        pmax[node] = {mx[0], mx[1], mx[2], 0.f};
    }
}

__device__ __forceinline__ void keep_3_best(float dist, float (&knn)[3]) // updateKBest<3>, simple_knn.cu:137-144
{
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            const float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256) knn_search_kernel(int P, const float4* __restrict__ spts, KnnLevels lv, float* __restrict__ out)
{
    __shared__ float4 s_leaf[8][LEAF];
    __shared__ uint32_t s_first[8][MAX_LEVELS + 1];
    __shared__ uint32_t s_mask[8][MAX_LEVELS + 1];
    __shared__ uint32_t s_level[8][MAX_LEVELS + 1];

    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t qbase = ((size_t)blockIdx.x * 8 + w) * 32;
    if (qbase >= (size_t)P) return; // whole warp
    const size_t qi = qbase + lane;
    const bool qvalid = qi < (size_t)P;
    const float4 q = spts[qvalid ? qi : (size_t)P - 1];
    const uint32_t own_leaf = (uint32_t)(qbase / LEAF);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};

    // query AABB of the warp
    const float qmnx = warp_min(q.x), qmny = warp_min(q.y), qmnz = warp_min(q.z);
    const float qmxx = warp_max(q.x), qmxy = warp_max(q.y), qmxz = warp_max(q.z);

    auto visit_leaf = [&](uint32_t leaf) {
        const size_t first = (size_t)leaf * LEAF;
        const uint32_t cnt = (uint32_t)min((size_t)LEAF, (size_t)P - first);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < LEAF / 32; k++) {
            const uint32_t j = k * 32 + lane;
            if (j < cnt) s_leaf[w][j] = spts[first + j];
        }
        __syncwarp();
        for (uint32_t j = 0; j < cnt; j++) {
            const float4 c = s_leaf[w][j];
            if (first + j == qi) continue; // self, excluded by index (duplicates at distance 0 count)
            const float3 d = {c.x - q.x, c.y - q.y, c.z - q.z};
            // d.x*d.x + d.y*d.y + d.z*d.z exactly as nvcc contracts it in the reference's updateKBest (SASS of boxMeanDist:
            // FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz). Spelled with intrinsics so the association cannot drift with context.
            const float dist = __fmaf_rn(d.z, d.z, __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y)));
            keep_3_best(dist, best);
        }
    };
    // children [first, first+count) of `level`: which may hold a point closer than the warp's worst 3rd-best?
    auto test_children = [&](int level, uint32_t first, uint32_t count) -> uint32_t {
        const float bound = warp_max(qvalid ? best[2] : 0.f);
        bool need = false;
        if (lane < count) {
            const float4 a = lv.bmin[level][first + lane], b = lv.bmax[level][first + lane];
            const float gx = max(0.f, max(a.x - qmxx, qmnx - b.x));
            const float gy = max(0.f, max(a.y - qmxy, qmny - b.y));
            const float gz = max(0.f, max(a.z - qmxz, qmnz - b.z));
            const float d2 = gx * gx + gy * gy + gz * gz;
            need = !(d2 * 0.9999f > bound); // conservative w.r.t. fp32 rounding of the point distances
        }
        return __ballot_sync(0xffffffffu, need);
    };

    visit_leaf(own_leaf); // establishes a tight bound before the walk

    const int top = lv.nlevels - 1;
    int sp = 0;
    if (lane == 0) {
        s_level[w][0] = (uint32_t)top;
        s_first[w][0] = 0;
    }
    {
        const uint32_t m = test_children(top, 0, lv.count[top]);
        if (lane == 0) s_mask[w][0] = m;
    }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        const uint32_t mask = s_mask[w][sp - 1];
        if (mask == 0) {
            sp--;
            continue;
        }
        const int level = (int)s_level[w][sp - 1];
        const uint32_t c = __ffs(mask) - 1;
        const uint32_t node = s_first[w][sp - 1] + c;
        __syncwarp();
        if (lane == 0) s_mask[w][sp - 1] = mask & (mask - 1);
        __syncwarp();
        if (level == 0) {
            if (node != own_leaf) {
                // the bound may have shrunk since this leaf was selected: re-test it alone
                if (test_children(0, node, 1) & 1u) visit_leaf(node);
            }
        } else {
            const uint32_t cfirst = node * FAN;
            const uint32_t ccount = min((uint32_t)FAN, lv.count[level - 1] - cfirst);
            const uint32_t m = test_children(level - 1, cfirst, ccount);
            if (lane == 0) {
                s_level[w][sp] = (uint32_t)(level - 1);
                s_first[w][sp] = cfirst;
                s_mask[w][sp] = m;
            }
            sp++;
            __syncwarp();
        }
    }
    if (qvalid) out[__float_as_uint(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}
} // namespace

size_t knn_workspace_bytes(int P)
{
    if (P <= 0) return 0;
    KnnWs w;
    return knn_layout(nullptr, P, w) + 256;
}

int knn_run(int P, const float* points, float* out, void* ws, size_t ws_bytes, cudaStream_t s)
{
    KnnWs w;
    char* base = (char*)align_up((size_t)ws, 256);
    knn_layout(base, P, w);
    (void)ws_bytes;
    const int nparts = min(1024, (P + 255) / 256);
    aabb_partial_kernel<<<nparts, 256, 0, s>>>(P, points, w.partial); count_launches(1);
    aabb_final_kernel<<<1, 192, 0, s>>>(nparts, w.partial, w.minmax); count_launches(1);
    morton_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, w.minmax, w.keys[0], w.vals[0]); count_launches(1);
    GSR_LAUNCHED(s, false, "knn_morton");
    const int res = radix_sort_pairs(w.keys, w.vals, (uint32_t)P, 30, w.hist, w.hist_words, s);
    if (res < 0) return res;
    gather_points_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, w.vals[res], w.spts); count_launches(1);
    leaf_boxes_kernel<<<(w.lv.count[0] + 7) / 8, 256, 0, s>>>(P, w.spts, w.lv.count[0], w.lv.bmin[0], w.lv.bmax[0]); count_launches(1);
    for (int l = 1; l < w.lv.nlevels; l++) {
        parent_boxes_kernel<<<(w.lv.count[l] + 7) / 8, 256, 0, s>>>(w.lv.count[l - 1], w.lv.bmin[l - 1], w.lv.bmax[l - 1], w.lv.count[l], w.lv.bmin[l],
                                                                   w.lv.bmax[l]); count_launches(1);
    }
    const uint32_t nwarps = (uint32_t)(((size_t)P + 31) / 32);
    knn_search_kernel<<<(nwarps + 7) / 8, 256, 0, s>>>(P, w.spts, w.lv, out); count_launches(1);
    GSR_LAUNCHED(s, false, "knn_search");
    return 0;
}
} // namespace gsr
