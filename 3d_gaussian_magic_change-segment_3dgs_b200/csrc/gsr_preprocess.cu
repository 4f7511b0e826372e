// Per-Gaussian stages: forward preprocess (cull, EWA projection, SH colour, tile rectangle, slot compaction),
// backward preprocess (conic/cov2D, projection, depth, SH and scale/rotation gradients, dense gradient rows),
// and the frustum mark.
//
// Replaces the reference kernels preprocessCUDA (forward.cu:155-256), computeCov2DCUDA + preprocessCUDA
// (backward.cu:144-274, 346-412) and checkFrustum (rasterizer_impl.cu:54-66), plus the 11 zero-fills of
// rasterize_points.cu:166-177 (every gradient row is written exactly once here).
#include <stdlib.h>

#include "gsr_math.cuh"

namespace gsr
{
namespace
{
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Exclusive prefix of `flag` over the 256-thread block (thread order) + block total.
__device__ __forceinline__ uint32_t block_rank_256(bool flag, uint32_t* s_warp /*[8]*/, uint32_t& total)
{
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
    const uint32_t warp = threadIdx.x >> 5;
    if (lane_id() == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) {
        const uint32_t c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    return base + __popc(ballot & ((1u << lane_id()) - 1u));
}

// HOIST: the scale / rotation loads are issued together with the position's, before the near cull is known (one DRAM round trip
// instead of two on the kernel's critical path; the ~10 % culled Gaussians cost 28 wasted bytes each).
template <bool HOIST>
__global__ void __launch_bounds__(PRE_BLOCK) preprocess_fwd_kernel(const PreFwdArgs a)
{
    __shared__ float s_view[16], s_proj[16];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_tiles[8];

    if (threadIdx.x < 16) {
        s_view[threadIdx.x] = a.view[threadIdx.x];
        s_proj[threadIdx.x] = a.proj[threadIdx.x];
    }
    __syncthreads();

    const int idx = blockIdx.x * PRE_BLOCK + threadIdx.x;
    bool visible = false;
    bool filtered = false;
    int radius_out = 0;
    float4 A, B, C;
    ushort4 rect = {0, 0, 0, 0};
    unsigned clamp_bits = 0;
    uint32_t tiles = 0;

    int gid = idx; // row of the input tensors: the position itself, or the list entry when an index list is rendered
    if (idx < a.P) {
        PartPtrs in = {a.means3D, a.scales, a.rotations, a.opacities, a.shs, a.cov3D_precomp, a.colors_precomp, a.segments};
        if (a.subset) gid = a.subset[idx];
        if (a.num_parts > 0) { // fused sub-scenes: this Gaussian's part and its row there (a CTA may straddle a boundary)
            int k = 0;
            while (k + 1 < a.num_parts && idx >= a.part_start[k + 1]) k++;
            in = a.part[k];
            gid = idx - a.part_start[k];
        }
        const float3 p_orig = {in.means3D[3 * gid], in.means3D[3 * gid + 1], in.means3D[3 * gid + 2]};
        float3 sc_pre = {0.f, 0.f, 0.f};
        float4 q_pre = {0.f, 0.f, 0.f, 0.f};
        if (HOIST && in.cov3D_precomp == nullptr) {
            sc_pre = {in.scales[3 * gid], in.scales[3 * gid + 1], in.scales[3 * gid + 2]};
            q_pre = *reinterpret_cast<const float4*>(in.rotations + 4 * (size_t)gid);
        }
        // near cull (auxiliary.h:139-164); the +-1.3 NDC test is disabled in the reference
        const float4 p_hom = xform_point_h(p_orig, s_proj);
        const float p_w = 1.0f / (p_hom.w + 0.0000001f);
        const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
        const float3 p_view = xform_point(p_orig, s_view);
        if (p_view.z <= 0.2f) {
            filtered = a.prefiltered != 0;
        } else {
            float cov3D[6];
            if (in.cov3D_precomp != nullptr) {
#pragma unroll
                for (int i = 0; i < 6; i++) cov3D[i] = in.cov3D_precomp[6 * (size_t)gid + i];
            } else {
                float3 sc = sc_pre;
                float4 q = q_pre;
                if (!HOIST) {
                    sc = {in.scales[3 * gid], in.scales[3 * gid + 1], in.scales[3 * gid + 2]};
                    q = *reinterpret_cast<const float4*>(in.rotations + 4 * (size_t)gid);
                }
                if (a.raw) { // fused activations (scene/gaussian_model.py:100-106): exp, normalize
                    sc = {expf(sc.x), expf(sc.y), expf(sc.z)};
                    q = act_normalize(q);
                }
                cov3d_from_scale_rot(sc, a.scale_modifier, q, cov3D);
            }
            const float3 cov = cov2d_project(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, s_view, nullptr);
            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det != 0.0f) {
                const float det_inv = 1.f / det;
                const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};
                const float mid = 0.5f * (cov.x + cov.z);
                const float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
                const float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
                const float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
                const float2 point_image = {ndc_to_pix(p_proj.x, a.W), ndc_to_pix(p_proj.y, a.H)};
                uint2 rmin, rmax;
                tile_rect(point_image, my_radius, a.grid_x, a.grid_y, rmin, rmax);
                if ((rmax.x - rmin.x) * (rmax.y - rmin.y) != 0) {
                    visible = true;
                    float3 rgb;
                    if (in.colors_precomp == nullptr) {
                        const float3 campos = {a.campos[0], a.campos[1], a.campos[2]};
                        const float3* sh3 = reinterpret_cast<const float3*>(in.shs);
                        const ShCoeffs<float3> sh = a.raw ? ShCoeffs<float3>{sh3 + gid, reinterpret_cast<const float3*>(a.shs_rest) + (size_t)gid * (a.M - 1)}
                                                          : ShCoeffs<float3>{sh3 + (size_t)gid * a.M, sh3 + (size_t)gid * a.M + 1};
                        rgb = sh_to_rgb(a.D, p_orig, campos, sh, clamp_bits);
                    } else {
                        rgb = {in.colors_precomp[3 * gid], in.colors_precomp[3 * gid + 1], in.colors_precomp[3 * gid + 2]};
                    }
                    float s0 = 0.f, s1 = 0.f;
                    if (a.S == 2 && in.segments != nullptr) {
                        const float2 sg = *reinterpret_cast<const float2*>(in.segments + 2 * (size_t)gid);
                        s0 = a.raw ? act_sigmoid(sg.x) : sg.x;
                        s1 = a.raw ? act_sigmoid(sg.y) : sg.y;
                    } else if (a.S > 0 && in.segments != nullptr) { // runtime class count: channels 0-1 here, the other pairs below
                        const float* sg = in.segments + (size_t)a.S * gid;
                        s0 = a.raw ? act_sigmoid(sg[0]) : sg[0];
                        if (a.S > 1) s1 = a.raw ? act_sigmoid(sg[1]) : sg[1];
                    }
                    const float op = a.raw ? act_sigmoid(in.opacities[gid]) : in.opacities[gid];
                    radius_out = my_radius;
                    A = {point_image.x, point_image.y, conic.x, conic.y};
                    B = {conic.z, op, rgb.x, rgb.y};
                    C = {rgb.z, p_view.z, s0, s1};
                    rect = {(unsigned short)rmin.x, (unsigned short)rmin.y, (unsigned short)rmax.x, (unsigned short)rmax.y};
                    tiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
                }
            }
        }
        a.radii[idx] = radius_out;
    }

    // slot compaction inside the block (stable in Gaussian-id order); the block's instance count and its "filtered although
    // prefiltered" flag ride on the same barrier (one __syncthreads per CTA instead of two)
    uint32_t t = tiles;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const bool warp_filtered = __any_sync(0xffffffffu, filtered);
    if (lane_id() == 0) s_tiles[threadIdx.x >> 5] = t | (warp_filtered ? 0x80000000u : 0u); // a warp's 32 rectangles stay far below 2^31 tiles
    uint32_t nvis;
    const uint32_t rank = block_rank_256(visible, s_warp, nvis);
    if (visible) {
        const uint32_t slot = blockIdx.x * PRE_BLOCK + rank;
        a.g.rec[3 * (size_t)slot + 0] = A;
        a.g.rec[3 * (size_t)slot + 1] = B;
        a.g.rec[3 * (size_t)slot + 2] = C;
        a.g.rect[slot] = rect;
        a.g.dkeys[1][slot] = __float_as_uint(C.y); // depth bits by slot: what depth_keys compacts (dkeys[1] is free until the sort's first pass)
        a.g.slot_gid[slot] = (uint32_t)(a.num_parts > 0 ? idx : gid); // the backward addresses inputs and gradient rows through this
        a.g.clamped[slot] = (uint8_t)clamp_bits;
        for (uint32_t k = 0; k < a.g.extra_pairs; k++) { // num_class > 2: segment channel pairs 1.. of this slot
            float2 v = {0.f, 0.f};
            if (a.segments != nullptr) {
                const float* sg = a.segments + (size_t)a.S * gid + 2 * (k + 1);
                v.x = a.raw ? act_sigmoid(sg[0]) : sg[0];
                if (2 * (k + 1) + 1 < (uint32_t)a.S) v.y = a.raw ? act_sigmoid(sg[1]) : sg[1];
            }
            a.g.seg_extra[(size_t)k * a.g.slots + slot] = v;
        }
    }
    // instance count of the block -> global R (integer atomics: deterministic total)
    if (threadIdx.x == 0) {
        uint32_t tt = 0, any_filtered = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            tt += s_tiles[w] & 0x7fffffffu;
            any_filtered |= s_tiles[w] >> 31;
        }
        a.g.blk_count[blockIdx.x] = nvis;
        if (tt) atomicAdd(reinterpret_cast<unsigned long long*>(a.g.counters + CNT_RENDERED_LO), (unsigned long long)tt);
        if (any_filtered) atomicOr(a.g.counters + CNT_ERROR, 1u);
    }
}

// Exclusive scan of the per-block visible counts (nblk = a few 10^4 values, L2 resident): one CTA, every thread sums a run of
// consecutive counts, one 1024-wide scan of the run totals, and the runs are written out. (Round 1 scanned 1024 counts per trip with
// four barriers each: 23 us at cfg3, on the critical path between the preprocess and the depth sort.)
// Writes blk_offset[0..nblk] (last = V) and counters[CNT_VISIBLE] = V.
__global__ void __launch_bounds__(1024) block_offsets_kernel(GeomState g)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = g.nblk, per = (n + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += g.blk_count[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += v;
        }
        s_warp[lane] = w; // inclusive over warps
    }
    __syncthreads();
    uint32_t excl = (warp ? s_warp[warp - 1] : 0u) + incl - sum;
    for (uint32_t i = lo; i < hi; i++) {
        g.blk_offset[i] = excl;
        excl += g.blk_count[i];
    }
    if (threadIdx.x == 0) {
        const uint32_t total = s_warp[31];
        g.blk_offset[n] = total;
        g.counters[CNT_VISIBLE] = total;
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One thread per VISIBLE Gaussian (rank r in Gaussian-id order -> slot -> id): the heavy chain rule runs on dense warps
// (only ~20% of the Gaussians of a view are visible; a thread-per-Gaussian layout would execute it with ~80% idle lanes).
constexpr int BWD_THREADS = 64; // a 256-Gaussian block has ~50 visible ones at 20 % visibility: two dense warps
constexpr int SH_ROW_STRIDE = 49; // 48 floats + 1: lane rows start in different banks

// The whole chain rule of one visible Gaussian (backward.cu:144-412 = K8 + K9, plus the fused activations' backward). Called by a
// full warp: `visible` marks the lanes that carry a Gaussian, `nrows` (warp-uniform) is how many of the warp's lanes do, `r` is the
// Gaussian's rank among the visible ones (= its packet index), s_sh_warp the warp's 32 x SH_ROW_STRIDE staging rows.
__device__ __forceinline__ void preprocess_bwd_one(const PreBwdArgs& a, const bool visible, const uint32_t slot, const uint32_t r,
                                                   const uint32_t nrows, const uint32_t lane, float* s_sh_warp, const float* s_view,
                                                   const float* s_proj)
{
    const int idx = visible ? (int)a.g.slot_gid[slot] : 0;
    float* my_sh = s_sh_warp + lane * SH_ROW_STRIDE;

    float3 dL_dmean = {0.f, 0.f, 0.f};
    float2 dL_dmean2D = {0.f, 0.f};
    float dL_dcov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 dL_dscale = {0.f, 0.f, 0.f};
    float4 dL_drot = {0.f, 0.f, 0.f, 0.f};
    float3 dL_dcolor = {0.f, 0.f, 0.f};
    float2 dL_dseg = {0.f, 0.f};
    float dL_dopacity = 0.f;

    if (visible) {
        const float4* gr = reinterpret_cast<const float4*>(a.grad_rec + (size_t)slot * GRAD_REC_FLOATS);
        const float4 g0 = gr[0]; // dcolor.rgb, ddepth
        const float4 g1 = gr[1]; // dseg0, dseg1, dmean2D.x, dmean2D.y
        const float4 g2 = gr[2]; // dconic.x, dconic.y, dconic.w, dopacity
        dL_dcolor = {g0.x, g0.y, g0.z};
        const float dL_ddepth = g0.w;
        dL_dseg = {g1.x, g1.y};
        dL_dmean2D = {g1.z, g1.w};
        dL_dopacity = g2.w;

        const float3 mean = {a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]};
        float3 sc = {0.f, 0.f, 0.f};
        float4 q = {0.f, 0.f, 0.f, 0.f}, q_raw = {0.f, 0.f, 0.f, 0.f};
        float cov3D[6];
        if (a.cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3D[i] = a.cov3D_precomp[6 * (size_t)idx + i];
        } else {
            sc = {a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]};
            q = *reinterpret_cast<const float4*>(a.rotations + 4 * (size_t)idx);
            if (a.raw) {
                q_raw = q;
                sc = {expf(sc.x), expf(sc.y), expf(sc.z)};
                q = act_normalize(q);
            }
            cov3d_from_scale_rot(sc, a.scale_modifier, q, cov3D); // same bits as the forward's (recomputed, not stored)
        }

        // ---- conic -> cov2D -> cov3D and the covariance path of the mean (backward.cu:159-273) ----
        const float3 dL_dconic = {g2.x, g2.y, g2.z};
        Cov2DCtx cx;
        const float3 cv = cov2d_project(mean, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, s_view, &cx);
        const float h_x = a.focal_x, h_y = a.focal_y;
        const float x_grad_mul = cx.txtz < -cx.limx || cx.txtz > cx.limx ? 0 : 1;
        const float y_grad_mul = cx.tytz < -cx.limy || cx.tytz > cx.limy ? 0 : 1;
        const M3& T = cx.T;
        const M3& Vrk = cx.Vrk;
        const M3& W = cx.W;
        const float3 t = cx.t;

        float aa = cv.x;
        float b = cv.y;
        float c = cv.z;

        float denom = aa * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - aa * c) * dL_dconic.z);
            dL_dc = denom2inv * (-aa * aa * dL_dconic.z + 2 * aa * b * dL_dconic.y + (denom - aa * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + aa * b * dL_dconic.z);

            dL_dcov3D[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
            dL_dcov3D[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
            dL_dcov3D[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);

            dL_dcov3D[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db +
                           2 * T.c[1][0] * T.c[1][1] * dL_dc;
            dL_dcov3D[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db +
                           2 * T.c[1][0] * T.c[1][2] * dL_dc;
            dL_dcov3D[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db +
                           2 * T.c[1][1] * T.c[1][2] * dL_dc;
        }

        float dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
        float dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
        float dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
        float dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
        float dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
        float dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;

        float dL_dJ00 = W.c[0][0] * dL_dT00 + W.c[0][1] * dL_dT01 + W.c[0][2] * dL_dT02;
        float dL_dJ02 = W.c[2][0] * dL_dT00 + W.c[2][1] * dL_dT01 + W.c[2][2] * dL_dT02;
        float dL_dJ11 = W.c[1][0] * dL_dT10 + W.c[1][1] * dL_dT11 + W.c[1][2] * dL_dT12;
        float dL_dJ12 = W.c[2][0] * dL_dT10 + W.c[2][1] * dL_dT11 + W.c[2][2] * dL_dT12;

        float tz = 1.f / t.z;
        float tz2 = tz * tz;
        float tz3 = tz2 * tz;

        float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;

        // transpose of the view rotation (auxiliary.h:89-97)
        dL_dmean = {s_view[0] * dL_dtx + s_view[1] * dL_dty + s_view[2] * dL_dtz,
                    s_view[4] * dL_dtx + s_view[5] * dL_dty + s_view[6] * dL_dtz,
                    s_view[8] * dL_dtx + s_view[9] * dL_dty + s_view[10] * dL_dtz};

        // ---- projection path (backward.cu:372-389) ----
        const float* proj = s_proj;
        const float* view = s_view;
        float4 m_hom = xform_point_h(mean, proj);
        float m_w = 1.0f / (m_hom.w + 0.0000001f);
        float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        float3 d1;
        d1.x = (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D.x + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D.y;
        d1.y = (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D.x + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D.y;
        d1.z = (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D.x + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D.y;
        dL_dmean.x += d1.x;
        dL_dmean.y += d1.y;
        dL_dmean.z += d1.z;

        // ---- depth path (backward.cu:394-403) ----
        float mul3 = view[2] * mean.x + view[6] * mean.y + view[10] * mean.z + view[14];
        float3 d2;
        d2.x = (view[2] - view[3] * mul3) * dL_ddepth;
        d2.y = (view[6] - view[7] * mul3) * dL_ddepth;
        d2.z = (view[10] - view[11] * mul3) * dL_ddepth;
        dL_dmean.x += d2.x;
        dL_dmean.y += d2.y;
        dL_dmean.z += d2.z;

        // ---- SH path ----
        if (a.shs != nullptr) {
            const float3 campos = {a.campos[0], a.campos[1], a.campos[2]};
            const unsigned cb = a.g.clamped[slot];
            V3* sh_row = (a.out.dL_dsh && !a.packets) ? reinterpret_cast<V3*>(my_sh) : nullptr;
            const V3* sh3 = reinterpret_cast<const V3*>(a.shs);
            const ShCoeffs<V3> sh = a.raw ? ShCoeffs<V3>{sh3 + idx, reinterpret_cast<const V3*>(a.shs_rest) + (size_t)idx * (a.M - 1)}
                                          : ShCoeffs<V3>{sh3 + (size_t)idx * a.M, sh3 + (size_t)idx * a.M + 1};
            const float3 d3 = sh_backward(a.D, mean, campos, sh, cb, V3{dL_dcolor.x, dL_dcolor.y, dL_dcolor.z}, ShRowWriter{sh_row});
            if (sh_row) // coefficients above the active degree get no gradient
                for (int k = (a.D + 1) * (a.D + 1); k < a.M; k++) sh_row[k] = V3{0.f, 0.f, 0.f};
            dL_dmean.x += d3.x;
            dL_dmean.y += d3.y;
            dL_dmean.z += d3.z;
        }
        // ---- scale / rotation path ----
        if (a.scales != nullptr) cov3d_backward(sc, a.scale_modifier, q, dL_dcov3D, dL_dscale, dL_drot);
        if (a.raw) { // chain rule of the fused activations: gradients w.r.t. the raw parameters
            dL_dscale = {dL_dscale.x * sc.x, dL_dscale.y * sc.y, dL_dscale.z * sc.z}; // d exp
            dL_drot = dact_normalize(q_raw, dL_drot);
            dL_dopacity = dact_sigmoid(act_sigmoid(a.opacities[idx]), dL_dopacity);
            if (a.S == 2 && a.segments != nullptr) {
                const float2 sg = *reinterpret_cast<const float2*>(a.segments + 2 * (size_t)idx);
                dL_dseg = {dact_sigmoid(act_sigmoid(sg.x), dL_dseg.x), dact_sigmoid(act_sigmoid(sg.y), dL_dseg.y)};
            } else if (a.S > 0 && a.segments != nullptr) {
                const float* sg = a.segments + (size_t)a.S * idx;
                dL_dseg.x = dact_sigmoid(act_sigmoid(sg[0]), dL_dseg.x);
                if (a.S > 1) dL_dseg.y = dact_sigmoid(act_sigmoid(sg[1]), dL_dseg.y);
            }
        }
    }

    // ---- SH gradient rows: each 192-B row leaves the warp as whole 128-B + 64-B bursts (a per-thread row write would
    //      touch 32 different rows per store instruction, half a sector each) ----
    if (a.packets) {
        // ---- packet mode: 16 words per visible Gaussian; the receiver rebuilds the SH rows from the colour gradient ----
        if (visible && r < a.packet_capacity) {
            float3 dRGB = dL_dcolor;
            if (a.shs != nullptr) { // the clamp mask of sh_backward (backward.cu:31-34)
                const unsigned cb = a.g.clamped[slot];
                dRGB.x *= (cb & 1u) ? 0 : 1;
                dRGB.y *= (cb & 2u) ? 0 : 1;
                dRGB.z *= (cb & 4u) ? 0 : 1;
            }
            if (a.vis_index) { // order-independent atomics: the index is deterministic
                atomicOr(&a.vis_index[2 * ((uint32_t)idx >> 5)], 1u << (idx & 31));
                atomicMax(&a.vis_index[2 * ((uint32_t)idx >> 5) + 1], ~r); // ~(smallest packet index of the group)
            }
            float4* pk = reinterpret_cast<float4*>(a.packets + (size_t)r * GSR_PACKET_WORDS); // 64-byte aligned
            pk[0] = make_float4(dRGB.x, dRGB.y, dRGB.z, dL_dmean.x);
            pk[1] = make_float4(dL_dmean.y, dL_dmean.z, dL_dopacity, dL_dseg.x);
            pk[2] = make_float4(dL_dseg.y, dL_dscale.x, dL_dscale.y, dL_dscale.z);
            pk[3] = dL_drot;
            if (a.out.dL_dmeans2D) { // per-view screen-space gradient for the densification statistics (dense, pre-zeroed)
                a.out.dL_dmeans2D[3 * (size_t)idx + 0] = dL_dmean2D.x;
                a.out.dL_dmeans2D[3 * (size_t)idx + 1] = dL_dmean2D.y;
            }
        }
        return;
    }
    if (a.out.dL_dsh) {
        __syncwarp();
        const int row_floats = a.M * 3;
        for (uint32_t rr = 0; rr < nrows; rr++) {
            const int id_rr = __shfl_sync(0xffffffffu, idx, rr);
            const float* src = s_sh_warp + rr * SH_ROW_STRIDE;
            if (a.out.dL_dsh_rest) { // raw-parameter mode: coefficient 0 -> features_dc row, the rest -> features_rest row
                float* dc = a.out.dL_dsh + (size_t)id_rr * 3;
                float* rest = a.out.dL_dsh_rest + (size_t)id_rr * (row_floats - 3) - 3;
                if (a.out.accumulate) {
                    for (int k = lane; k < row_floats; k += 32) (k < 3 ? dc : rest)[k] += src[k];
                } else {
                    for (int k = lane; k < row_floats; k += 32) (k < 3 ? dc : rest)[k] = src[k];
                }
                continue;
            }
            float* dst = a.out.dL_dsh + (size_t)id_rr * row_floats;
            if (a.out.accumulate) {
                for (int k = lane; k < row_floats; k += 32) dst[k] += src[k];
            } else {
                for (int k = lane; k < row_floats; k += 32) dst[k] = src[k];
            }
        }
    }
    if (!visible) return;

    // ---- remaining gradient rows of this visible Gaussian (rows of invisible Gaussians were zero-filled beforehand) ----
    const size_t i = (size_t)idx;
    const bool acc = a.out.accumulate != 0;
    auto put = [acc](float* p, float v) {
        if (acc) *p += v;
        else *p = v;
    };
    if (a.out.dL_dmeans3D) {
        put(a.out.dL_dmeans3D + 3 * i + 0, dL_dmean.x);
        put(a.out.dL_dmeans3D + 3 * i + 1, dL_dmean.y);
        put(a.out.dL_dmeans3D + 3 * i + 2, dL_dmean.z);
    }
    if (a.out.dL_dmeans2D) {
        put(a.out.dL_dmeans2D + 3 * i + 0, dL_dmean2D.x);
        put(a.out.dL_dmeans2D + 3 * i + 1, dL_dmean2D.y);
        if (!acc) a.out.dL_dmeans2D[3 * i + 2] = 0.f;
    }
    if (a.out.dL_dopacity) put(a.out.dL_dopacity + i, dL_dopacity);
    if (a.out.dL_dcolors) {
        put(a.out.dL_dcolors + 3 * i + 0, dL_dcolor.x);
        put(a.out.dL_dcolors + 3 * i + 1, dL_dcolor.y);
        put(a.out.dL_dcolors + 3 * i + 2, dL_dcolor.z);
    }
    if (a.out.dL_dsegments && a.S == 2) {
        put(a.out.dL_dsegments + 2 * i + 0, dL_dseg.x);
        put(a.out.dL_dsegments + 2 * i + 1, dL_dseg.y);
    } else if (a.out.dL_dsegments && a.S > 0) { // runtime class count: pair 0 from the record, the other pairs from their own buffers
        float* row = a.out.dL_dsegments + (size_t)a.S * i;
        put(row, dL_dseg.x);
        if (a.S > 1) put(row + 1, dL_dseg.y);
        for (uint32_t k = 0; k < a.g.extra_pairs; k++) {
            float2 gk = a.grad_seg_extra[((size_t)k * a.g.slots + slot)];
            const int c0 = 2 * ((int)k + 1);
            if (a.raw && a.segments != nullptr) {
                const float* sg = a.segments + (size_t)a.S * i + c0;
                gk.x = dact_sigmoid(act_sigmoid(sg[0]), gk.x);
                if (c0 + 1 < a.S) gk.y = dact_sigmoid(act_sigmoid(sg[1]), gk.y);
            }
            put(row + c0, gk.x);
            if (c0 + 1 < a.S) put(row + c0 + 1, gk.y);
        }
    }
    if (a.out.dL_dscales) {
        put(a.out.dL_dscales + 3 * i + 0, dL_dscale.x);
        put(a.out.dL_dscales + 3 * i + 1, dL_dscale.y);
        put(a.out.dL_dscales + 3 * i + 2, dL_dscale.z);
    }
    if (a.out.dL_drotations) {
        put(a.out.dL_drotations + 4 * i + 0, dL_drot.x);
        put(a.out.dL_drotations + 4 * i + 1, dL_drot.y);
        put(a.out.dL_drotations + 4 * i + 2, dL_drot.z);
        put(a.out.dL_drotations + 4 * i + 3, dL_drot.w);
    }
    if (a.out.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) put(a.out.dL_dcov3D + 6 * i + k, dL_dcov3D[k]);
    }
}

// Dense / accumulate / packet backward over the visible Gaussians, one CTA per 256-Gaussian preprocess block. The forward's slot
// layout makes that block's visible Gaussians the contiguous slots [256 b, 256 b + blk_count[b]), so there is no index list to
// chase, and the CTA OWNS the gradient rows of Gaussians [256 b, 256 b + 256): in dense mode it first zero-fills them (coalesced
// 16-byte stores) and then overwrites the visible ones, which still sit in L2 -- the API's "dense rows, zeros for invisible
// Gaussians" costs no separate 1.5 GB memset pass and nothing competes with the compositing backward for HBM.
// (96 registers, 10 CTAs of 64 threads per SM. Forcing 12 / 16 CTAs with launch bounds measured 0.350 / 0.478 ms against 0.351 ms at
// cfg3: the spills cost what the occupancy buys.)
__global__ void __launch_bounds__(BWD_THREADS) preprocess_bwd_kernel(const PreBwdArgs a)
{
    __shared__ float s_view[16], s_proj[16];
    __shared__ float s_sh[BWD_THREADS / 32][32 * SH_ROW_STRIDE]; // per warp: 32 SH gradient rows, transposed out by the body
    if (threadIdx.x < 16) {
        s_view[threadIdx.x] = a.view[threadIdx.x];
        s_proj[threadIdx.x] = a.proj[threadIdx.x];
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x;
    const uint32_t cnt = a.g.blk_count[b], base = a.g.blk_offset[b];
    if (a.fill) { // dense mode: this CTA's 256 rows of every output tensor
        const size_t lo = (size_t)b * PRE_BLOCK, hi = min((size_t)a.P, lo + PRE_BLOCK);
        const int Mdc = a.out.dL_dsh_rest ? 1 : a.M;
        const struct { float* p; int row; } t[10] = {
            {a.out.dL_dsh, Mdc * 3}, {a.out.dL_dsh_rest, (a.M - 1) * 3}, {a.out.dL_dmeans3D, 3}, {a.out.dL_dmeans2D, 3}, {a.out.dL_dopacity, 1},
            {a.out.dL_dcolors, 3}, {a.out.dL_dsegments, a.S > 0 ? a.S : 0}, {a.out.dL_dscales, 3}, {a.out.dL_drotations, 4}, {a.out.dL_dcov3D, 6}};
#pragma unroll
        for (int k = 0; k < 10; k++) {
            if (!t[k].p || t[k].row <= 0) continue;
            float* dst = t[k].p + lo * t[k].row; // 256 * row * 4 bytes per block: always 16-byte aligned
            const size_t n = (hi - lo) * t[k].row, n4 = n >> 2;
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (size_t i = threadIdx.x; i < n4; i += BWD_THREADS) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += BWD_THREADS) dst[i] = 0.f;
        }
    }
    if (a.packets && b == 0 && threadIdx.x == 0) *a.packet_count = a.g.counters[CNT_VISIBLE];
    __syncthreads(); // view matrices staged; zero rows ordered before this CTA's visible rows
    for (uint32_t j0 = 0; j0 < cnt; j0 += BWD_THREADS) {
        const uint32_t j = j0 + threadIdx.x;
        if ((j & ~31u) >= cnt) continue; // whole warp past the end
        __syncwarp();                    // the previous trip's row transposition is done with the staging rows
        preprocess_bwd_one(a, j < cnt, b * PRE_BLOCK + min(j, cnt - 1), base + j, min(32u, cnt - (j & ~31u)), lane, s_sh[warp], s_view, s_proj);
    }
}

// Multi-GPU gradient rebuild, gather form: one thread per Gaussian sums the packets of all views that saw it and writes its dense
// rows once. A warp owns 32 consecutive Gaussians = one word of every view's visibility index, so "which views, which packet"
// costs two warp-uniform word loads per view, and the packets a warp needs from one view are adjacent in memory. Compared with
// one read-modify-write pass per view (apply_packets_kernel) every dense row is touched exactly once and no zero fill is needed.
constexpr int GATHER_THREADS = 128;
constexpr int GATHER_GROUP = 8; // views whose index words are loaded together (one lane per view)
constexpr int GATHER_CAP = 64;  // packets a warp stages at a time (8 views x 32 Gaussians x ~20% visible = ~51)
constexpr int GATHER_STRIDE = 20; // words between staged packets: 16-byte aligned, and 8 consecutive packets cover all 32 banks
// words between the SH rows staged for the store: 16-byte aligned rows whose 16-byte quarters fall into 8 different bank groups for
// 8 consecutive lanes (52 l mod 32 = 0, 20, 8, 28, 16, 4, 24, 12), so both the per-lane 128-bit row writes and the 128-bit reads
// of consecutive output quarters are conflict free
constexpr int GATHER_ROW_STRIDE = 52;

// 16-byte asynchronous copy, L2 only (.cg): packets are read once, and for a peer blob the request crosses NVLink as one
// 16-byte read per lane, 64 contiguous bytes per packet (round 1 moved the 68-byte packets as 4-byte .ca copies)
__device__ __forceinline__ void cp_async_16(uint32_t* smem_dst, const uint32_t* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A warp owns 32 consecutive Gaussians = one pair of every view's visibility index. Per group of 8 views:
//   1. lane u loads view u's index pair (one load instruction for the group); a shuffle scan gives every view's packet count/offset;
//   2. the packets the warp needs from one view are ADJACENT (ascending Gaussian order), so each view's span is copied with fully
//      coalesced asynchronous 128-byte requests into a staging buffer -- the right shape for loads that cross NVLink;
//   3. every lane walks ITS OWN list of views (bit mask), so a trip of the loop serves up to 32 (Gaussian, view) pairs whatever
//      views they belong to: ~4.5 trips per group instead of one 20%-occupied trip per view.
// Views are summed in ascending order per Gaussian on every rank, so replicas end up bitwise identical.
__global__ void __launch_bounds__(GATHER_THREADS, 4) gather_packets_kernel(const GatherPacketsArgs a)
{
    __shared__ __align__(16) uint32_t s_buf[GATHER_THREADS / 32][32 * GATHER_ROW_STRIDE]; // staging, then the SH rows for the coalesced store
    __shared__ float s_cam[GSR_MAX_GATHER_VIEWS * 3];
    static_assert(GATHER_CAP * GATHER_STRIDE <= 32 * GATHER_ROW_STRIDE && GSR_PACKET_WORDS == 16, "staging must fit");
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.num_views * 3; i += GATHER_THREADS) s_cam[i] = a.campos[i];
    __syncthreads();
    const uint32_t W = ((uint32_t)a.P + 31u) >> 5; // groups of 32 Gaussians = index pairs per view
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool want_sh = a.out.dL_dsh && a.M > 0;
    uint32_t* stage = s_buf[warp];
    const uint32_t nvg0 = (uint32_t)min(GATHER_GROUP, a.num_views);
    // Persistent warps: warp w takes groups w, w + (warps in the grid), ... The index pairs of the NEXT group are requested before
    // the current one is processed, so the (possibly remote) index round trip is off the critical path of every group but the first.
    const uint32_t warps_total = gridDim.x * (GATHER_THREADS / 32);
    uint32_t grp = blockIdx.x * (GATHER_THREADS / 32) + warp;
    uint2 pr_next = make_uint2(0u, 0u);
    if (grp < W && lane < nvg0) pr_next = __ldg(reinterpret_cast<const uint2*>(a.views[lane] + a.index_off) + grp);
    for (; grp < W; grp += warps_total) {
    const uint2 pr_first = pr_next;
    if (grp + warps_total < W && lane < nvg0) pr_next = __ldg(reinterpret_cast<const uint2*>(a.views[lane] + a.index_off) + grp + warps_total);
    const uint32_t id = grp * 32u + lane, word_i = grp;
    const bool valid = id < (uint32_t)a.P;
    float acc[13], dsh[48];
#pragma unroll
    for (int k = 0; k < 13; k++) acc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 48; k++) dsh[k] = 0.f;
    const size_t i = (size_t)(valid ? id : 0);
    const float3 pos = {a.means3D[3 * i], a.means3D[3 * i + 1], a.means3D[3 * i + 2]};
    for (int g0 = 0; g0 < a.num_views; g0 += GATHER_GROUP) {
        const uint32_t nvg = (uint32_t)min(GATHER_GROUP, a.num_views - g0);
        uint32_t my_bits = 0u, my_first = 0u;
        if (lane < nvg) {
            const uint2 pr = g0 == 0 ? pr_first : __ldg(reinterpret_cast<const uint2*>(a.views[g0 + lane] + a.index_off) + word_i);
            const uint32_t cnt = __popc(pr.x), f0 = ~pr.y;
            const bool ok = f0 <= a.capacity && cnt <= a.capacity - f0; // a corrupt index cannot make a span leave its blob
            my_bits = ok ? pr.x : 0u;
            my_first = f0;
        }
        const uint32_t my_cnt = __popc(my_bits);
        uint32_t incl = my_cnt; // inclusive scan over the (at most 8) view lanes
#pragma unroll
        for (int d = 1; d < GATHER_GROUP; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (__shfl_sync(0xffffffffu, incl, nvg - 1) == 0u) continue; // none of the 32 Gaussians is visible in these views
        uint32_t u0 = 0u, base = 0u;
        while (u0 < nvg) { // usually one segment: all views of the group fit the staging buffer
            const uint32_t fit = __ballot_sync(0xffffffffu, lane >= u0 && lane < nvg && incl - base <= GATHER_CAP);
            const uint32_t e = u0 + __popc(fit); // >= u0 + 1: one view never has more than 32 packets
            __syncwarp();                        // the previous segment's readers are done with the staging buffer
            uint32_t m = 0u;                     // views of this segment that see my Gaussian
            for (uint32_t u = u0; u < e; u++) {
                const uint32_t b_u = __shfl_sync(0xffffffffu, my_bits, u);
                const uint32_t nquads = __popc(b_u) * (GSR_PACKET_WORDS / 4); // 16-byte quarters of this view's adjacent packets
                if (nquads == 0u) continue;
                const uint32_t f_u = __shfl_sync(0xffffffffu, my_first, u);
                const uint32_t off_u = __shfl_sync(0xffffffffu, incl - my_cnt, u) - base;
                const uint32_t* src = a.views[g0 + u] + a.packet_off + (size_t)f_u * GSR_PACKET_WORDS;
                uint32_t* dst = stage + off_u * GATHER_STRIDE;
                for (uint32_t qd = lane; qd < nquads; qd += 32) cp_async_16(dst + (qd >> 2) * GATHER_STRIDE + (qd & 3u) * 4u, src + qd * 4u);
                m |= ((b_u >> lane) & 1u) << u;
            }
            if (!valid) m = 0u;
            cp_async_wait_all();
            __syncwarp();
            while (__any_sync(0xffffffffu, m != 0u)) {
                const uint32_t u = m ? (uint32_t)__ffs(m) - 1u : 0u;
                const uint32_t b_u = __shfl_sync(0xffffffffu, my_bits, u);
                const uint32_t off_u = __shfl_sync(0xffffffffu, incl - my_cnt, u) - base;
                if (m) {
                    m &= m - 1u;
                    const float4* pk = reinterpret_cast<const float4*>(stage + (off_u + __popc(b_u & lt_mask)) * GATHER_STRIDE);
                    const float4 p0 = pk[0], p1 = pk[1], p2 = pk[2], p3 = pk[3];
                    const float f[13] = {p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, p3.x, p3.y, p3.z, p3.w};
#pragma unroll
                    for (int k = 0; k < 13; k++) acc[k] += f[k];
                    if (want_sh) {
                        const int r = g0 + (int)u;
                        const float cr = p0.x, cg = p0.y, cb = p0.z;
                        V3 dir_orig = {pos.x - s_cam[3 * r], pos.y - s_cam[3 * r + 1], pos.z - s_cam[3 * r + 2]};
                        const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
                        float w[16];
                        sh_basis(a.D, dir_orig.x / len, dir_orig.y / len, dir_orig.z / len, w);
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            dsh[3 * k + 0] += w[k] * cr;
                            dsh[3 * k + 1] += w[k] * cg;
                            dsh[3 * k + 2] += w[k] * cb;
                        }
                    }
                }
            }
            base = __shfl_sync(0xffffffffu, incl, e - 1);
            u0 = e;
        }
    }
    if (valid) {
        if (a.out.dL_dmeans3D) {
            a.out.dL_dmeans3D[3 * i + 0] = acc[0];
            a.out.dL_dmeans3D[3 * i + 1] = acc[1];
            a.out.dL_dmeans3D[3 * i + 2] = acc[2];
        }
        if (a.out.dL_dopacity) a.out.dL_dopacity[i] = acc[3];
        if (a.out.dL_dsegments && a.S == 2) {
            a.out.dL_dsegments[2 * i + 0] = acc[4];
            a.out.dL_dsegments[2 * i + 1] = acc[5];
        }
        if (a.out.dL_dscales) {
            a.out.dL_dscales[3 * i + 0] = acc[6];
            a.out.dL_dscales[3 * i + 1] = acc[7];
            a.out.dL_dscales[3 * i + 2] = acc[8];
        }
        if (a.out.dL_drotations) *reinterpret_cast<float4*>(a.out.dL_drotations + 4 * i) = {acc[9], acc[10], acc[11], acc[12]};
    }
    if (want_sh) { // 32 consecutive rows of the warp form one contiguous span: transposed through shared memory, fully coalesced
        __syncwarp();
        float* rows = reinterpret_cast<float*>(stage);
        float4* my4 = reinterpret_cast<float4*>(rows + lane * GATHER_ROW_STRIDE);
#pragma unroll
        for (int k = 0; k < 12; k++) my4[k] = make_float4(dsh[4 * k], dsh[4 * k + 1], dsh[4 * k + 2], dsh[4 * k + 3]);
        __syncwarp();
        const uint32_t first_row = grp * 32u;
        if (first_row < (uint32_t)a.P) {
            const uint32_t nrows = min(32u, (uint32_t)a.P - first_row);
            const int row_floats = a.M * 3;
            float* dst = a.out.dL_dsh + (size_t)first_row * row_floats;
            const uint32_t total = nrows * (uint32_t)row_floats;
            if (a.out.dL_dsh_rest) { // raw-parameter mode: two tensors, each span still contiguous for the warp's 32 rows
                float* dc = a.out.dL_dsh + (size_t)first_row * 3;
                float* rest = a.out.dL_dsh_rest + (size_t)first_row * (row_floats - 3);
                for (uint32_t e = lane; e < nrows * 3u; e += 32) dc[e] = rows[(e / 3u) * GATHER_ROW_STRIDE + e % 3u];
                const uint32_t rf = (uint32_t)row_floats - 3u;
                for (uint32_t e = lane; e < nrows * rf; e += 32) {
                    const uint32_t rr = e / rf, k = e - rr * rf + 3u;
                    rest[e] = k < 48 ? rows[rr * GATHER_ROW_STRIDE + k] : 0.f;
                }
            } else if (row_floats == 48 && ((size_t)a.out.dL_dsh & 15u) == 0) {
                // the common layout: the warp's 32 rows are one contiguous 6 KB span, moved as 384 16-byte quarters (12 trips of
                // LDS.128 + STG.128 instead of 48 trips of 4-byte ones)
                float4* dst4 = reinterpret_cast<float4*>(dst);
                const uint32_t total4 = nrows * 12u;
#pragma unroll 4
                for (uint32_t e = lane; e < total4; e += 32) {
                    const uint32_t rr = e / 12u, k4 = e - rr * 12u;
                    dst4[e] = *reinterpret_cast<const float4*>(rows + rr * GATHER_ROW_STRIDE + 4u * k4);
                }
            } else if (row_floats == 48) {
                for (uint32_t e = lane; e < total; e += 32) dst[e] = rows[(e / 48u) * GATHER_ROW_STRIDE + e % 48u];
            } else {
                for (uint32_t e = lane; e < total; e += 32) {
                    const uint32_t rr = e / (uint32_t)row_floats, k = e - rr * (uint32_t)row_floats;
                    dst[e] = k < 48 ? rows[rr * GATHER_ROW_STRIDE + k] : 0.f;
                }
            }
        }
    }
    __syncwarp(); // the row store is done with the staging buffer before the next group stages into it
    } // groups
}

// ---- the same pass as a software pipeline (GSR_GATHER_PIPE=1) ----
// While a warp sums the packets of group g out of one staging buffer, the asynchronous copies of its group g + 1 are already
// landing in the other one, and the index pairs and positions of the group after that are in flight: a warp's (local or NVLink)
// round trips overlap its own arithmetic and row stores instead of alternating with them. Only the first segment (the first <= 64
// packets) of a group's first 8 views is prefetched -- the common case; anything beyond is copied synchronously as above.
constexpr int GATHER_STAGE_WORDS = 32 * GATHER_ROW_STRIDE;
constexpr int GATHER_PIPE_SMEM_BYTES = (GATHER_THREADS / 32) * 2 * GATHER_STAGE_WORDS * (int)sizeof(uint32_t);

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct GatherViews // what lane u < 8 knows about view g0 + u of the warp's 32 Gaussians
{
    uint32_t bits, first, cnt, incl; // visible bits, first packet, packet count, inclusive scan of the counts over the view lanes
};
__device__ __forceinline__ GatherViews gather_views(const uint2 pr, const bool has_view, const uint32_t capacity, const uint32_t lane)
{
    GatherViews g;
    g.bits = 0u;
    g.first = 0u;
    if (has_view) {
        const uint32_t cnt = __popc(pr.x), f0 = ~pr.y;
        const bool ok = f0 <= capacity && cnt <= capacity - f0; // a corrupt index cannot make a span leave its blob
        g.bits = ok ? pr.x : 0u;
        g.first = f0;
    }
    g.cnt = __popc(g.bits);
    g.incl = g.cnt;
#pragma unroll
    for (int d = 1; d < GATHER_GROUP; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, g.incl, d);
        if (lane >= (uint32_t)d) g.incl += t;
    }
    return g;
}
// Issue the copies of the views [u0, e) of a view group -- the longest run whose packets fit the staging buffer. Returns e
// (>= u0 + 1: one view never has more than 32 packets); m = the views of the run that see this lane's Gaussian.
__device__ __forceinline__ uint32_t gather_issue(const GatherPacketsArgs& a, const int g0, const uint32_t nvg, const GatherViews& g, const uint32_t u0,
                                                 const uint32_t base, uint32_t* stage, const uint32_t lane, uint32_t& m)
{
    const uint32_t fit = __ballot_sync(0xffffffffu, lane >= u0 && lane < nvg && g.incl - base <= GATHER_CAP);
    const uint32_t e = u0 + __popc(fit);
    m = 0u;
    for (uint32_t u = u0; u < e; u++) {
        const uint32_t b_u = __shfl_sync(0xffffffffu, g.bits, u);
        const uint32_t nquads = __popc(b_u) * (GSR_PACKET_WORDS / 4); // 16-byte quarters of this view's adjacent packets
        if (nquads == 0u) continue;
        const uint32_t f_u = __shfl_sync(0xffffffffu, g.first, u);
        const uint32_t off_u = __shfl_sync(0xffffffffu, g.incl - g.cnt, u) - base;
        const uint32_t* src = a.views[g0 + u] + a.packet_off + (size_t)f_u * GSR_PACKET_WORDS;
        uint32_t* dst = stage + off_u * GATHER_STRIDE;
        for (uint32_t qd = lane; qd < nquads; qd += 32) cp_async_16(dst + (qd >> 2) * GATHER_STRIDE + (qd & 3u) * 4u, src + qd * 4u);
        m |= ((b_u >> lane) & 1u) << u;
    }
    return e;
}

__global__ void __launch_bounds__(GATHER_THREADS, 4) gather_packets_pipe_kernel(const GatherPacketsArgs a)
{
    extern __shared__ __align__(16) uint32_t s_gather[]; // [warps][2][GATHER_STAGE_WORDS]
    __shared__ float s_cam[GSR_MAX_GATHER_VIEWS * 3];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < a.num_views * 3; i += GATHER_THREADS) s_cam[i] = a.campos[i];
    __syncthreads();
    const uint32_t W = ((uint32_t)a.P + 31u) >> 5; // groups of 32 Gaussians = index pairs per view
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool want_sh = a.out.dL_dsh && a.M > 0;
    uint32_t* const stage0 = s_gather + warp * 2 * GATHER_STAGE_WORDS;
    const uint32_t nvg0 = (uint32_t)min(GATHER_GROUP, a.num_views);
    const uint32_t stride = gridDim.x * (GATHER_THREADS / 32);
    uint32_t grp = blockIdx.x * (GATHER_THREADS / 32) + warp;
    if (grp >= W) return;

    auto load_pair = [&](uint32_t g) {
        uint2 pr = make_uint2(0u, 0u);
        if (g < W && lane < nvg0) pr = __ldg(reinterpret_cast<const uint2*>(a.views[lane] + a.index_off) + g);
        return pr;
    };
    auto load_pos = [&](uint32_t g) {
        const uint32_t id = g * 32u + lane;
        const size_t i = (size_t)(g < W && id < (uint32_t)a.P ? id : 0u);
        return make_float3(a.means3D[3 * i], a.means3D[3 * i + 1], a.means3D[3 * i + 2]);
    };

    // prologue: this warp's first group is staged, the second one's index pairs and positions are requested
    uint32_t cb = 0u;
    GatherViews gc = gather_views(load_pair(grp), lane < nvg0, a.capacity, lane);
    float3 pos = load_pos(grp);
    uint32_t m_c = 0u, e_c = 0u;
    if (__shfl_sync(0xffffffffu, gc.incl, nvg0 - 1) != 0u) e_c = gather_issue(a, 0, nvg0, gc, 0u, 0u, stage0, lane, m_c);
    cp_async_commit();
    uint2 pr_n = load_pair(grp + stride);
    float3 pos_n = load_pos(grp + stride);

    for (; grp < W; grp += stride) {
        uint32_t* const stage = stage0 + cb * GATHER_STAGE_WORDS;
        // ---- next group: its first segment starts travelling into the other buffer; the group after it is requested ----
        const GatherViews gn = gather_views(pr_n, lane < nvg0 && grp + stride < W, a.capacity, lane);
        uint32_t m_n = 0u, e_n = 0u;
        if (__shfl_sync(0xffffffffu, gn.incl, nvg0 - 1) != 0u) e_n = gather_issue(a, 0, nvg0, gn, 0u, 0u, stage0 + (cb ^ 1u) * GATHER_STAGE_WORDS, lane, m_n);
        cp_async_commit();
        pr_n = load_pair(grp + 2 * stride);
        const float3 pos_nn = load_pos(grp + 2 * stride);
        cp_async_wait_group<1>(); // everything but the copies just issued: this group's first segment has landed
        __syncwarp();

        const uint32_t id = grp * 32u + lane;
        const bool valid = id < (uint32_t)a.P;
        float acc[13], dsh[48];
#pragma unroll
        for (int k = 0; k < 13; k++) acc[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 48; k++) dsh[k] = 0.f;
        const size_t i = (size_t)(valid ? id : 0);
        for (int g0 = 0; g0 < a.num_views; g0 += GATHER_GROUP) {
            const uint32_t nvg = (uint32_t)min(GATHER_GROUP, a.num_views - g0);
            GatherViews g = gc;
            uint32_t m = m_c, e = e_c;
            bool staged = true; // the first segment of the first 8 views is already in `stage`
            if (g0 > 0) {
                uint2 pr = make_uint2(0u, 0u);
                if (lane < nvg) pr = __ldg(reinterpret_cast<const uint2*>(a.views[g0 + lane] + a.index_off) + grp);
                g = gather_views(pr, lane < nvg, a.capacity, lane);
                staged = false;
            }
            if (__shfl_sync(0xffffffffu, g.incl, nvg - 1) == 0u) continue; // none of the 32 Gaussians is visible in these views
            uint32_t u0 = 0u, base = 0u;
            while (u0 < nvg) { // usually one segment: all views of the group fit the staging buffer
                if (!staged) {
                    __syncwarp(); // the previous segment's readers are done with the staging buffer
                    e = gather_issue(a, g0, nvg, g, u0, base, stage, lane, m);
                    cp_async_commit();
                    cp_async_wait_group<0>();
                    __syncwarp();
                }
                staged = false;
                if (!valid) m = 0u;
                while (__any_sync(0xffffffffu, m != 0u)) {
                    const uint32_t u = m ? (uint32_t)__ffs(m) - 1u : 0u;
                    const uint32_t b_u = __shfl_sync(0xffffffffu, g.bits, u);
                    const uint32_t off_u = __shfl_sync(0xffffffffu, g.incl - g.cnt, u) - base;
                    if (m) {
                        m &= m - 1u;
                        const float4* pk = reinterpret_cast<const float4*>(stage + (off_u + __popc(b_u & lt_mask)) * GATHER_STRIDE);
                        const float4 p0 = pk[0], p1 = pk[1], p2 = pk[2], p3 = pk[3];
                        const float f[13] = {p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, p3.x, p3.y, p3.z, p3.w};
#pragma unroll
                        for (int k = 0; k < 13; k++) acc[k] += f[k];
                        if (want_sh) {
                            const int r = g0 + (int)u;
                            const float cr = p0.x, cg = p0.y, cb_ = p0.z;
                            V3 dir_orig = {pos.x - s_cam[3 * r], pos.y - s_cam[3 * r + 1], pos.z - s_cam[3 * r + 2]};
                            const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
                            float w[16];
                            sh_basis(a.D, dir_orig.x / len, dir_orig.y / len, dir_orig.z / len, w);
#pragma unroll
                            for (int k = 0; k < 16; k++) {
                                dsh[3 * k + 0] += w[k] * cr;
                                dsh[3 * k + 1] += w[k] * cg;
                                dsh[3 * k + 2] += w[k] * cb_;
                            }
                        }
                    }
                }
                base = __shfl_sync(0xffffffffu, g.incl, e - 1);
                u0 = e;
            }
        }
        if (valid) {
            if (a.out.dL_dmeans3D) {
                a.out.dL_dmeans3D[3 * i + 0] = acc[0];
                a.out.dL_dmeans3D[3 * i + 1] = acc[1];
                a.out.dL_dmeans3D[3 * i + 2] = acc[2];
            }
            if (a.out.dL_dopacity) a.out.dL_dopacity[i] = acc[3];
            if (a.out.dL_dsegments && a.S == 2) {
                a.out.dL_dsegments[2 * i + 0] = acc[4];
                a.out.dL_dsegments[2 * i + 1] = acc[5];
            }
            if (a.out.dL_dscales) {
                a.out.dL_dscales[3 * i + 0] = acc[6];
                a.out.dL_dscales[3 * i + 1] = acc[7];
                a.out.dL_dscales[3 * i + 2] = acc[8];
            }
            if (a.out.dL_drotations) *reinterpret_cast<float4*>(a.out.dL_drotations + 4 * i) = {acc[9], acc[10], acc[11], acc[12]};
        }
        if (want_sh) { // 32 consecutive rows of the warp form one contiguous span: transposed through shared memory, fully coalesced
            __syncwarp();
            float* rows = reinterpret_cast<float*>(stage);
            float4* my4 = reinterpret_cast<float4*>(rows + lane * GATHER_ROW_STRIDE);
#pragma unroll
            for (int k = 0; k < 12; k++) my4[k] = make_float4(dsh[4 * k], dsh[4 * k + 1], dsh[4 * k + 2], dsh[4 * k + 3]);
            __syncwarp();
            const uint32_t first_row = grp * 32u;
            const uint32_t nrows = min(32u, (uint32_t)a.P - first_row);
            const int row_floats = a.M * 3;
            float* dst = a.out.dL_dsh + (size_t)first_row * row_floats;
            const uint32_t total = nrows * (uint32_t)row_floats;
            if (a.out.dL_dsh_rest) { // raw-parameter mode: two tensors, each span still contiguous for the warp's 32 rows
                float* dc = a.out.dL_dsh + (size_t)first_row * 3;
                float* rest = a.out.dL_dsh_rest + (size_t)first_row * (row_floats - 3);
                for (uint32_t e2 = lane; e2 < nrows * 3u; e2 += 32) dc[e2] = rows[(e2 / 3u) * GATHER_ROW_STRIDE + e2 % 3u];
                const uint32_t rf = (uint32_t)row_floats - 3u;
                for (uint32_t e2 = lane; e2 < nrows * rf; e2 += 32) {
                    const uint32_t rr = e2 / rf, k = e2 - rr * rf + 3u;
                    rest[e2] = k < 48 ? rows[rr * GATHER_ROW_STRIDE + k] : 0.f;
                }
            } else if (row_floats == 48 && ((size_t)a.out.dL_dsh & 15u) == 0) {
                float4* dst4 = reinterpret_cast<float4*>(dst);
                const uint32_t total4 = nrows * 12u;
#pragma unroll 4
                for (uint32_t e2 = lane; e2 < total4; e2 += 32) {
                    const uint32_t rr = e2 / 12u, k4 = e2 - rr * 12u;
                    dst4[e2] = *reinterpret_cast<const float4*>(rows + rr * GATHER_ROW_STRIDE + 4u * k4);
                }
            } else if (row_floats == 48) {
                for (uint32_t e2 = lane; e2 < total; e2 += 32) dst[e2] = rows[(e2 / 48u) * GATHER_ROW_STRIDE + e2 % 48u];
            } else {
                for (uint32_t e2 = lane; e2 < total; e2 += 32) {
                    const uint32_t rr = e2 / (uint32_t)row_floats, k = e2 - rr * (uint32_t)row_floats;
                    dst[e2] = k < 48 ? rows[rr * GATHER_ROW_STRIDE + k] : 0.f;
                }
            }
        }
        __syncwarp(); // the row store is done with this buffer before the group after next stages into it
        gc = gn;
        m_c = m_n;
        e_c = e_n;
        pos = pos_n;
        pos_n = pos_nn;
        cb ^= 1u;
    }
    cp_async_wait_group<0>();
}

// Zero the gradient records of the live slots only: one warp per slot-block, 3 x 16 bytes per visible Gaussian, contiguous.
__global__ void __launch_bounds__(256) zero_grad_rec_kernel(GeomState g, float4* __restrict__ grad_rec)
{
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= g.nblk) return;
    const uint32_t n4 = g.blk_count[b] * (GRAD_REC_FLOATS / 4);
    float4* dst = grad_rec + (size_t)b * PRE_BLOCK * (GRAD_REC_FLOATS / 4);
    for (uint32_t i = threadIdx.x & 31u; i < n4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                                           uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    const float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    const float z = view[2] * p.x + view[6] * p.y + view[10] * p.z + view[14];
    present[idx] = z > 0.2f ? 1 : 0;
}
} // namespace

int launch_preprocess_fwd(const PreFwdArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return 0;
    static const bool hoist = !(getenv("GSR_PRE_HOIST") && atoi(getenv("GSR_PRE_HOIST")) == 0); // GSR_PRE_HOIST=0: round-1 load order (A/B)
    if (hoist) preprocess_fwd_kernel<true><<<a.g.nblk, PRE_BLOCK, 0, s>>>(a);
    else preprocess_fwd_kernel<false><<<a.g.nblk, PRE_BLOCK, 0, s>>>(a);
    count_launches(1);
    return 0;
}
int launch_block_offsets(const GeomState& g, cudaStream_t s)
{
    block_offsets_kernel<<<1, 1024, 0, s>>>(g); count_launches(1);
    return 0;
}
// Dense gradient tensors are an API requirement (zeros for invisible Gaussians). Two ways, A/B with GSR_FILL_MODE:
//   memset (default): whole-tensor cudaMemsetAsync fills (7.5 TB/s) that the caller runs on a side stream while the issue-bound
//                     compositing backward executes; the dense pass then overwrites the ~20 % visible rows;
//   kernel:           every CTA of preprocess_bwd_kernel zero-fills the 256 rows it owns before writing its visible rows.
static bool fill_in_kernel()
{
    static const bool v = getenv("GSR_FILL_MODE") && !strcmp(getenv("GSR_FILL_MODE"), "kernel");
    return v;
}
// Zero fill of up to 10 dense gradient tensors in ONE launch, written for running BESIDE the compositing backward:
//   * streaming stores (st.global.cs, evict-first): 1.46 GB of zeros do not push the backward's working set -- the per-tile lists,
//     the 48-byte records and the gradient records it accumulates into, all served from the 126 MB L2 -- out of the cache
//     (cudaMemsetAsync fills at 7.5 TB/s alone, but beside it the compositing backward ran 1.02 ms instead of 0.86 ms);
//   * a small persistent grid (GSR_FILL_CTAS, default one 64-thread CTA per two SMs, ~3.9 TB/s): the fill has the whole duration of
//     the backward to finish and should take as few issue slots and as little of the register file as possible.
// Measured at cfg3 (render_bwd + preprocess_bwd incl. the join, ms): memsets on the side stream 1.396, on the main stream 1.462;
// this kernel with 16 / 37 / 74 / 148 / 296 / 1184 CTAs 1.918 (exposed) / 1.376 / 1.377 / 1.423 / 1.428 / 1.443.
struct FillSegs
{
    float4* p[10];
    unsigned long long n16[10]; // 16-byte units
    int count;
};
__global__ void __launch_bounds__(64) fill_zero_kernel(const FillSegs f)
{
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned long long stride = (unsigned long long)gridDim.x * 64ull;
    for (int k = 0; k < f.count; k++) {
        float4* p = f.p[k];
        const unsigned long long n = f.n16[k];
        unsigned long long i = (unsigned long long)blockIdx.x * 64ull + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) { // four independent stores in flight per thread
            __stcs(p + i, z);
            __stcs(p + i + stride, z);
            __stcs(p + i + 2 * stride, z);
            __stcs(p + i + 3 * stride, z);
        }
        for (; i < n; i += stride) __stcs(p + i, z);
    }
}

int launch_grad_fills(const PreBwdArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return 0;
    const size_t P = (size_t)a.P;
    const int Mdc = a.out.dL_dsh_rest ? 1 : a.M;
    struct { float* p; size_t floats; } fills[] = {
        {a.out.dL_dsh, P * (size_t)Mdc * 3}, {a.out.dL_dsh_rest, P * (size_t)(a.M > 1 ? a.M - 1 : 0) * 3}, {a.out.dL_dmeans3D, P * 3},
        {a.out.dL_dmeans2D, P * 3}, {a.out.dL_dopacity, P}, {a.out.dL_dcolors, P * 3}, {a.out.dL_dsegments, P * (size_t)(a.S > 0 ? a.S : 0)},
        {a.out.dL_dscales, P * 3}, {a.out.dL_drotations, P * 4}, {a.out.dL_dcov3D, P * 6}};
    if (a.packets) {
        if (a.out.dL_dmeans2D) GSR_CUDA(cudaMemsetAsync(a.out.dL_dmeans2D, 0, P * 3 * sizeof(float), s));
        if (a.vis_index) {
            const size_t W = (P + 31) / 32;
            GSR_CUDA(cudaMemsetAsync(a.vis_index, 0, 2 * W * sizeof(uint32_t), s));
        }
    } else if (!a.out.accumulate && (!fill_in_kernel() || a.has_subset)) {
        // GSR_FILL_KERNEL=0: cudaMemsetAsync per tensor (round 1); default: one fill_zero_kernel launch
        static const bool own = !(getenv("GSR_FILL_KERNEL") && atoi(getenv("GSR_FILL_KERNEL")) == 0);
        static const int ctas_env = getenv("GSR_FILL_CTAS") ? atoi(getenv("GSR_FILL_CTAS")) : 0;
        FillSegs fs;
        fs.count = 0;
        for (auto& f : fills) {
            if (!f.p || !f.floats) continue;
            const size_t bytes = f.floats * sizeof(float);
            const bool aligned = ((size_t)f.p & 15u) == 0;
            const size_t body = own && aligned ? bytes & ~(size_t)15 : 0;
            if (body) {
                fs.p[fs.count] = reinterpret_cast<float4*>(f.p);
                fs.n16[fs.count] = body / 16;
                fs.count++;
            }
            if (body < bytes) GSR_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(f.p) + body, 0, bytes - body, s)); // unaligned / tail
        }
        if (fs.count) {
            static int sms = 0;
            if (!sms) {
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (sms <= 0) sms = 148;
            }
            const int grid = ctas_env > 0 ? ctas_env : (sms + 1) / 2;
            fill_zero_kernel<<<grid, 64, 0, s>>>(fs); count_launches(1);
        }
    }
    return 0;
}

int launch_zero_grad_rec(const GeomState& g, float* grad_rec, cudaStream_t s)
{
    if (g.nblk == 0) return 0;
    zero_grad_rec_kernel<<<(g.nblk + 7) / 8, 256, 0, s>>>(g, reinterpret_cast<float4*>(grad_rec)); count_launches(1);
    return 0;
}

int launch_preprocess_bwd(const PreBwdArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return 0;
    PreBwdArgs b = a;
    b.fill = (!a.packets && !a.out.accumulate && fill_in_kernel() && !a.has_subset) ? 1 : 0;
    preprocess_bwd_kernel<<<a.g.nblk, BWD_THREADS, 0, s>>>(b); count_launches(1);
    return 0;
}
int launch_gather_packets(const GatherPacketsArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return 0;
    static int resident = 0; // persistent grid: 4 CTAs per SM (launch bounds), fewer when there is less work
    if (!resident) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident = 4 * (sms > 0 ? sms : 148);
    }
    const unsigned groups = ((unsigned)a.P + 31u) / 32u, want = (groups + GATHER_THREADS / 32 - 1) / (GATHER_THREADS / 32);
    // GSR_GATHER_PIPE=1: the software-pipelined form (next group's packets travel while this one is summed)
    static const bool pipe = getenv("GSR_GATHER_PIPE") && atoi(getenv("GSR_GATHER_PIPE")) != 0;
    const unsigned grid = want < (unsigned)resident ? want : (unsigned)resident;
    if (pipe) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(gather_packets_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GATHER_PIPE_SMEM_BYTES);
            attr_set = true;
        }
        gather_packets_pipe_kernel<<<grid, GATHER_THREADS, GATHER_PIPE_SMEM_BYTES, s>>>(a);
    } else {
        gather_packets_kernel<<<grid, GATHER_THREADS, 0, s>>>(a);
    }
    count_launches(1);
    return 0;
}
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s)
{
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present); count_launches(1);
    return 0;
}
} // namespace gsr
