// Fused multi-tensor Adam over the flat parameter / gradient buffers.
// Replaces torch.optim.Adam(l, lr=0.0, eps=1e-15) of scene/gaussian_model.py:166-177: seven parameter groups with their own
// learning rates, updated by ONE launch that reads p, g, m, v and writes p, m, v once (28 B per parameter; torch's foreach
// path makes ~8 passes over the same tensors). The arithmetic follows torch.optim.Adam's default (foreach) path op by op in
// fp32, with the bias corrections evaluated in double on the host like the Python code does:
//   m = m + (1-b1) * (g - m)                    torch._foreach_lerp_
//   v = v * b2 ; v = v + (1-b2) * (g * g)       _foreach_mul_, _foreach_addcmul_
//   d = sqrt(v) / sqrt(1 - b2^t) + eps          _foreach_sqrt, _foreach_div_, _foreach_add_
//   p = p + (-(lr / (1 - b1^t))) * (m / d)      _foreach_addcdiv_
#include <math.h>
#include <stdlib.h>

#include "gsr_common.cuh"

namespace gsr
{
namespace
{
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4; // float4 per thread per trip
constexpr int ADAM_MAX_GROUPS = 16;
constexpr int ADAM_TRIPS = 4;

struct AdamArgs
{
    float* p;
    const float* g;
    float* m;
    float* v;
    int num_groups;
    unsigned long long offset[ADAM_MAX_GROUPS];
    unsigned long long count[ADAM_MAX_GROUPS];
    unsigned int first_chunk[ADAM_MAX_GROUPS + 1]; // CTA -> group table
    float neg_step_size[ADAM_MAX_GROUPS];          // -(lr / bias_correction1), per group (each group has its own step count)
    float bc2_sqrt[ADAM_MAX_GROUPS];               // sqrt(1 - b2^t), per group
    float w1, b2, w2, eps;                         // 1-b1, b2, 1-b2, eps
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float nss, float bc2_sqrt)
{
    m = __fmaf_rn(a.w1, g - m, m);
    v = __fmaf_rn(a.w2, g * g, v * a.b2);
    // m == 0 (a Gaussian no view has seen yet): torch adds -step * (0 / d) = 0, p keeps its bits. Skipping it also skips the
    // special-operand slow paths of the IEEE sqrt and divisions, which otherwise cost 1.8x on a step where 80 % of the rows are zero.
    if (m != 0.f) {
        const float d = sqrtf(v) / bc2_sqrt + a.eps;
        p = __fmaf_rn(nss, m / d, p);
    }
}

template <int TRIPS, bool STREAM>
__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const AdamArgs a)
{
    constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC * TRIPS; // floats per CTA
    int grp = 0;
#pragma unroll 1
    while (grp + 1 < a.num_groups && blockIdx.x >= a.first_chunk[grp + 1]) grp++;
    const unsigned long long base = a.offset[grp], n = a.count[grp];
    const unsigned long long c0 = (unsigned long long)(blockIdx.x - a.first_chunk[grp]) * ADAM_CHUNK;
    const float nss = a.neg_step_size[grp], bcs = a.bc2_sqrt[grp];
    float* p = a.p + base;
    const float* g = a.g + base;
    float* m = a.m + base;
    float* v = a.v + base;
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15u) == 0;
    if (aligned && c0 + ADAM_CHUNK <= n) { // full chunk, 16-byte aligned: 4 x LDG.128 per array in flight per thread
        float4 P4[TRIPS], G4[TRIPS], M4[TRIPS], V4[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; t++) {
            const unsigned long long i = c0 + (unsigned long long)(t * ADAM_THREADS + threadIdx.x) * ADAM_VEC;
            if (STREAM) { // every byte is touched once per step: evict-first keeps the stream from churning L2
                P4[t] = __ldcs(reinterpret_cast<const float4*>(p + i));
                G4[t] = __ldcs(reinterpret_cast<const float4*>(g + i));
                M4[t] = __ldcs(reinterpret_cast<const float4*>(m + i));
                V4[t] = __ldcs(reinterpret_cast<const float4*>(v + i));
            } else {
                P4[t] = *reinterpret_cast<const float4*>(p + i);
                G4[t] = __ldg(reinterpret_cast<const float4*>(g + i));
                M4[t] = *reinterpret_cast<const float4*>(m + i);
                V4[t] = *reinterpret_cast<const float4*>(v + i);
            }
        }
#pragma unroll
        for (int t = 0; t < TRIPS; t++) {
            adam_one(P4[t].x, G4[t].x, M4[t].x, V4[t].x, a, nss, bcs);
            adam_one(P4[t].y, G4[t].y, M4[t].y, V4[t].y, a, nss, bcs);
            adam_one(P4[t].z, G4[t].z, M4[t].z, V4[t].z, a, nss, bcs);
            adam_one(P4[t].w, G4[t].w, M4[t].w, V4[t].w, a, nss, bcs);
            const unsigned long long i = c0 + (unsigned long long)(t * ADAM_THREADS + threadIdx.x) * ADAM_VEC;
            if (STREAM) {
                __stcs(reinterpret_cast<float4*>(p + i), P4[t]);
                __stcs(reinterpret_cast<float4*>(m + i), M4[t]);
                __stcs(reinterpret_cast<float4*>(v + i), V4[t]);
            } else {
                *reinterpret_cast<float4*>(p + i) = P4[t];
                *reinterpret_cast<float4*>(m + i) = M4[t];
                *reinterpret_cast<float4*>(v + i) = V4[t];
            }
        }
        return;
    }
    for (unsigned long long i = c0 + threadIdx.x; i < n && i < c0 + ADAM_CHUNK; i += ADAM_THREADS) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_one(pp, g[i], mm, vv, a, nss, bcs);
        p[i] = pp;
        m[i] = mm;
        v[i] = vv;
    }
}
} // namespace
} // namespace gsr

using namespace gsr;

extern "C" int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const GsrAdamGroup* groups,
                             int32_t num_groups, double beta1, double beta2, double eps, int32_t step, gsr_stream_t stream_)
{
    if (num_groups <= 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !groups || num_groups > ADAM_MAX_GROUPS || step < 1) {
        set_error("gsr_adam_step: invalid argument (1 <= num_groups <= %d, step >= 1)", ADAM_MAX_GROUPS);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    AdamArgs a;
    a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.num_groups = num_groups;
    // Python: bias_correction1 = 1 - beta1 ** step; step_size = lr / bias_correction1; bias_correction2_sqrt = sqrt(1 - beta2 ** step)
    const double b1 = beta1, b2 = beta2; // doubles end to end: (float)0.9 would turn 1 - beta1 into 0.10000002
    a.w1 = (float)(1.0 - b1); a.b2 = (float)b2; a.w2 = (float)(1.0 - b2); a.eps = (float)eps;
    // 4 x LDG.128 per array per thread; 1, 2 or 4 trips and evict-first hints all measure 1.47 ms (6.97 TB/s) at 61 x 6M floats
    const unsigned long long ADAM_CHUNK = (unsigned long long)ADAM_THREADS * ADAM_VEC * ADAM_TRIPS;
    unsigned long long chunks = 0;
    for (int i = 0; i < num_groups; i++) {
        a.offset[i] = groups[i].offset; a.count[i] = groups[i].count;
        // torch.optim.Adam keeps `step` per parameter: a group that skipped steps (its tensor was just replaced) lags behind
        const double t = (double)(groups[i].step > 0 ? groups[i].step : step);
        const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
        a.neg_step_size[i] = (float)(-((double)groups[i].lr / bc1));
        a.bc2_sqrt[i] = (float)sqrt(bc2);
        a.first_chunk[i] = (unsigned int)chunks;
        chunks += (groups[i].count + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    a.first_chunk[num_groups] = (unsigned int)chunks;
    if (chunks == 0) return 0;
    if (chunks > 0x7fffffffull) {
        set_error("gsr_adam_step: too many parameters");
        return GSR_ERR_UNSUPPORTED;
    }
    const unsigned int nb = (unsigned int)chunks;
    cudaStream_t s = (cudaStream_t)stream_;
    adam_kernel<ADAM_TRIPS, false><<<nb, ADAM_THREADS, 0, s>>>(a); count_launches(1);
    return after_launch((cudaStream_t)stream_, false, "adam");
}

// ---------------------------------------------------------------------------------------------------------------------
// Row selection over a flat multi-tensor buffer: densify / prune as ONE index list.
// The reference rebuilds every parameter and both Adam moments with boolean masks and torch.cat, tensor by tensor
// (scene/gaussian_model.py:377-441: _prune_optimizer, cat_tensors_to_optimizer -- 2 cats + 2 masks x 7 tensors x 3 states per
// densification). Here the new row j of every block is row index[j] of the old buffer (-1: a fresh row of zeros, which is what
// the reference gives the moments of cloned / split Gaussians), for all blocks of the flat buffer in one launch.
// ---------------------------------------------------------------------------------------------------------------------
namespace gsr
{
namespace
{
constexpr int SEL_THREADS = 256;
constexpr int SEL_CHUNK = SEL_THREADS * 16; // output floats per CTA
constexpr int SEL_MAX_BLOCKS = 16;

struct SelectArgs
{
    const float* src;
    float* dst;
    const long long* index;
    long long n_out, n_src;
    int num_blocks;
    int row[SEL_MAX_BLOCKS];                          // floats per row of block k
    unsigned long long src_off[SEL_MAX_BLOCKS];       // start of block k in src / dst (floats)
    unsigned long long dst_off[SEL_MAX_BLOCKS];
    unsigned int first_chunk[SEL_MAX_BLOCKS + 1];
};

__global__ void __launch_bounds__(SEL_THREADS) select_rows_kernel(const SelectArgs a)
{
    int k = 0;
#pragma unroll 1
    while (k + 1 < a.num_blocks && blockIdx.x >= a.first_chunk[k + 1]) k++;
    const unsigned rf = (unsigned)a.row[k]; // > 0: zero-width blocks (features_rest at SH degree 0) own no CTA
    const unsigned long long n = (unsigned long long)a.n_out * rf;
    const unsigned long long c0 = (unsigned long long)(blockIdx.x - a.first_chunk[k]) * SEL_CHUNK;
    const float* src = a.src + a.src_off[k];
    float* dst = a.dst + a.dst_off[k];
#pragma unroll 4
    for (int t = 0; t < SEL_CHUNK / SEL_THREADS; t++) {
        const unsigned long long e = c0 + (unsigned long long)t * SEL_THREADS + threadIdx.x;
        if (e >= n) break;
        const unsigned long long r = e / rf;
        const unsigned c = (unsigned)(e - r * rf);
        const long long s = a.index[r];
        dst[e] = (s >= 0 && s < a.n_src) ? src[(unsigned long long)s * rf + c] : 0.f;
    }
}
} // namespace
} // namespace gsr

extern "C" int gsr_select_rows(const float* src, float* dst, const int64_t* index, int64_t n_out, int64_t n_src, const int32_t* row_floats,
                               const uint64_t* src_offsets, const uint64_t* dst_offsets, int32_t num_blocks, gsr_stream_t stream_)
{
    if (n_out <= 0 || num_blocks <= 0) return 0;
    if (!src || !dst || !index || !row_floats || !src_offsets || !dst_offsets || num_blocks > SEL_MAX_BLOCKS || n_src < 0) {
        set_error("gsr_select_rows: invalid argument (1 <= num_blocks <= %d)", SEL_MAX_BLOCKS);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    SelectArgs a;
    a.src = src; a.dst = dst; a.index = (const long long*)index; a.n_out = n_out; a.n_src = n_src; a.num_blocks = num_blocks;
    unsigned long long chunks = 0;
    for (int k = 0; k < num_blocks; k++) {
        if (row_floats[k] < 0) {
            set_error("gsr_select_rows: row_floats[%d] must not be negative", k);
            return GSR_ERR_INVALID_ARGUMENT;
        }
        a.row[k] = row_floats[k]; a.src_off[k] = src_offsets[k]; a.dst_off[k] = dst_offsets[k]; a.first_chunk[k] = (unsigned int)chunks;
        chunks += ((unsigned long long)row_floats[k] * (unsigned long long)n_out + SEL_CHUNK - 1) / SEL_CHUNK;
    }
    a.first_chunk[num_blocks] = (unsigned int)chunks;
    if (chunks == 0) return 0;
    if (chunks > 0x7fffffffull) {
        set_error("gsr_select_rows: too many rows");
        return GSR_ERR_UNSUPPORTED;
    }
    select_rows_kernel<<<(unsigned int)chunks, SEL_THREADS, 0, (cudaStream_t)stream_>>>(a); count_launches(1);
    return after_launch((cudaStream_t)stream_, false, "select_rows");
}
