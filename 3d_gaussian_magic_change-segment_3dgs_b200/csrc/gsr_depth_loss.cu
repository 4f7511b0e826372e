// Depth-supervision loss of the training step, value and gradient, without a sort.
//
// Reference (train.py:118-121, depth_loss_choice 'localrf'; utils/loss_utils.py:88-102 compute_depth_loss; the normalisation of
// gaussian_renderer/__init__.py:375):
//     y  = depth / (depth.max() + 1e-5)                       render()
//     x  = 1 / y.clamp(1e-6)                                   train.py:120
//     t  = median(x)   s  = mean|x - t|    xn = (x - t) / s    (torch.median: the LOWER middle element)
//     tg = median(g)   sg = mean|g - tg|   gn = (g - tg) / sg
//     a  = (xn - gn)^2 ;  a[a > quantile(a, 0.8)] = 0 ;  loss = lambda * mean(a)
// torch implements median and quantile with full sorts of the 2 M pixels (two kthvalue / sort launches of ~10 ms each at 1080p).
// Here every order statistic is a 4-pass radix SELECT over the monotone bit pattern of the floats (a 256-bin histogram per
// pass, restricted to the keys that share the prefix found so far), all sums are two-stage reductions with a fixed order
// (deterministic, double accumulators), and the gradient -- including the paths through the median element, the mean absolute
// deviation and, in fused mode, the arg-max pixel of the normalisation -- is one more pass. ~20 small launches, no host sync.
//
//   mode 0: `in` is compute_depth_loss's dyn_depth argument (x); the gradient is dL/dx.
//   mode 1: `in` is the rasterizer's raw depth image; y and x are formed here and the gradient is dL/d(depth).
#include <math.h>

#include "gsr_common.cuh"

namespace gsr
{
namespace
{
constexpr int DL_THREADS = 256;
constexpr int DL_ITEMS = 8;                       // elements per thread per CTA trip
constexpr int DL_CHUNK = DL_THREADS * DL_ITEMS;   // elements per CTA
constexpr int DL_MAX_SEL = 2;

// per-selection state: prefix / remaining rank after each pass, one histogram per pass
struct SelState
{
    uint32_t prefix[5];   // prefix[p]: the top 8 p bits (right aligned) of the key being selected, known before pass p
    uint32_t krem[5];     // rank still to be skipped inside that prefix
    uint32_t hist[4][256];
    float value;          // result (written by select_finish)
    uint32_t index;       // smallest element index holding `value` (filled by find_index_kernel when requested)
};

struct DlScalars // device-resident scalars of one evaluation
{
    float dmax;
    uint32_t nmax;        // pixels equal to the maximum
    float t, s, tg, sg;   // medians and mean absolute deviations
    float thr;            // quantile threshold
    double S_loss, S_u, S_ud, S_sgn, S_gyd;
};

__device__ __forceinline__ uint32_t float_key(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u); // monotone: key order == float order (negative zero below positive zero)
}
__device__ __forceinline__ float key_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Digit chosen by pass p - 1 from its histogram: every CTA recomputes it (256 values) instead of waiting for a broadcast.
__device__ __forceinline__ void resolve_pass(const SelState& st, int p, uint32_t& prefix, uint32_t& krem, uint32_t* s_tmp /*[256 + 8]*/)
{
    if (p == 0) {
        prefix = 0u;
        krem = st.krem[0];
        return;
    }
    const uint32_t k = st.krem[p - 1];
    const uint32_t c = st.hist[p - 1][threadIdx.x];
    // inclusive scan over 256 threads
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) s_tmp[256 + warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++)
        if (w < warp) base += s_tmp[256 + w];
    incl += base;
    const uint32_t excl = incl - c;
    if (k >= excl && k < incl) { // exactly one thread: the digit that contains rank k
        s_tmp[0] = threadIdx.x;
        s_tmp[1] = k - excl;
    }
    __syncthreads();
    prefix = (st.prefix[p - 1] << 8) | s_tmp[0];
    krem = s_tmp[1];
    __syncthreads();
}

struct SelSrc
{
    const float* v[DL_MAX_SEL]; // array each selection runs over (may be the same array)
    SelState* st[DL_MAX_SEL];
    int nsel;
    uint32_t n;
};

template <int PASS>
__global__ void __launch_bounds__(DL_THREADS) select_hist_kernel(const SelSrc a)
{
    __shared__ uint32_t s_hist[DL_MAX_SEL][256];
    __shared__ uint32_t s_tmp[256 + 8];
    uint32_t prefix[DL_MAX_SEL], krem[DL_MAX_SEL];
    for (int q = 0; q < a.nsel; q++) {
        resolve_pass(*a.st[q], PASS, prefix[q], krem[q], s_tmp);
        s_hist[q][threadIdx.x] = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0 && PASS > 0) { // publish for the next pass
            a.st[q]->prefix[PASS] = prefix[q];
            a.st[q]->krem[PASS] = krem[q];
        }
    }
    __syncthreads();
    constexpr int SHIFT = 24 - 8 * PASS;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    for (int q = 0; q < a.nsel; q++) {
        const float* v = a.v[q];
#pragma unroll
        for (int k = 0; k < DL_ITEMS; k++) {
            const uint32_t i = base + k * DL_THREADS + threadIdx.x;
            if (i < a.n) {
                const uint32_t key = float_key(v[i]);
                if (PASS == 0 || (key >> ((SHIFT + 8) & 31)) == prefix[q]) atomicAdd(&s_hist[q][(key >> SHIFT) & 0xffu], 1u);
            }
        }
    }
    __syncthreads();
    for (int q = 0; q < a.nsel; q++) {
        const uint32_t c = s_hist[q][threadIdx.x];
        if (c) atomicAdd(&a.st[q]->hist[PASS][threadIdx.x], c);
    }
}

__global__ void __launch_bounds__(DL_THREADS) select_finish_kernel(const SelSrc a)
{
    __shared__ uint32_t s_tmp[256 + 8];
    for (int q = 0; q < a.nsel; q++) {
        uint32_t prefix, krem;
        resolve_pass(*a.st[q], 4, prefix, krem, s_tmp);
        if (threadIdx.x == 0) {
            a.st[q]->prefix[4] = prefix;
            a.st[q]->value = key_float(prefix);
            a.st[q]->index = 0xffffffffu;
        }
    }
}

// smallest index i with v[i] == value (the element the median's gradient goes to)
__global__ void __launch_bounds__(DL_THREADS) find_index_kernel(const float* __restrict__ v, uint32_t n, SelState* st)
{
    const float value = st->value;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    uint32_t best = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n && v[i] == value) best = min(best, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31u) == 0 && best != 0xffffffffu) atomicMin(&st->index, best);
}

// ---- two-stage deterministic reductions: per-CTA partials, then one CTA in a fixed order ----
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* partial /*[gridDim.x][NV]*/)
{
    __shared__ double s_red[8][NV];
#pragma unroll
    for (int j = 0; j < NV; j++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if ((threadIdx.x & 31u) == 0)
#pragma unroll
        for (int j = 0; j < NV; j++) s_red[threadIdx.x >> 5][j] = v[j];
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * NV + threadIdx.x] = t;
    }
}

template <int NV>
__device__ __forceinline__ void final_reduce(const double* partial, uint32_t nb, double (&out)[NV], double* s_buf /*[DL_THREADS]*/)
{
#pragma unroll
    for (int j = 0; j < NV; j++) {
        double t = 0.0;
        for (uint32_t b = threadIdx.x; b < nb; b += DL_THREADS) t += partial[(size_t)b * NV + j];
        s_buf[threadIdx.x] = t;
        __syncthreads();
        for (int o = DL_THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < (unsigned)o) s_buf[threadIdx.x] += s_buf[threadIdx.x + o];
            __syncthreads();
        }
        out[j] = s_buf[0];
        __syncthreads();
    }
}

// mode 1: maximum of the depth image (partials), then x = 1 / max(d / (max + 1e-5), 1e-6) and the number of arg-max pixels
__global__ void __launch_bounds__(DL_THREADS) max_partial_kernel(const float* __restrict__ d, uint32_t n, float* partial)
{
    __shared__ float s_m[8];
    const uint32_t base = blockIdx.x * DL_CHUNK;
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) m = fmaxf(m, d[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31u) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; w++) m = fmaxf(m, s_m[w]);
        partial[blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(DL_THREADS) max_final_kernel(const float* partial, uint32_t nb, DlScalars* sc)
{
    __shared__ float s_m[DL_THREADS];
    float m = -INFINITY;
    for (uint32_t b = threadIdx.x; b < nb; b += DL_THREADS) m = fmaxf(m, partial[b]);
    s_m[threadIdx.x] = m;
    __syncthreads();
    for (int o = DL_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < (unsigned)o) s_m[threadIdx.x] = fmaxf(s_m[threadIdx.x], s_m[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sc->dmax = s_m[0];
        sc->nmax = 0;
    }
}
__global__ void __launch_bounds__(DL_THREADS) inverse_depth_kernel(const float* __restrict__ d, uint32_t n, DlScalars* sc, float* __restrict__ x)
{
    const float m = sc->dmax, den = m + 1e-5f;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            const float di = d[i];
            x[i] = 1.0f / fmaxf(di / den, 1e-6f);
            cnt += di == m ? 1u : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31u) == 0 && cnt) atomicAdd(&sc->nmax, cnt);
}

// mean absolute deviation around the medians (x and gt together)
__global__ void __launch_bounds__(DL_THREADS) mad_partial_kernel(const float* __restrict__ x, const float* __restrict__ g, uint32_t n,
                                                                 const SelState* mx, const SelState* mg, double* partial)
{
    const float t = mx->value, tg = mg->value;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            v[0] += (double)fabsf(x[i] - t);
            v[1] += (double)fabsf(g[i] - tg);
        }
    }
    block_reduce_store<2>(v, partial);
}
__global__ void __launch_bounds__(DL_THREADS) mad_final_kernel(const double* partial, uint32_t nb, uint32_t n, const SelState* mx,
                                                               const SelState* mg, DlScalars* sc)
{
    __shared__ double s_buf[DL_THREADS];
    double out[2];
    final_reduce<2>(partial, nb, out, s_buf);
    if (threadIdx.x == 0) {
        sc->t = mx->value;
        sc->tg = mg->value;
        sc->s = (float)(out[0] / (double)n);
        sc->sg = (float)(out[1] / (double)n);
    }
}

__global__ void __launch_bounds__(DL_THREADS) sqdiff_kernel(const float* __restrict__ x, const float* __restrict__ g, uint32_t n,
                                                            const DlScalars* sc, float* __restrict__ arr)
{
    const float t = sc->t, s = sc->s, tg = sc->tg, sg = sc->sg;
    const uint32_t base = blockIdx.x * DL_CHUNK;
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            const float dd = (x[i] - t) / s - (g[i] - tg) / sg;
            arr[i] = dd * dd;
        }
    }
}

// threshold = lerp(a_lo, a_hi, w) as torch.quantile does it, then the sums of the loss and of its gradient coefficients
__global__ void __launch_bounds__(DL_THREADS) sums_partial_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                  const float* __restrict__ arr, uint32_t n, const SelState* lo, const SelState* hi,
                                                                  float w, float lambda, DlScalars* sc, double* partial)
{
    const float a0 = lo->value, a1 = hi->value;
    const float thr = w < 0.5f ? a0 + w * (a1 - a0) : a1 - (a1 - a0) * (1.0f - w); // at::lerp
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->thr = thr;
    const float t = sc->t, s = sc->s, tg = sc->tg, sg = sc->sg;
    const float c = 2.0f * lambda / (float)n;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    double v[4] = {0.0, 0.0, 0.0, 0.0}; // loss numerator, sum u, sum u * xn, sum sign(x - t)
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            const float xi = x[i];
            const float xn = (xi - t) / s, dd = xn - (g[i] - tg) / sg;
            const float a = arr[i];
            if (!(a > thr)) { // kept (the reference zeroes a > threshold)
                const float u = c * dd;
                v[0] += (double)a;
                v[1] += (double)u;
                v[2] += (double)u * (double)xn;
            }
            v[3] += xi > t ? 1.0 : (xi < t ? -1.0 : 0.0);
        }
    }
    block_reduce_store<4>(v, partial);
}
__global__ void __launch_bounds__(DL_THREADS) sums_final_kernel(const double* partial, uint32_t nb, uint32_t n, float lambda, DlScalars* sc,
                                                                float* loss_out)
{
    __shared__ double s_buf[DL_THREADS];
    double out[4];
    final_reduce<4>(partial, nb, out, s_buf);
    if (threadIdx.x == 0) {
        sc->S_loss = out[0];
        sc->S_u = out[1];
        sc->S_ud = out[2];
        sc->S_sgn = out[3];
        loss_out[0] = (float)(out[0] / (double)n) * lambda;
    }
}

// dL/dx_i = u_i / s + Ds sign(x_i - t) / n + [i == median index] B          (mode 0: written out, scaled)
//   Ds = -(sum u xn) / s                 (through the mean absolute deviation)
//   B  = -(sum u) / s - Ds (sum sign) / n (through the median)
// mode 1 continues to y = d / (max + 1e-5):  gy_i = -x_i^2 gx_i where y_i >= 1e-6, and accumulates sum gy_i d_i for the arg-max path
template <int MODE>
__global__ void __launch_bounds__(DL_THREADS) grad_kernel(const float* __restrict__ in, const float* __restrict__ x, const float* __restrict__ g,
                                                          const float* __restrict__ arr, uint32_t n, const SelState* med, float lambda,
                                                          float grad_scale, const DlScalars* sc, float* __restrict__ out, double* partial)
{
    const float t = sc->t, s = sc->s, tg = sc->tg, sg = sc->sg, thr = sc->thr;
    const float c = 2.0f * lambda / (float)n;
    const double Ds = -sc->S_ud / (double)s;
    const double Bm = -sc->S_u / (double)s - Ds * sc->S_sgn / (double)n;
    const float Dsn = (float)(Ds / (double)n), Bf = (float)Bm;
    const uint32_t imed = med->index;
    const float den = MODE == 1 ? sc->dmax + 1e-5f : 1.f;
    const uint32_t base = blockIdx.x * DL_CHUNK;
    double v[1] = {0.0};
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            const float xi = x[i];
            const float xn = (xi - t) / s, dd = xn - (g[i] - tg) / sg;
            float gx = !(arr[i] > thr) ? c * dd / s : 0.f;
            gx += Dsn * (xi > t ? 1.f : (xi < t ? -1.f : 0.f));
            if (i == imed) gx += Bf;
            if (MODE == 0) {
                out[i] = grad_scale * gx;
            } else {
                const float di = in[i];
                const float y = di / den;
                const float gy = y >= 1e-6f ? -(xi * xi) * gx : 0.f;
                out[i] = gy; // finished by argmax_apply_kernel
                v[0] += (double)gy * (double)di;
            }
        }
    }
    if (MODE == 1) block_reduce_store<1>(v, partial);
}
__global__ void __launch_bounds__(DL_THREADS) gyd_final_kernel(const double* partial, uint32_t nb, DlScalars* sc)
{
    __shared__ double s_buf[DL_THREADS];
    double out[1];
    final_reduce<1>(partial, nb, out, s_buf);
    if (threadIdx.x == 0) sc->S_gyd = out[0];
}
// dL/dd_i = gy_i / (max + 1e-5)  -  [d_i == max] (sum_j gy_j d_j) / (max + 1e-5)^2 / (number of arg-max pixels)
__global__ void __launch_bounds__(DL_THREADS) argmax_apply_kernel(const float* __restrict__ d, uint32_t n, float grad_scale, const DlScalars* sc,
                                                                  float* __restrict__ out)
{
    const float m = sc->dmax, den = m + 1e-5f;
    const float share = (float)(-sc->S_gyd / ((double)den * (double)den) / (double)max(sc->nmax, 1u));
    const uint32_t base = blockIdx.x * DL_CHUNK;
#pragma unroll
    for (int k = 0; k < DL_ITEMS; k++) {
        const uint32_t i = base + k * DL_THREADS + threadIdx.x;
        if (i < n) {
            float r = out[i] / den;
            if (d[i] == m) r += share;
            out[i] = grad_scale * r;
        }
    }
}

__global__ void init_ranks_kernel(SelState* sel, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    if (threadIdx.x == 0) {
        sel[0].krem[0] = k0;
        sel[1].krem[0] = k1;
        sel[2].krem[0] = k2;
        sel[3].krem[0] = k3;
    }
}

struct DlLayout
{
    SelState* sel; // [4]: median x, median gt, quantile lo, quantile hi
    DlScalars* sc;
    double* partial; // [nb][4]
    float* fpartial; // [nb]
    float* x;        // [n] (mode 1)
    float* arr;      // [n]
};
size_t dl_layout(char* base, size_t n, DlLayout& l)
{
    char* p = base;
    const size_t nb = (n + DL_CHUNK - 1) / DL_CHUNK;
    carve(p, l.sel, 4);
    carve(p, l.sc, 1);
    carve(p, l.partial, nb * 4);
    carve(p, l.fpartial, nb);
    carve(p, l.x, n);
    carve(p, l.arr, n);
    return (size_t)(p - base) + 256;
}

int run_select(const SelSrc& a, uint32_t nb, cudaStream_t s)
{
    select_hist_kernel<0><<<nb, DL_THREADS, 0, s>>>(a);
    select_hist_kernel<1><<<nb, DL_THREADS, 0, s>>>(a);
    select_hist_kernel<2><<<nb, DL_THREADS, 0, s>>>(a);
    select_hist_kernel<3><<<nb, DL_THREADS, 0, s>>>(a);
    select_finish_kernel<<<1, DL_THREADS, 0, s>>>(a);
    count_launches(5);
    return after_launch(s, false, "depth_loss select");
}
} // namespace
} // namespace gsr

using namespace gsr;

extern "C" size_t gsr_depth_loss_scratch_bytes(int64_t n)
{
    if (n <= 0) return 0;
    DlLayout l;
    return dl_layout(nullptr, (size_t)n, l) + 256;
}

extern "C" int gsr_depth_loss(const float* in, const float* gt, int64_t n_, float lambda, float grad_scale, int32_t mode, float* loss_out,
                              float* grad_out, void* scratch, size_t scratch_bytes, gsr_stream_t stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (n_ <= 0 || n_ > 0x7fffffffll || !in || !gt || !loss_out || !scratch || (mode != 0 && mode != 1)) {
        set_error("gsr_depth_loss: invalid argument (0 < n < 2^31, mode 0 or 1)");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (scratch_bytes < gsr_depth_loss_scratch_bytes(n_)) {
        set_error("gsr_depth_loss: scratch too small (%zu < %zu)", scratch_bytes, gsr_depth_loss_scratch_bytes(n_));
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const uint32_t n = (uint32_t)n_;
    const uint32_t nb = (n + DL_CHUNK - 1) / DL_CHUNK;
    DlLayout l;
    dl_layout((char*)align_up((size_t)scratch, 256), n, l);
    // ranks: torch.median takes the lower middle element; torch.quantile computes q * (n - 1) in the tensor's dtype (float32)
    const uint32_t kmed = (n - 1) / 2;
    const float rank = 0.8f * (float)(n - 1);
    const float rank_lo = floorf(rank), rank_hi = ceilf(rank);
    const float w = rank - rank_lo;
    uint32_t k_lo = (uint32_t)rank_lo, k_hi = (uint32_t)rank_hi;
    if (k_lo > n - 1) k_lo = n - 1;
    if (k_hi > n - 1) k_hi = n - 1;
    GSR_CUDA(cudaMemsetAsync(l.sel, 0, 4 * sizeof(SelState), s));
    init_ranks_kernel<<<1, 32, 0, s>>>(l.sel, kmed, kmed, k_lo, k_hi);
    count_launches(1);
    const float* x = in;
    if (mode == 1) {
        max_partial_kernel<<<nb, DL_THREADS, 0, s>>>(in, n, l.fpartial);
        max_final_kernel<<<1, DL_THREADS, 0, s>>>(l.fpartial, nb, l.sc);
        inverse_depth_kernel<<<nb, DL_THREADS, 0, s>>>(in, n, l.sc, l.x);
        count_launches(3);
        x = l.x;
    }
    SelSrc med;
    med.v[0] = x; med.v[1] = gt; med.st[0] = &l.sel[0]; med.st[1] = &l.sel[1]; med.nsel = 2; med.n = n;
    int rc = run_select(med, nb, s);
    if (rc) return rc;
    if (grad_out) {
        find_index_kernel<<<nb, DL_THREADS, 0, s>>>(x, n, &l.sel[0]);
        count_launches(1);
    }
    mad_partial_kernel<<<nb, DL_THREADS, 0, s>>>(x, gt, n, &l.sel[0], &l.sel[1], l.partial);
    mad_final_kernel<<<1, DL_THREADS, 0, s>>>(l.partial, nb, n, &l.sel[0], &l.sel[1], l.sc);
    sqdiff_kernel<<<nb, DL_THREADS, 0, s>>>(x, gt, n, l.sc, l.arr);
    count_launches(3);
    SelSrc qs;
    qs.v[0] = l.arr; qs.v[1] = l.arr; qs.st[0] = &l.sel[2]; qs.st[1] = &l.sel[3]; qs.nsel = 2; qs.n = n;
    rc = run_select(qs, nb, s);
    if (rc) return rc;
    sums_partial_kernel<<<nb, DL_THREADS, 0, s>>>(x, gt, l.arr, n, &l.sel[2], &l.sel[3], w, lambda, l.sc, l.partial);
    sums_final_kernel<<<1, DL_THREADS, 0, s>>>(l.partial, nb, n, lambda, l.sc, loss_out);
    count_launches(2);
    if (grad_out) {
        if (mode == 0) {
            grad_kernel<0><<<nb, DL_THREADS, 0, s>>>(in, x, gt, l.arr, n, &l.sel[0], lambda, grad_scale, l.sc, grad_out, l.partial);
            count_launches(1);
        } else {
            grad_kernel<1><<<nb, DL_THREADS, 0, s>>>(in, x, gt, l.arr, n, &l.sel[0], lambda, grad_scale, l.sc, grad_out, l.partial);
            gyd_final_kernel<<<1, DL_THREADS, 0, s>>>(l.partial, nb, l.sc);
            argmax_apply_kernel<<<nb, DL_THREADS, 0, s>>>(in, n, grad_scale, l.sc, grad_out);
            count_launches(3);
        }
    }
    return after_launch(s, false, "depth_loss");
}
