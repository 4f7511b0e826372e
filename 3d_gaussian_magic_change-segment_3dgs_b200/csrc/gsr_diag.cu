// Measurement support: the compositing kernels are bound by FP32 issue, MUFU, shuffles and L2 atomics,
// not by HBM, so their roofline needs (a) the machine's achievable rates for exactly those instruction classes, measured on
// the GPU the bench runs on, and (b) the algorithmic work of the frame, counted from the bit-exact forward state.
//
//   gsr_microbench   FFMA / packed FFMA2 / MUFU.EX2 / SHFL chains and a 12-lane red.global.add.f32 pattern, each timed with
//                    CUDA events on the caller's stream (~1-3 ms per kernel).
//   gsr_count_work   replays the reference's per-pixel loop (forward.cu:314-377 semantics) over the saved lists and counts
//                    E   = list entries evaluated front to back until the pixel is done (or its list ends),
//                    Cc  = entries that contributed (passed the power / alpha / transmittance tests),
//                    E_b = entries the backward re-traverses (positions below the pixel's last contributor).
//
// Nothing here is on the product path; bench.py and the tests call it.
#include "gsr_common.cuh"

namespace gsr
{
namespace
{
constexpr int MB_THREADS = 256;
constexpr int MB_CHAINS = 8;

// a and b pass through memory so that the multiplier and the addend are per-thread REGISTERS (the 3-register FFMA form the
// compositing kernels issue), not constant-bank operands.
__global__ void __launch_bounds__(MB_THREADS) mb_ffma_kernel(float* out, int iters, float a, float b)
{
    a += out[threadIdx.x];
    b += out[threadIdx.x + MB_THREADS];
    float x[MB_CHAINS];
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) x[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s += x[k];
    if (s == 123.456f) out[0] = s; // never true in practice; keeps the chains alive
}

__global__ void __launch_bounds__(MB_THREADS) mb_ffma2_kernel(float* out, int iters, float a, float b)
{
    float2 x[MB_CHAINS];
    const float2 a2 = {a + out[threadIdx.x], a + out[threadIdx.x + 1]}, b2 = {b + out[threadIdx.x + MB_THREADS], b + out[threadIdx.x + 2]};
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) x[k] = {(float)(threadIdx.x + k), (float)k};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) x[k] = __ffma2_rn(x[k], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s += x[k].x + x[k].y;
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_ex2_kernel(float* out, int iters)
{
    float x[MB_CHAINS];
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) x[k] = 0.25f + 0.01f * (float)((threadIdx.x + k) & 15);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[k]) : "f"(-x[k])); // stays in (0.5, 1)
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s += x[k];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_shfl_kernel(float* out, int iters)
{
    float x[MB_CHAINS];
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) x[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1 + (k & 15));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s += x[k];
    if (s == 123.456f) out[0] = s;
}

// The compositing backward's atomic pattern: one warp instruction, 12 active lanes, 12 consecutive floats of one 48-byte record;
// records visited pseudo-randomly inside a table that fits in L2 (1.2 M records = 58 MB at cfg3).
__global__ void __launch_bounds__(MB_THREADS) mb_red_kernel(float* table, uint32_t records, int iters)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t h = (blockIdx.x * MB_THREADS + threadIdx.x) >> 5; // warp id
    for (int i = 0; i < iters; i++) {
        h = h * 1664525u + 1013904223u;
        const uint32_t rec = (h >> 8) % records;
        if (lane < GRAD_REC_FLOATS) atomicAdd(table + (size_t)rec * GRAD_REC_FLOATS + lane, 1.0f);
    }
}

// One CTA per tile, one thread per pixel; the tile's list is staged 256 entries at a time. Same tests, same order and the same
// arithmetic as the forward compositing loop, so the counts describe exactly the work the reference algorithm defines.
__global__ void __launch_bounds__(TILE_PIXELS) count_work_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                                                                 const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                                                                 const uint32_t* __restrict__ n_contrib, unsigned long long* out)
{
    __shared__ float4 sA[TILE_PIXELS];
    __shared__ float2 sB[TILE_PIXELS];
    __shared__ unsigned long long s_sum[3];
    if (threadIdx.x < 3) s_sum[threadIdx.x] = 0ull;
    __syncthreads();
    const uint32_t tx = blockIdx.x, ty = blockIdx.y;
    const uint32_t px = tx * TILE_X + (threadIdx.x & 15u), py = ty * TILE_Y + (threadIdx.x >> 4);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const float2 pixf = {(float)px, (float)py};
    const uint2 range = ranges[ty * (uint32_t)grid_x + tx];
    const uint32_t len = range.y - range.x;
    bool done = !inside;
    float T = 1.0f;
    uint32_t E = 0, Cc = 0;
    for (uint32_t b0 = 0; b0 < len; b0 += TILE_PIXELS) {
        if (__syncthreads_count(done) == TILE_PIXELS) break;
        const uint32_t k = b0 + threadIdx.x;
        if (k < len) {
            const float4* r = rec + 3 * (size_t)point_list[range.x + k];
            const float4 a = __ldg(r), b = __ldg(r + 1);
            sA[threadIdx.x] = a;
            sB[threadIdx.x] = {b.x, b.y};
        }
        __syncthreads();
        const uint32_t n = min((uint32_t)TILE_PIXELS, len - b0);
        for (uint32_t j = 0; !done && j < n; j++) {
            E++;
            const float4 xyc = sA[j];
            const float2 co = sB[j];
            const float2 d = {xyc.x - pixf.x, xyc.y - pixf.y};
            const float power = -0.5f * (xyc.z * d.x * d.x + co.x * d.y * d.y) - xyc.w * d.x * d.y;
            if (power > 0.0f) continue;
            const float alpha = min(0.99f, co.y * exp(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            Cc++;
            T = test_T;
        }
    }
    const uint32_t Eb = inside ? n_contrib[(size_t)W * py + px] : 0u;
    atomicAdd(&s_sum[0], (unsigned long long)E);
    atomicAdd(&s_sum[1], (unsigned long long)Cc);
    atomicAdd(&s_sum[2], (unsigned long long)Eb);
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(out + threadIdx.x, s_sum[threadIdx.x]);
}

template <typename F>
int time_launches(cudaStream_t s, int reps, float* ms_out, F launch)
{
    cudaEvent_t e0, e1;
    GSR_CUDA(cudaEventCreate(&e0));
    GSR_CUDA(cudaEventCreate(&e1));
    launch(); // warm-up
    GSR_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) launch();
    GSR_CUDA(cudaEventRecord(e1, s));
    GSR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GSR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms / reps;
    return cudaGetLastError() == cudaSuccess ? 0 : GSR_ERR_CUDA;
}
} // namespace
} // namespace gsr

using namespace gsr;

extern "C" int gsr_microbench(GsrMicrobench* r, gsr_stream_t stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (!r) {
        set_error("gsr_microbench: null result");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    memset(r, 0, sizeof(*r));
    int dev = 0, sms = 0, khz = 0;
    GSR_CUDA(cudaGetDevice(&dev));
    GSR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GSR_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    r->sm_count = sms;
    r->sm_clock_mhz_nominal = khz / 1000.0f;
    const int ctas = sms * 8; // 8 x 256 threads = 64 warps per SM: every scheduler has 16 warps to pick from
    const int iters = 4096;
    float* scratch = nullptr;
    const uint32_t records = 1200000u;
    GSR_CUDA(cudaMalloc(&scratch, (size_t)records * GRAD_REC_FLOATS * sizeof(float)));
    GSR_CUDA(cudaMemsetAsync(scratch, 0, (size_t)records * GRAD_REC_FLOATS * sizeof(float), s));
    const double lane_ops = (double)ctas * MB_THREADS * (double)iters * MB_CHAINS;
    float ms = 0.f;
    int rc = 0;
    rc = time_launches(s, 3, &ms, [&] { mb_ffma_kernel<<<ctas, MB_THREADS, 0, s>>>(scratch, iters, 0.999f, 0.001f); count_launches(1); });
    if (!rc) r->ffma_tflops = (float)(2.0 * lane_ops / (ms * 1e-3) / 1e12);
    if (!rc) rc = time_launches(s, 3, &ms, [&] { mb_ffma2_kernel<<<ctas, MB_THREADS, 0, s>>>(scratch, iters, 0.999f, 0.001f); count_launches(1); });
    if (!rc) r->ffma2_tflops = (float)(4.0 * lane_ops / (ms * 1e-3) / 1e12);
    if (!rc) rc = time_launches(s, 3, &ms, [&] { mb_ex2_kernel<<<ctas, MB_THREADS, 0, s>>>(scratch, iters); count_launches(1); });
    if (!rc) r->ex2_gops = (float)(lane_ops / (ms * 1e-3) / 1e9);
    if (!rc) rc = time_launches(s, 3, &ms, [&] { mb_shfl_kernel<<<ctas, MB_THREADS, 0, s>>>(scratch, iters); count_launches(1); });
    if (!rc) r->shfl_gops = (float)(lane_ops / (ms * 1e-3) / 1e9);
    const int red_iters = 512;
    if (!rc) rc = time_launches(s, 3, &ms, [&] { mb_red_kernel<<<ctas, MB_THREADS, 0, s>>>(scratch, records, red_iters); count_launches(1); });
    if (!rc) r->red_gops = (float)((double)ctas * (MB_THREADS / 32) * red_iters * GRAD_REC_FLOATS / (ms * 1e-3) / 1e9);
    cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (rc) set_error("gsr_microbench: a launch failed");
    return rc;
}

extern "C" int gsr_count_work(int32_t P, int32_t W, int32_t H, const GsrState* state, uint64_t* counters_dev, gsr_stream_t stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (P <= 0 || W <= 0 || H <= 0 || !state || !state->geom || !state->img || !counters_dev) {
        set_error("gsr_count_work: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    GeomState g;
    geom_layout((char*)state->geom, P, g);
    ImgState img;
    img_layout((char*)state->img, (size_t)W * H, (size_t)gx * gy, img);
    GSR_CUDA(cudaMemsetAsync(counters_dev, 0, 3 * sizeof(uint64_t), s));
    if (state->num_rendered <= 0) return 0;
    if (!state->binning) {
        set_error("gsr_count_work: binning state missing");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    count_work_kernel<<<dim3(gx, gy), TILE_PIXELS, 0, s>>>(W, H, gx, img.ranges, (const uint32_t*)state->binning, g.rec, img.n_contrib,
                                                          (unsigned long long*)counters_dev);
    count_launches(1);
    GSR_LAUNCHED(s, false, "count_work");
    return 0;
}
