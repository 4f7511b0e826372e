// Fused photometric loss of the training step:
//     loss = (1 - lambda) * mean|x - y| + lambda * (1 - mean(SSIM(x, y)))        train.py:110-111, utils/loss_utils.py:104-150
// and its gradient with respect to the rendered image x, in two kernels instead of torch's five grouped 11x11 convolutions, a
// dozen element-wise kernels and their autograd backward (again five convolutions).
//   K1 ssim_fwd_kernel: per 32x32 tile (+5 pixel halo, zero padding like F.conv2d(padding=5)) the separable 11-tap Gaussian
//      sums of x, y, x^2, y^2, xy -> the SSIM map value m and three derivative maps
//          Dmu = dm/dmu1 - 2 mu1 dm/dsigma1^2 - mu2 dm/dsigma12,   Ds1 = dm/dsigma1^2,   Ds12 = dm/dsigma12
//      plus per-CTA partial sums of m and |x - y| (no float atomics: the final sum is taken in a fixed order -> deterministic).
//   K2 ssim_bwd_kernel: the same separable filter over the three maps gives
//          d(sum m)/dx = conv(Dmu) + 2 x conv(Ds1) + y conv(Ds12)
//      (the window is symmetric, so the adjoint of the zero-padded correlation is the same correlation), combined with the L1
//      sign term and the two means into dL/dx. Its first CTA also reduces the partial sums to {L1, SSIM, loss}.
#include <math.h>

#include "gsr_common.cuh"

namespace gsr
{
namespace
{
constexpr int LT = 32;            // output tile edge
constexpr int LR = 5;             // window radius (window_size 11, utils/loss_utils.py:121)
constexpr int LW = LT + 2 * LR;   // input tile edge
constexpr int LOSS_THREADS = 256;

struct LossArgs
{
    const float* x;      // rendered image [C][H][W]
    const float* y;      // ground truth   [C][H][W]
    int C, H, W, tiles_x, tiles_y;
    float w[2 * LR + 1]; // normalised 1-D Gaussian, sigma 1.5
    float* maps;         // [3][C][H][W]
    float* partial;      // [2][blocks]
    float* loss_out;     // {l1, ssim, loss}
    float* grad;         // [C][H][W] or null
    float lambda_dssim, grad_scale;
};

__device__ __forceinline__ float block_sum(float v, float* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < LOSS_THREADS / 32 ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t; // valid in warp 0
}

__global__ void __launch_bounds__(LOSS_THREADS) ssim_fwd_kernel(const LossArgs a)
{
    __shared__ float s_x[LW][LW + 1], s_y[LW][LW + 1];
    __shared__ float s_h[5][LW][LT + 1]; // horizontally filtered x, y, xx, yy, xy
    __shared__ float s_red[LOSS_THREADS / 32];
    const int c = blockIdx.z, ty0 = blockIdx.y * LT, tx0 = blockIdx.x * LT;
    const float* X = a.x + (size_t)c * a.H * a.W;
    const float* Y = a.y + (size_t)c * a.H * a.W;
    for (int i = threadIdx.x; i < LW * LW; i += LOSS_THREADS) {
        const int r = i / LW, q = i - r * LW;
        const int gy = ty0 + r - LR, gx = tx0 + q - LR;
        const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        s_x[r][q] = in ? X[(size_t)gy * a.W + gx] : 0.f;
        s_y[r][q] = in ? Y[(size_t)gy * a.W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LW * LT; i += LOSS_THREADS) {
        const int r = i / LT, q = i - r * LT;
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * LR + 1; k++) {
            const float xv = s_x[r][q + k], yv = s_y[r][q + k], wk = a.w[k];
            sx += wk * xv; sy += wk * yv; sxx += wk * xv * xv; syy += wk * yv * yv; sxy += wk * xv * yv;
        }
        s_h[0][r][q] = sx; s_h[1][r][q] = sy; s_h[2][r][q] = sxx; s_h[3][r][q] = syy; s_h[4][r][q] = sxy;
    }
    __syncthreads();
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float sum_m = 0.f, sum_l1 = 0.f;
    const size_t plane = (size_t)a.H * a.W, chan = (size_t)c * plane;
    for (int i = threadIdx.x; i < LT * LT; i += LOSS_THREADS) {
        const int r = i / LT, q = i - r * LT;
        const int gy = ty0 + r, gx = tx0 + q;
        if (gy >= a.H || gx >= a.W) continue;
        float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * LR + 1; k++) {
            const float wk = a.w[k];
            mu1 += wk * s_h[0][r + k][q]; mu2 += wk * s_h[1][r + k][q]; exx += wk * s_h[2][r + k][q];
            eyy += wk * s_h[3][r + k][q]; exy += wk * s_h[4][r + k][q];
        }
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
        const float A = 2.f * mu12 + C1, B = 2.f * s12 + C2, Cc = mu1_sq + mu2_sq + C1, Dd = s1 + s2 + C2;
        const float m = (A * B) / (Cc * Dd);
        const float dm_dmu1 = (B / Dd) * (2.f * mu2 * Cc - A * 2.f * mu1) / (Cc * Cc);
        const float dm_ds1 = -(A * B) / (Cc * Dd * Dd);
        const float dm_ds12 = 2.f * A / (Cc * Dd);
        const size_t o = chan + (size_t)gy * a.W + gx;
        const size_t n = (size_t)a.C * plane;
        a.maps[o] = dm_dmu1 - 2.f * mu1 * dm_ds1 - mu2 * dm_ds12;
        a.maps[n + o] = dm_ds1;
        a.maps[2 * n + o] = dm_ds12;
        sum_m += m;
        sum_l1 += fabsf(s_x[r + LR][q + LR] - s_y[r + LR][q + LR]);
    }
    const float tm = block_sum(sum_m, s_red);
    const float tl = block_sum(sum_l1, s_red);
    if (threadIdx.x == 0) {
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        a.partial[b] = tm;
        a.partial[nb + b] = tl;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) ssim_bwd_kernel(const LossArgs a)
{
    __shared__ float s_m[3][LW][LW + 1];
    __shared__ float s_h[3][LW][LT + 1];
    __shared__ float s_red[LOSS_THREADS / 32];
    const int c = blockIdx.z, ty0 = blockIdx.y * LT, tx0 = blockIdx.x * LT;
    const size_t plane = (size_t)a.H * a.W, chan = (size_t)c * plane, n = (size_t)a.C * plane;
    const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { // fixed-order final reduction -> deterministic loss
        float sm = 0.f, sl = 0.f;
        for (unsigned i = threadIdx.x; i < nb; i += LOSS_THREADS) {
            sm += a.partial[i];
            sl += a.partial[nb + i];
        }
        const float tm = block_sum(sm, s_red);
        const float tl = block_sum(sl, s_red);
        if (threadIdx.x == 0) {
            const float l1 = tl / (float)n, ssim = tm / (float)n;
            a.loss_out[0] = l1;
            a.loss_out[1] = ssim;
            a.loss_out[2] = (1.0f - a.lambda_dssim) * l1 + a.lambda_dssim * (1.0f - ssim);
        }
        __syncthreads();
    }
    if (!a.grad) return;
    for (int i = threadIdx.x; i < LW * LW; i += LOSS_THREADS) {
        const int r = i / LW, q = i - r * LW;
        const int gy = ty0 + r - LR, gx = tx0 + q - LR;
        const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        const size_t o = chan + (size_t)(in ? gy : 0) * a.W + (in ? gx : 0);
        s_m[0][r][q] = in ? a.maps[o] : 0.f;
        s_m[1][r][q] = in ? a.maps[n + o] : 0.f;
        s_m[2][r][q] = in ? a.maps[2 * n + o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LW * LT; i += LOSS_THREADS) {
        const int r = i / LT, q = i - r * LT;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * LR + 1; k++) {
            const float wk = a.w[k];
            h0 += wk * s_m[0][r][q + k]; h1 += wk * s_m[1][r][q + k]; h2 += wk * s_m[2][r][q + k];
        }
        s_h[0][r][q] = h0; s_h[1][r][q] = h1; s_h[2][r][q] = h2;
    }
    __syncthreads();
    const float inv_n = 1.0f / (float)n;
    const float k_l1 = (1.0f - a.lambda_dssim) * inv_n * a.grad_scale, k_ss = -a.lambda_dssim * inv_n * a.grad_scale;
    for (int i = threadIdx.x; i < LT * LT; i += LOSS_THREADS) {
        const int r = i / LT, q = i - r * LT;
        const int gy = ty0 + r, gx = tx0 + q;
        if (gy >= a.H || gx >= a.W) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * LR + 1; k++) {
            const float wk = a.w[k];
            v0 += wk * s_h[0][r + k][q]; v1 += wk * s_h[1][r + k][q]; v2 += wk * s_h[2][r + k][q];
        }
        const size_t o = chan + (size_t)gy * a.W + gx;
        const float xv = a.x[o], yv = a.y[o], d = xv - yv;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); // torch.abs backward: sign(x - y)
        a.grad[o] = k_l1 * sgn + k_ss * (v0 + 2.f * xv * v1 + yv * v2);
    }
}
} // namespace
} // namespace gsr

using namespace gsr;

extern "C" size_t gsr_image_loss_scratch_bytes(int32_t C, int32_t H, int32_t W)
{
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    const size_t tiles = (size_t)((W + LT - 1) / LT) * ((H + LT - 1) / LT) * C;
    return (3 * (size_t)C * H * W + 2 * tiles) * sizeof(float);
}

extern "C" int gsr_image_loss(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim, float grad_scale,
                              float* loss_out, float* dL_dimage, void* scratch, size_t scratch_bytes, gsr_stream_t stream_)
{
    if (!image || !gt || !loss_out || !scratch || C <= 0 || H <= 0 || W <= 0 || C > 65535) {
        set_error("gsr_image_loss: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (scratch_bytes < gsr_image_loss_scratch_bytes(C, H, W)) {
        set_error("gsr_image_loss: scratch too small (%zu < %zu)", scratch_bytes, gsr_image_loss_scratch_bytes(C, H, W));
        return GSR_ERR_INVALID_ARGUMENT;
    }
    LossArgs a;
    a.x = image; a.y = gt; a.C = C; a.H = H; a.W = W;
    a.tiles_x = (W + LT - 1) / LT; a.tiles_y = (H + LT - 1) / LT;
    // gaussian(11, 1.5) of utils/loss_utils.py:110-112, in fp32 like torch.Tensor([...]) / sum
    float g[2 * LR + 1], sum = 0.f;
    for (int i = 0; i < 2 * LR + 1; i++) {
        g[i] = (float)exp(-(double)((i - LR) * (i - LR)) / (2.0 * 1.5 * 1.5));
        sum += g[i];
    }
    for (int i = 0; i < 2 * LR + 1; i++) a.w[i] = g[i] / sum;
    a.maps = (float*)scratch;
    a.partial = a.maps + 3 * (size_t)C * H * W;
    a.loss_out = loss_out; a.grad = dL_dimage; a.lambda_dssim = lambda_dssim; a.grad_scale = grad_scale;
    if (a.tiles_y > 65535) {
        set_error("gsr_image_loss: image too tall");
        return GSR_ERR_UNSUPPORTED;
    }
    const dim3 grid(a.tiles_x, a.tiles_y, C);
    cudaStream_t s = (cudaStream_t)stream_;
    ssim_fwd_kernel<<<grid, LOSS_THREADS, 0, s>>>(a);
    ssim_bwd_kernel<<<grid, LOSS_THREADS, 0, s>>>(a);
    count_launches(2);
    return after_launch(s, false, "image_loss");
}
