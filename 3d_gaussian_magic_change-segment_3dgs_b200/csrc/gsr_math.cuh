// Per-Gaussian arithmetic of the splatting pipeline (EWA projection, covariance, SH colour and their
// backward rules), written as small scalar device functions.
//
// Numerical contract: results must equal what the reference's kernels compute
// (cuda_rasterizer/forward.cu, backward.cu, auxiliary.h) bit for bit where integer state is derived from them
// (radii, tile rectangles, depth keys), so every expression below keeps the reference's operand order and
// association -- including the general 3x3 products with structural zeros that GLM performs -- and nvcc's
// default FMA contraction then makes the same choices. No fast-math.
#pragma once
#include "gsr_common.cuh"

namespace gsr
{
// Real SH basis constants (auxiliary.h:22-39 of the reference; standard values).
__device__ constexpr float kSH0 = 0.28209479177387814f;
__device__ constexpr float kSH1 = 0.4886025119029199f;
__device__ constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                                      0.5462742152960396f};
__device__ constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                      -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// Column-major 3x3, c[col][row].
struct M3
{
    float c[3][3];
};

__device__ __forceinline__ M3 m3_cols(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8)
{
    M3 r;
    r.c[0][0] = a0; r.c[0][1] = a1; r.c[0][2] = a2;
    r.c[1][0] = a3; r.c[1][1] = a4; r.c[1][2] = a5;
    r.c[2][0] = a6; r.c[2][1] = a7; r.c[2][2] = a8;
    return r;
}
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b)
{
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.c[j][i] = a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1] + a.c[2][i] * b.c[j][2];
    return r;
}
__device__ __forceinline__ M3 m3_transpose(const M3& a)
{
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.c[j][i] = a.c[i][j];
    return r;
}

// Camera matrices are 16 floats indexed column-major (the reference passes the transposed row-major matrix).
__device__ __forceinline__ float3 xform_point(const float3& p, const float* m) // transformPoint4x3
{
    float3 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
    };
    return t;
}
__device__ __forceinline__ float4 xform_point_h(const float3& p, const float* m) // transformPoint4x4
{
    float4 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
        m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]};
    return t;
}

// Pixel coordinate from NDC, evaluated in double like the reference (auxiliary.h:41-44).
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// Tile rectangle of a splat (auxiliary.h:46-56): float division, truncation toward zero, clamp to the grid.
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, int grid_x, int grid_y, uint2& rmin, uint2& rmax)
{
    rmin = {(unsigned)min(grid_x, max((int)0, (int)((p.x - max_radius) / TILE_X))),
            (unsigned)min(grid_y, max((int)0, (int)((p.y - max_radius) / TILE_Y)))};
    rmax = {(unsigned)min(grid_x, max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X))),
            (unsigned)min(grid_y, max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)))};
}

__device__ __forceinline__ M3 quat_to_rot(float r, float x, float y, float z)
{
    return m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// Sigma = (S R)^T (S R), upper triangle (forward.cu:118-152). The quaternion is used as given (not normalised).
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot, float* cov3D)
{
    M3 S = m3_cols(1.0f, 0.f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f, 1.0f);
    S.c[0][0] = mod * scale.x;
    S.c[1][1] = mod * scale.y;
    S.c[2][2] = mod * scale.z;
    M3 R = quat_to_rot(rot.x, rot.y, rot.z, rot.w);
    M3 M = m3_mul(S, R);
    M3 Sigma = m3_mul(m3_transpose(M), M);
    cov3D[0] = Sigma.c[0][0];
    cov3D[1] = Sigma.c[0][1];
    cov3D[2] = Sigma.c[0][2];
    cov3D[3] = Sigma.c[1][1];
    cov3D[4] = Sigma.c[1][2];
    cov3D[5] = Sigma.c[2][2];
}

// EWA projection of the 3D covariance (forward.cu:74-113, also recomputed by backward.cu:164-199).
struct Cov2DCtx
{
    M3 T, Vrk, W;
    float3 t;          // view-space mean after the frustum clamp
    float txtz, tytz;  // un-clamped ratios
    float limx, limy;
};

__device__ __forceinline__ float3 cov2d_project(const float3& mean, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                                const float* cov3D, const float* view, Cov2DCtx* ctx)
{
    float3 t = xform_point(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;

    M3 J = m3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                   0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                   0, 0, 0);
    M3 W = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    M3 T = m3_mul(W, J);
    M3 Vrk = m3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
    cov.c[0][0] += 0.3f; // low-pass: at least one pixel wide
    cov.c[1][1] += 0.3f;
    if (ctx) {
        ctx->T = T; ctx->Vrk = Vrk; ctx->W = W; ctx->t = t;
        ctx->txtz = txtz; ctx->tytz = tytz; ctx->limx = limx; ctx->limy = limy;
    }
    return {float(cov.c[0][0]), float(cov.c[0][1]), float(cov.c[1][1])};
}

// SH coefficients of one Gaussian: coefficient 0 and coefficients 1..M-1 may live in two tensors (the model's _features_dc
// and _features_rest, scene/gaussian_model.py:111-114) -- the torch.cat in front of the reference rasterizer is never made.
// For the classic [P,M,3] tensor both pointers address the same row. k is a compile-time constant at every use.
template <class T>
struct ShCoeffs
{
    const T* dc;
    const T* rest; // coefficient k >= 1 is rest[k - 1]
    __device__ __forceinline__ const T& operator[](int k) const { return k == 0 ? dc[0] : rest[k - 1]; }
};

// Activations of the fused-parameter entry, spelled like ATen's CUDA kernels so the activated values carry the
// same bits as torch.sigmoid / torch.exp / F.normalize: sigmoid = 1 / (1 + exp(-x)); normalize = x / max(sqrt(sum x^2), 1e-12).
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dact_sigmoid(float y, float g) { return g * (1.0f - y) * y; } // sigmoid_backward: g * (1 - y) * y
// The sum of squares is associated the way ATen's reduction kernel does it for a contiguous [P,4] tensor (4 threads, one
// element each, two shuffle steps): (x^2 + z^2) + (y^2 + w^2), products rounded separately. Measured on torch 2.11: 0 of 10^6
// random quaternions differ from torch.linalg.vector_norm with this order (15 % differ with any other), and an ulp here is enough
// to flip an alpha < 1/255 decision downstream.
__device__ __forceinline__ float quat_norm(const float4& q)
{
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.z, q.z)), __fadd_rn(__fmul_rn(q.y, q.y), __fmul_rn(q.w, q.w)));
    return fmaxf(sqrtf(n2), 1e-12f);
}
__device__ __forceinline__ float4 act_normalize(const float4& q)
{
    const float n = quat_norm(q);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
// y = x / n: dL/dx = (g - y (y . g)) / n
__device__ __forceinline__ float4 dact_normalize(const float4& x, const float4& g)
{
    const float n = quat_norm(x);
    const float4 y = {x.x / n, x.y / n, x.z / n, x.w / n};
    const float d = y.x * g.x + y.y * g.y + y.z * g.z + y.w * g.w;
    return {(g.x - y.x * d) / n, (g.y - y.y * d) / n, (g.z - y.z * d) / n, (g.w - y.w * d) / n};
}

// View-dependent colour from SH coefficients (forward.cu:20-71). sh addresses this Gaussian's M coefficients.
// Returns max(result, 0) and the per-channel clamp bits.
template <class SH>
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float3& pos, const float3& campos, const SH sh, unsigned& clamp_bits)
{
    float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir = {dir.x / len, dir.y / len, dir.z / len};

    float res[3];
#define GSR_SH(k, c) ((c) == 0 ? sh[k].x : ((c) == 1 ? sh[k].y : sh[k].z))
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float result = kSH0 * GSR_SH(0, c);
        if (deg > 0) {
            const float x = dir.x, y = dir.y, z = dir.z;
            result = result - kSH1 * y * GSR_SH(1, c) + kSH1 * z * GSR_SH(2, c) - kSH1 * x * GSR_SH(3, c);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                result = result + kSH2[0] * xy * GSR_SH(4, c) + kSH2[1] * yz * GSR_SH(5, c) +
                         kSH2[2] * (2.0f * zz - xx - yy) * GSR_SH(6, c) + kSH2[3] * xz * GSR_SH(7, c) +
                         kSH2[4] * (xx - yy) * GSR_SH(8, c);
                if (deg > 2) {
                    result = result + kSH3[0] * y * (3.0f * xx - yy) * GSR_SH(9, c) + kSH3[1] * xy * z * GSR_SH(10, c) +
                             kSH3[2] * y * (4.0f * zz - xx - yy) * GSR_SH(11, c) +
                             kSH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * GSR_SH(12, c) +
                             kSH3[4] * x * (4.0f * zz - xx - yy) * GSR_SH(13, c) + kSH3[5] * z * (xx - yy) * GSR_SH(14, c) +
                             kSH3[6] * x * (xx - 3.0f * yy) * GSR_SH(15, c);
                }
            }
        }
        result += 0.5f;
        res[c] = result;
    }
#undef GSR_SH
    clamp_bits = (res[0] < 0 ? 1u : 0u) | (res[1] < 0 ? 2u : 0u) | (res[2] < 0 ? 4u : 0u);
    return {max(res[0], 0.0f), max(res[1], 0.0f), max(res[2], 0.0f)};
}

// d(normalize(v))/dv applied to dv (auxiliary.h:107-118).
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv)
{
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

struct V3
{
    float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, const V3& v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator*(const V3& v, float s) { return {v.x * s, v.y * s, v.z * s}; }
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3& operator+=(V3& a, const V3& b)
{
    a.x += b.x; a.y += b.y; a.z += b.z;
    return a;
}
__device__ __forceinline__ float dot3(const V3& a, const V3& b)
{
    V3 t = {a.x * b.x, a.y * b.y, a.z * b.z};
    return t.x + t.y + t.z;
}

// Backward of sh_to_rgb (backward.cu:20-139). sh: this Gaussian's coefficients; dL_dsh: its [M][3] output row in global
// memory (may be null), written for k < (deg+1)^2 only; the caller zero-fills the rest. Returns the mean-gradient
// contribution. Rows are stored as they are produced (like the reference) instead of being held in 48 registers.
struct ShRowWriter
{
    V3* row;
    struct Ref
    {
        V3* p;
        __device__ __forceinline__ void operator=(const V3& v) const
        {
            if (p) *p = v;
        }
    };
    __device__ __forceinline__ Ref operator[](int k) const { return Ref{row ? row + k : nullptr}; }
};

template <class SH>
__device__ __forceinline__ float3 sh_backward(int deg, const float3& pos, const float3& campos, const SH sh, unsigned clamp_bits,
                                              V3 dL_dRGB, ShRowWriter dL_dsh)
{
    V3 dir_orig = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};

    dL_dRGB.x *= (clamp_bits & 1u) ? 0 : 1;
    dL_dRGB.y *= (clamp_bits & 2u) ? 0 : 1;
    dL_dRGB.z *= (clamp_bits & 4u) ? 0 : 1;

    V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
    const float x = dir.x, y = dir.y, z = dir.z;

    float dRGBdsh0 = kSH0;
    dL_dsh[0] = dRGBdsh0 * dL_dRGB;
    if (deg > 0) {
        float dRGBdsh1 = -kSH1 * y;
        float dRGBdsh2 = kSH1 * z;
        float dRGBdsh3 = -kSH1 * x;
        dL_dsh[1] = dRGBdsh1 * dL_dRGB;
        dL_dsh[2] = dRGBdsh2 * dL_dRGB;
        dL_dsh[3] = dRGBdsh3 * dL_dRGB;

        dRGBdx = -kSH1 * sh[3];
        dRGBdy = -kSH1 * sh[1];
        dRGBdz = kSH1 * sh[2];

        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;

            float dRGBdsh4 = kSH2[0] * xy;
            float dRGBdsh5 = kSH2[1] * yz;
            float dRGBdsh6 = kSH2[2] * (2.f * zz - xx - yy);
            float dRGBdsh7 = kSH2[3] * xz;
            float dRGBdsh8 = kSH2[4] * (xx - yy);
            dL_dsh[4] = dRGBdsh4 * dL_dRGB;
            dL_dsh[5] = dRGBdsh5 * dL_dRGB;
            dL_dsh[6] = dRGBdsh6 * dL_dRGB;
            dL_dsh[7] = dRGBdsh7 * dL_dRGB;
            dL_dsh[8] = dRGBdsh8 * dL_dRGB;

            dRGBdx += kSH2[0] * y * sh[4] + kSH2[2] * 2.f * -x * sh[6] + kSH2[3] * z * sh[7] + kSH2[4] * 2.f * x * sh[8];
            dRGBdy += kSH2[0] * x * sh[4] + kSH2[1] * z * sh[5] + kSH2[2] * 2.f * -y * sh[6] + kSH2[4] * 2.f * -y * sh[8];
            dRGBdz += kSH2[1] * y * sh[5] + kSH2[2] * 2.f * 2.f * z * sh[6] + kSH2[3] * x * sh[7];

            if (deg > 2) {
                float dRGBdsh9 = kSH3[0] * y * (3.f * xx - yy);
                float dRGBdsh10 = kSH3[1] * xy * z;
                float dRGBdsh11 = kSH3[2] * y * (4.f * zz - xx - yy);
                float dRGBdsh12 = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                float dRGBdsh13 = kSH3[4] * x * (4.f * zz - xx - yy);
                float dRGBdsh14 = kSH3[5] * z * (xx - yy);
                float dRGBdsh15 = kSH3[6] * x * (xx - 3.f * yy);
                dL_dsh[9] = dRGBdsh9 * dL_dRGB;
                dL_dsh[10] = dRGBdsh10 * dL_dRGB;
                dL_dsh[11] = dRGBdsh11 * dL_dRGB;
                dL_dsh[12] = dRGBdsh12 * dL_dRGB;
                dL_dsh[13] = dRGBdsh13 * dL_dRGB;
                dL_dsh[14] = dRGBdsh14 * dL_dRGB;
                dL_dsh[15] = dRGBdsh15 * dL_dRGB;

                dRGBdx += (kSH3[0] * sh[9] * 3.f * 2.f * xy + kSH3[1] * sh[10] * yz + kSH3[2] * sh[11] * -2.f * xy +
                           kSH3[3] * sh[12] * -3.f * 2.f * xz + kSH3[4] * sh[13] * (-3.f * xx + 4.f * zz - yy) +
                           kSH3[5] * sh[14] * 2.f * xz + kSH3[6] * sh[15] * 3.f * (xx - yy));

                dRGBdy += (kSH3[0] * sh[9] * 3.f * (xx - yy) + kSH3[1] * sh[10] * xz + kSH3[2] * sh[11] * (-3.f * yy + 4.f * zz - xx) +
                           kSH3[3] * sh[12] * -3.f * 2.f * yz + kSH3[4] * sh[13] * -2.f * xy + kSH3[5] * sh[14] * -2.f * yz +
                           kSH3[6] * sh[15] * -3.f * 2.f * xy);

                dRGBdz += (kSH3[1] * sh[10] * xy + kSH3[2] * sh[11] * 4.f * 2.f * yz + kSH3[3] * sh[12] * 3.f * (2.f * zz - xx - yy) +
                           kSH3[4] * sh[13] * 4.f * 2.f * xz + kSH3[5] * sh[14] * (xx - yy));
            }
        }
    }
    float3 dL_ddir = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
    return dnormvdv(float3{dir_orig.x, dir_orig.y, dir_orig.z}, dL_ddir);
}

// The scalars dRGB/dsh_k that multiply dL_dRGB in sh_backward, for a unit direction (x,y,z); 0 above the active degree.
// Same expressions as sh_backward, so (sh_basis[k] * dL_dRGB) reproduces its dL_dsh rows bit for bit. Used by the
// receiving side of the multi-GPU gradient exchange, which rebuilds the 48-float SH row from the 3-float colour gradient.
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* w)
{
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = 0.f;
    w[0] = kSH0;
    if (deg > 0) {
        w[1] = -kSH1 * y;
        w[2] = kSH1 * z;
        w[3] = -kSH1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            w[4] = kSH2[0] * xy;
            w[5] = kSH2[1] * yz;
            w[6] = kSH2[2] * (2.f * zz - xx - yy);
            w[7] = kSH2[3] * xz;
            w[8] = kSH2[4] * (xx - yy);
            if (deg > 2) {
                w[9] = kSH3[0] * y * (3.f * xx - yy);
                w[10] = kSH3[1] * xy * z;
                w[11] = kSH3[2] * y * (4.f * zz - xx - yy);
                w[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                w[13] = kSH3[4] * x * (4.f * zz - xx - yy);
                w[14] = kSH3[5] * z * (xx - yy);
                w[15] = kSH3[6] * x * (xx - 3.f * yy);
            }
        }
    }
}

// Backward of cov3d_from_scale_rot (backward.cu:278-341): gradients w.r.t. the scale and the raw quaternion.
__device__ __forceinline__ void cov3d_backward(const float3 scale, float mod, const float4 rot, const float* dL_dcov3D,
                                               float3& dL_dscale, float4& dL_drot)
{
    const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    M3 R = quat_to_rot(r, x, y, z);
    M3 S = m3_cols(1.0f, 0.f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f, 1.0f);
    const float3 s = {mod * scale.x, mod * scale.y, mod * scale.z};
    S.c[0][0] = s.x;
    S.c[1][1] = s.y;
    S.c[2][2] = s.z;
    M3 M = m3_mul(S, R);

    M3 dL_dSigma = m3_cols(dL_dcov3D[0], 0.5f * dL_dcov3D[1], 0.5f * dL_dcov3D[2],
                           0.5f * dL_dcov3D[1], dL_dcov3D[3], 0.5f * dL_dcov3D[4],
                           0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[4], dL_dcov3D[5]);
    M3 M2;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            M2.c[j][i] = M.c[j][i] * 2.0f;
    M3 dL_dM = m3_mul(M2, dL_dSigma);
    M3 Rt = m3_transpose(R);
    M3 dL_dMt = m3_transpose(dL_dM);

    dL_dscale.x = dot3(V3{Rt.c[0][0], Rt.c[0][1], Rt.c[0][2]}, V3{dL_dMt.c[0][0], dL_dMt.c[0][1], dL_dMt.c[0][2]});
    dL_dscale.y = dot3(V3{Rt.c[1][0], Rt.c[1][1], Rt.c[1][2]}, V3{dL_dMt.c[1][0], dL_dMt.c[1][1], dL_dMt.c[1][2]});
    dL_dscale.z = dot3(V3{Rt.c[2][0], Rt.c[2][1], Rt.c[2][2]}, V3{dL_dMt.c[2][0], dL_dMt.c[2][1], dL_dMt.c[2][2]});

#pragma unroll
    for (int i = 0; i < 3; i++) {
        dL_dMt.c[0][i] *= s.x;
        dL_dMt.c[1][i] *= s.y;
        dL_dMt.c[2][i] *= s.z;
    }
#define GSR_D(a, b) dL_dMt.c[a][b]
    dL_drot.x = 2 * z * (GSR_D(0, 1) - GSR_D(1, 0)) + 2 * y * (GSR_D(2, 0) - GSR_D(0, 2)) + 2 * x * (GSR_D(1, 2) - GSR_D(2, 1));
    dL_drot.y = 2 * y * (GSR_D(1, 0) + GSR_D(0, 1)) + 2 * z * (GSR_D(2, 0) + GSR_D(0, 2)) + 2 * r * (GSR_D(1, 2) - GSR_D(2, 1)) -
                4 * x * (GSR_D(2, 2) + GSR_D(1, 1));
    dL_drot.z = 2 * x * (GSR_D(1, 0) + GSR_D(0, 1)) + 2 * r * (GSR_D(2, 0) - GSR_D(0, 2)) + 2 * z * (GSR_D(1, 2) + GSR_D(2, 1)) -
                4 * y * (GSR_D(2, 2) + GSR_D(0, 0));
    dL_drot.w = 2 * r * (GSR_D(0, 1) - GSR_D(1, 0)) + 2 * x * (GSR_D(2, 0) + GSR_D(0, 2)) + 2 * y * (GSR_D(1, 2) + GSR_D(2, 1)) -
                4 * z * (GSR_D(1, 1) + GSR_D(0, 0));
#undef GSR_D
}

} // namespace gsr
