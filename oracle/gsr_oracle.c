/*
 * Oracle B: plain-C CPU restatement of the reference rasterizer and simple-knn.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path; it is loaded by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` fallback leg, as the
 * checker or the reported CPU baseline, never as the thing shipped. The product (libgsr.so) has no CPU path.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/submodules_local/):
 *   DGR = diff-gaussian-rasterization/cuda_rasterizer ,  KNN = simple-knn
 *
 * Arithmetic: fp32 throughout, compiled with -ffp-contract=off so results do not depend on the host
 * compiler's FMA choices. nvcc contracts a*b+c into FMAs in both the reference and the product, so this
 * oracle agrees with either CUDA build to fp32 rounding, not bit for bit; bit-exact integer parity
 * (radii, keys, order, ranges, n_contrib) is pinned against oracle A (oracle/_ref, the reference's own
 * CUDA compiled for sm_100a) on the GPU box and through tests/golden/ fixtures produced from it.
 * GLM (un-vendored, unpinned submodule of the reference) boundary: "parity unpinned" -- the mat3
 * products below follow GLM 0.9.9's column-major term order r[c][r] = a[0][r]*b[c][0] + a[1][r]*b[c][1]
 * + a[2][r]*b[c][2].
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16 /* DGR/config.h:17-18 BLOCK_X = BLOCK_Y = 16 */

/* ------------------------------------------------------------------ small column-major mat3 helpers */
typedef struct { float c[3][3]; } m3; /* c[col][row], as glm::mat3 */

static m3 m3_mul(const m3* a, const m3* b)
{
    m3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            r.c[j][i] = a->c[0][i] * b->c[j][0] + a->c[1][i] * b->c[j][1] + a->c[2][i] * b->c[j][2];
    return r;
}
static m3 m3_t(const m3* a)
{
    m3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            r.c[j][i] = a->c[i][j];
    return r;
}
static m3 m3_cols(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8)
{
    m3 r = {{{a0, a1, a2}, {a3, a4, a5}, {a6, a7, a8}}};
    return r;
}

/* DGR/auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* DGR/auxiliary.h:41-44  evaluated in double, rounded to float */
static float ndc2pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

/* float -> int with CUDA's saturating cvt.rzi semantics (plain C casts are UB out of range) */
static int f2i(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int)f;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* DGR/auxiliary.h:46-56 getRect */
static void get_rect(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
    *x0 = clampi(f2i((px - (float)max_radius) / (float)TILE), 0, gx);
    *y0 = clampi(f2i((py - (float)max_radius) / (float)TILE), 0, gy);
    *x1 = clampi(f2i((px + (float)max_radius + (float)TILE - 1.0f) / (float)TILE), 0, gx);
    *y1 = clampi(f2i((py + (float)max_radius + (float)TILE - 1.0f) / (float)TILE), 0, gy);
}

/* DGR/auxiliary.h:58-77 */
static void xform4x3(const float* p, const float* m, float* o)
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xform4x4(const float* p, const float* m, float* o)
{
    xform4x3(p, m, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* DGR/forward.cu:118-152 computeCov3D (quaternion NOT normalised, :127) */
static void cov3d_from_scale_rot(const float* scale, float mod, const float* rot, float* cov3D)
{
    m3 S = m3_cols(1, 0, 0, 0, 1, 0, 0, 0, 1);
    S.c[0][0] = mod * scale[0];
    S.c[1][1] = mod * scale[1];
    S.c[2][2] = mod * scale[2];
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    m3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    m3 M = m3_mul(&S, &R);
    m3 Mt = m3_t(&M);
    m3 Sig = m3_mul(&Mt, &M);
    cov3D[0] = Sig.c[0][0];
    cov3D[1] = Sig.c[0][1];
    cov3D[2] = Sig.c[0][2];
    cov3D[3] = Sig.c[1][1];
    cov3D[4] = Sig.c[1][2];
    cov3D[5] = Sig.c[2][2];
}

/* shared by DGR/forward.cu:74-113 (computeCov2D) and DGR/backward.cu:164-199 */
typedef struct { m3 T, Vrk, W; float tx, ty, tz, txtz, tytz, limx, limy; } cov2d_ctx;

static void cov2d(const float* mean, float fx, float fy, float tan_fovx, float tan_fovy, const float* cov3D,
                  const float* view, float* cov /*a,b,c*/, cov2d_ctx* ctx)
{
    float t[3];
    xform4x3(mean, view, t);
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    m3 J = m3_cols(fx / t[2], 0.0f, -(fx * t[0]) / (t[2] * t[2]), 0.0f, fy / t[2], -(fy * t[1]) / (t[2] * t[2]), 0, 0, 0);
    m3 W = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    m3 T = m3_mul(&W, &J);
    m3 Vrk = m3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    m3 Tt = m3_t(&T), Vt = m3_t(&Vrk);
    m3 tmp = m3_mul(&Tt, &Vt);
    m3 c = m3_mul(&tmp, &T);
    cov[0] = c.c[0][0] + 0.3f; /* low-pass filter, forward.cu:110-111 */
    cov[1] = c.c[0][1];
    cov[2] = c.c[1][1] + 0.3f;
    if (ctx) {
        ctx->T = T; ctx->Vrk = Vrk; ctx->W = W;
        ctx->tx = t[0]; ctx->ty = t[1]; ctx->tz = t[2];
        ctx->txtz = txtz; ctx->tytz = tytz; ctx->limx = limx; ctx->limy = limy;
    }
}

/* DGR/forward.cu:20-71 computeColorFromSH */
static void sh_to_rgb(int deg, int M, const float* mean, const float* campos, const float* sh /*[M][3]*/, float* rgb,
                      uint8_t* clamped)
{
    float dir[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
    float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
    (void)M;
    for (int c = 0; c < 3; c++) {
#define SHC(k) sh[(k)*3 + c]
        float result = SH_C0 * SHC(0);
        if (deg > 0) {
            result = result - SH_C1 * y * SHC(1) + SH_C1 * z * SHC(2) - SH_C1 * x * SHC(3);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                result = result + SH_C2[0] * xy * SHC(4) + SH_C2[1] * yz * SHC(5) + SH_C2[2] * (2.0f * zz - xx - yy) * SHC(6) +
                         SH_C2[3] * xz * SHC(7) + SH_C2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    result = result + SH_C3[0] * y * (3.0f * xx - yy) * SHC(9) + SH_C3[1] * xy * z * SHC(10) +
                             SH_C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
                             SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
                             SH_C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + SH_C3[5] * z * (xx - yy) * SHC(14) +
                             SH_C3[6] * x * (xx - 3.0f * yy) * SHC(15);
                }
            }
        }
#undef SHC
        result += 0.5f;
        clamped[c] = (result < 0);
        rgb[c] = fmaxf(result, 0.0f);
    }
}

/*
 * DGR/forward.cu:155-256 preprocessCUDA (+ auxiliary.h:139-164 in_frustum).
 * All per-Gaussian state arrays are indexed by Gaussian id like the reference's GeometryState.
 * Returns the number of visible Gaussians. `prefiltered` culls trap in the reference; here they return -1.
 */
int orc_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                   const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                   const float* colors_precomp, const float* view, const float* proj, const float* campos, int W, int H,
                   float tan_fovx, float tan_fovy, int prefiltered, int* radii, float* means2D, float* depths,
                   float* cov3Ds, float* rgb, float* conic_opacity, uint8_t* clamped, uint32_t* tiles_touched)
{
    /* DGR/rasterizer_impl.cu:226-227 */
    const float focal_y = H / (2.0f * tan_fovy);
    const float focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    int visible = 0, trapped = 0;
#pragma omp parallel for schedule(static) reduction(+ : visible) reduction(| : trapped)
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        const float* p = means3D + 3 * (size_t)idx;
        float ph[4], pv[3];
        xform4x4(p, proj, ph);
        float p_w = 1.0f / (ph[3] + 0.0000001f);
        float pp[3] = {ph[0] * p_w, ph[1] * p_w, ph[2] * p_w};
        xform4x3(p, view, pv);
        if (pv[2] <= 0.2f) { /* auxiliary.h:154 */
            if (prefiltered) trapped |= 1;
            continue;
        }
        const float* cov3D;
        if (cov3D_precomp) {
            cov3D = cov3D_precomp + 6 * (size_t)idx;
        } else {
            cov3d_from_scale_rot(scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx, cov3Ds + 6 * (size_t)idx);
            cov3D = cov3Ds + 6 * (size_t)idx;
        }
        float cov[3];
        cov2d(p, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, view, cov, NULL);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float px = ndc2pix(pp[0], W), py = ndc2pix(pp[1], H);
        int x0, y0, x1, y1;
        get_rect(px, py, f2i(my_radius), gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (!colors_precomp)
            sh_to_rgb(D, M, p, campos, shs + (size_t)idx * M * 3, rgb + 3 * (size_t)idx, clamped + 3 * (size_t)idx);
        depths[idx] = pv[2];
        radii[idx] = f2i(my_radius);
        means2D[2 * (size_t)idx] = px;
        means2D[2 * (size_t)idx + 1] = py;
        conic_opacity[4 * (size_t)idx + 0] = conic[0];
        conic_opacity[4 * (size_t)idx + 1] = conic[1];
        conic_opacity[4 * (size_t)idx + 2] = conic[2];
        conic_opacity[4 * (size_t)idx + 3] = opacities[idx];
        tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
        visible++;
    }
    return trapped ? -1 : visible;
}

/* DGR/rasterizer_impl.cu:281 cub::DeviceScan::InclusiveSum. Returns the total (= num_rendered, :285). */
uint32_t orc_inclusive_sum(int P, const uint32_t* in, uint32_t* out)
{
    uint32_t acc = 0;
    for (int i = 0; i < P; i++) {
        acc += in[i];
        out[i] = acc;
    }
    return acc;
}

/* DGR/rasterizer_impl.cu:35-50 getHigherMsb */
uint32_t orc_higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* DGR/rasterizer_impl.cu:70-111 duplicateWithKeys */
void orc_duplicate_with_keys(int P, const float* means2D, const float* depths, const uint32_t* offsets, const int* radii,
                             int W, int H, uint64_t* keys, uint32_t* values)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int idx = 0; idx < P; idx++) {
        if (radii[idx] <= 0) continue;
        uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
        int x0, y0, x1, y1;
        get_rect(means2D[2 * (size_t)idx], means2D[2 * (size_t)idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, depths + idx, 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key <<= 32;
                key |= dbits;
                keys[off] = key;
                values[off] = (uint32_t)idx;
                off++;
            }
    }
}

/*
 * DGR/rasterizer_impl.cu:307-312 cub::DeviceRadixSort::SortPairs(keys, values, n, 0, end_bit):
 * stable ascending LSD radix sort on key bits [0, end_bit). Restated as 16-bit-digit counting passes.
 */
void orc_sort_pairs(size_t n, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                    int end_bit)
{
    uint64_t* kb[2];
    uint32_t* vb[2];
    kb[0] = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    vb[0] = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    kb[1] = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    vb[1] = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    memcpy(kb[0], keys_in, n * sizeof(uint64_t));
    memcpy(vb[0], vals_in, n * sizeof(uint32_t));
    int cur = 0;
    size_t* count = (size_t*)malloc(65537 * sizeof(size_t));
    for (int shift = 0; shift < end_bit; shift += 16) {
        int bits = end_bit - shift < 16 ? end_bit - shift : 16;
        uint64_t mask = ((uint64_t)1 << bits) - 1;
        memset(count, 0, 65537 * sizeof(size_t));
        for (size_t i = 0; i < n; i++) count[((kb[cur][i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 65536; d++) count[d + 1] += count[d];
        for (size_t i = 0; i < n; i++) {
            size_t dst = count[(kb[cur][i] >> shift) & mask]++;
            kb[cur ^ 1][dst] = kb[cur][i];
            vb[cur ^ 1][dst] = vb[cur][i];
        }
        cur ^= 1;
    }
    memcpy(keys_out, kb[cur], n * sizeof(uint64_t));
    memcpy(vals_out, vb[cur], n * sizeof(uint32_t));
    free(count);
    free(kb[0]); free(kb[1]); free(vb[0]); free(vb[1]);
}

/* DGR/rasterizer_impl.cu:314 (memset) + :116-138 identifyTileRanges. ranges = uint2[T] as u32 pairs. */
void orc_identify_tile_ranges(size_t L, const uint64_t* keys, int T, uint32_t* ranges)
{
    memset(ranges, 0, (size_t)T * 2 * sizeof(uint32_t));
    for (size_t idx = 0; idx < L; idx++) {
        uint32_t currtile = (uint32_t)(keys[idx] >> 32);
        if (idx == 0) ranges[2 * (size_t)currtile] = 0;
        else {
            uint32_t prevtile = (uint32_t)(keys[idx - 1] >> 32);
            if (currtile != prevtile) {
                ranges[2 * (size_t)prevtile + 1] = (uint32_t)idx;
                ranges[2 * (size_t)currtile] = (uint32_t)idx;
            }
        }
        if (idx == L - 1) ranges[2 * (size_t)currtile + 1] = (uint32_t)L;
    }
}

/*
 * DGR/forward.cu:261-392 renderCUDA (forward). One tile per outer iteration, one pixel per inner one;
 * the block-wide early exit (:317-319) only stops work once every pixel is done, so per-pixel
 * sequential evaluation is equivalent. `features` is colors_precomp or the preprocess rgb (:325).
 * S = number of segment channels (reference NUM_CLASS = 2); segments may be NULL when S == 0.
 * row_stride > 1 restricts the work to tile rows with row % row_stride == row_offset (bench.py's bounded CPU sample).
 */
void orc_render_forward(int W, int H, int S, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                        const float* features, const float* segments, const float* depths, const float* conic_opacity,
                        const float* bg, float* out_color, float* out_segment, float* out_depth, float* out_alpha,
                        uint32_t* n_contrib, int row_stride, int row_offset)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        if (row_stride > 1 && ty % row_stride != row_offset) continue; /* bounded sample for the CPU baseline timing */
        const uint32_t r0 = ranges[2 * (size_t)tile], r1 = ranges[2 * (size_t)tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix_id = (size_t)W * py + px;
                const float pixfx = (float)px, pixfy = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0}, Sg[16] = {0}, weight = 0, Dd = 0;
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t k = r0; k < r1; k++) {
                    contributor++;
                    const uint32_t g = point_list[k];
                    const float dx = means2D[2 * (size_t)g] - pixfx, dy = means2D[2 * (size_t)g + 1] - pixfy;
                    const float* co = conic_opacity + 4 * (size_t)g;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break; /* done = true (:355-359) */
                    for (int ch = 0; ch < 3; ch++) C[ch] += features[3 * (size_t)g + ch] * alpha * T;
                    weight += alpha * T;
                    Dd += depths[g] * alpha * T;
                    for (int c = 0; c < S; c++) Sg[c] += segments[(size_t)g * S + c] * alpha * T;
                    T = test_T;
                    last_contributor = contributor;
                }
                n_contrib[pix_id] = last_contributor;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg[ch];
                out_alpha[pix_id] = weight; /* forward.cu:386: sum of alpha*T, not 1-T */
                out_depth[pix_id] = Dd;
                for (int c = 0; c < S; c++) out_segment[c * HW + pix_id] = Sg[c];
            }
    }
}

static void atomic_addd(double* p, double v)
{
#pragma omp atomic
    *p += v;
}

/*
 * DGR/backward.cu:415-639 renderCUDA (backward). Per-Gaussian sums are accumulated in double (the
 * reference uses fp32 atomicAdd in arbitrary order; the test tolerance of 1e-4 relative covers that).
 * Outputs (double, zero-initialised by the caller, indexed by Gaussian id):
 *   dmean2D[P][2], dconic[P][3] (.x,.y,.w of the reference's float4; .z is never written, :631-633),
 *   dopacity[P], dcolors[P][3], dsegments[P][S], ddepths[P].
 */
void orc_render_backward(int W, int H, int S, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                         const float* means2D, const float* conic_opacity, const float* colors, const float* segments,
                         const float* depths, const float* alphas, const uint32_t* n_contrib, const float* dL_dpixels,
                         const float* dL_dpixels_segments, const float* dL_dpixel_depths, const float* dL_dalphas,
                         double* dmean2D, double* dconic, double* dopacity, double* dcolors, double* dsegments,
                         double* ddepths, int row_stride, int row_offset)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H); /* backward.cu:504-505 */
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        if (row_stride > 1 && ty % row_stride != row_offset) continue;
        const uint32_t r0 = ranges[2 * (size_t)tile], r1 = ranges[2 * (size_t)tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix_id = (size_t)W * py + px;
                const float pixfx = (float)px, pixfy = (float)py;
                const float T_final = 1 - alphas[pix_id]; /* :468 */
                float T = T_final;
                uint32_t contributor = r1 - r0;
                const uint32_t last_contributor = n_contrib[pix_id];
                float accum_rec[3] = {0, 0, 0}, dL_dpixel[3], accum_seg[16] = {0}, dL_dseg[16] = {0};
                float accum_depth_rec = 0, accum_alpha_rec = 0;
                for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * HW + pix_id];
                const float dL_dpixel_depth = dL_dpixel_depths[pix_id];
                const float dL_dalpha = dL_dalphas[pix_id];
                for (int i = 0; i < S; i++) dL_dseg[i] = dL_dpixels_segments[i * HW + pix_id];
                float last_alpha = 0, last_color[3] = {0, 0, 0}, last_seg[16] = {0}, last_depth = 0;
                for (uint32_t k = r1; k-- > r0;) {
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t g = point_list[k];
                    const float dx = means2D[2 * (size_t)g] - pixfx, dy = means2D[2 * (size_t)g + 1] - pixfy;
                    const float* co = conic_opacity + 4 * (size_t)g;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dopa = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = colors[3 * (size_t)g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        const float dL_dchannel = dL_dpixel[ch];
                        dL_dopa += (c - accum_rec[ch]) * dL_dchannel;
                        atomic_addd(&dcolors[3 * (size_t)g + ch], (double)(dchannel_dcolor * dL_dchannel));
                    }
                    for (int ch = 0; ch < S; ch++) {
                        const float c_s = segments[(size_t)g * S + ch];
                        accum_seg[ch] = last_alpha * last_seg[ch] + (1.f - last_alpha) * accum_seg[ch];
                        last_seg[ch] = c_s;
                        const float dL_dclass = dL_dseg[ch];
                        dL_dopa += (c_s - accum_seg[ch]) * dL_dclass;
                        atomic_addd(&dsegments[(size_t)g * S + ch], (double)(dchannel_dcolor * dL_dclass));
                    }
                    const float c_d = depths[g];
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dopa += (c_d - accum_depth_rec) * dL_dpixel_depth;
                    atomic_addd(&ddepths[g], (double)(dchannel_dcolor * dL_dpixel_depth));
                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec; /* :604 */
                    dL_dopa += (1 - accum_alpha_rec) * dL_dalpha;
                    dL_dopa *= T;
                    last_alpha = alpha;
                    float bg_dot_dpixel = 0;
                    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg[i] * dL_dpixel[i];
                    dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = co[3] * dL_dopa;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    atomic_addd(&dmean2D[2 * (size_t)g], (double)(dL_dG * dG_ddelx * ddelx_dx));
                    atomic_addd(&dmean2D[2 * (size_t)g + 1], (double)(dL_dG * dG_ddely * ddely_dy));
                    atomic_addd(&dconic[3 * (size_t)g], (double)(-0.5f * gdx * dx * dL_dG));
                    atomic_addd(&dconic[3 * (size_t)g + 1], (double)(-0.5f * gdx * dy * dL_dG));
                    atomic_addd(&dconic[3 * (size_t)g + 2], (double)(-0.5f * gdy * dy * dL_dG));
                    atomic_addd(&dopacity[g], (double)(G * dL_dopa));
                }
            }
    }
}

/* DGR/auxiliary.h:107-118 dnormvdv(float3) */
static void dnormvdv3(const float* v, const float* dv, float* o)
{
    float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
    o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
    o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

/* DGR/backward.cu:20-139 computeColorFromSH (backward). dL_dmean is accumulated (+=), dL_dsh[M][3] assigned. */
static void sh_backward(int deg, int M, const float* mean, const float* campos, const float* sh, const uint8_t* clamped,
                        const float* dL_dcolor, float* dL_dmean, float* dL_dsh)
{
    (void)M;
    float dir_orig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
    float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
    float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
    float dL_dRGB[3], dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
    for (int c = 0; c < 3; c++) dL_dRGB[c] = dL_dcolor[c] * (clamped[c] ? 0.f : 1.f);
#define SHC(k) sh[(k)*3 + c]
#define DSH(k, w) for (int c = 0; c < 3; c++) dL_dsh[(k)*3 + c] = (w)*dL_dRGB[c]
    DSH(0, SH_C0);
    if (deg > 0) {
        float d1 = -SH_C1 * y, d2 = SH_C1 * z, d3 = -SH_C1 * x;
        DSH(1, d1); DSH(2, d2); DSH(3, d3);
        for (int c = 0; c < 3; c++) {
            dRGBdx[c] = -SH_C1 * SHC(3);
            dRGBdy[c] = -SH_C1 * SHC(1);
            dRGBdz[c] = SH_C1 * SHC(2);
        }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            float d4 = SH_C2[0] * xy, d5 = SH_C2[1] * yz, d6 = SH_C2[2] * (2.f * zz - xx - yy), d7 = SH_C2[3] * xz,
                  d8 = SH_C2[4] * (xx - yy);
            DSH(4, d4); DSH(5, d5); DSH(6, d6); DSH(7, d7); DSH(8, d8);
            for (int c = 0; c < 3; c++) {
                dRGBdx[c] += SH_C2[0] * y * SHC(4) + SH_C2[2] * 2.f * -x * SHC(6) + SH_C2[3] * z * SHC(7) + SH_C2[4] * 2.f * x * SHC(8);
                dRGBdy[c] += SH_C2[0] * x * SHC(4) + SH_C2[1] * z * SHC(5) + SH_C2[2] * 2.f * -y * SHC(6) + SH_C2[4] * 2.f * -y * SHC(8);
                dRGBdz[c] += SH_C2[1] * y * SHC(5) + SH_C2[2] * 2.f * 2.f * z * SHC(6) + SH_C2[3] * x * SHC(7);
            }
            if (deg > 2) {
                float d9 = SH_C3[0] * y * (3.f * xx - yy), d10 = SH_C3[1] * xy * z, d11 = SH_C3[2] * y * (4.f * zz - xx - yy),
                      d12 = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy), d13 = SH_C3[4] * x * (4.f * zz - xx - yy),
                      d14 = SH_C3[5] * z * (xx - yy), d15 = SH_C3[6] * x * (xx - 3.f * yy);
                DSH(9, d9); DSH(10, d10); DSH(11, d11); DSH(12, d12); DSH(13, d13); DSH(14, d14); DSH(15, d15);
                for (int c = 0; c < 3; c++) {
                    dRGBdx[c] += (SH_C3[0] * SHC(9) * 3.f * 2.f * xy + SH_C3[1] * SHC(10) * yz + SH_C3[2] * SHC(11) * -2.f * xy +
                                  SH_C3[3] * SHC(12) * -3.f * 2.f * xz + SH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
                                  SH_C3[5] * SHC(14) * 2.f * xz + SH_C3[6] * SHC(15) * 3.f * (xx - yy));
                    dRGBdy[c] += (SH_C3[0] * SHC(9) * 3.f * (xx - yy) + SH_C3[1] * SHC(10) * xz +
                                  SH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHC(12) * -3.f * 2.f * yz +
                                  SH_C3[4] * SHC(13) * -2.f * xy + SH_C3[5] * SHC(14) * -2.f * yz + SH_C3[6] * SHC(15) * -3.f * 2.f * xy);
                    dRGBdz[c] += (SH_C3[1] * SHC(10) * xy + SH_C3[2] * SHC(11) * 4.f * 2.f * yz +
                                  SH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHC(13) * 4.f * 2.f * xz +
                                  SH_C3[5] * SHC(14) * (xx - yy));
                }
            }
        }
    }
#undef SHC
#undef DSH
    float dL_ddir[3] = {dRGBdx[0] * dL_dRGB[0] + dRGBdx[1] * dL_dRGB[1] + dRGBdx[2] * dL_dRGB[2],
                        dRGBdy[0] * dL_dRGB[0] + dRGBdy[1] * dL_dRGB[1] + dRGBdy[2] * dL_dRGB[2],
                        dRGBdz[0] * dL_dRGB[0] + dRGBdz[1] * dL_dRGB[1] + dRGBdz[2] * dL_dRGB[2]};
    float dm[3];
    dnormvdv3(dir_orig, dL_ddir, dm);
    dL_dmean[0] += dm[0];
    dL_dmean[1] += dm[1];
    dL_dmean[2] += dm[2];
}

/* DGR/backward.cu:278-341 computeCov3D (backward): gradients w.r.t. scale and the un-normalised quaternion */
static void cov3d_backward(const float* scale, float mod, const float* rot, const float* dL_dcov3D, float* dL_dscale,
                           float* dL_drot)
{
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    m3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    m3 S = m3_cols(1, 0, 0, 0, 1, 0, 0, 0, 1);
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    S.c[0][0] = s[0]; S.c[1][1] = s[1]; S.c[2][2] = s[2];
    m3 M = m3_mul(&S, &R);
    m3 dL_dSigma = m3_cols(dL_dcov3D[0], 0.5f * dL_dcov3D[1], 0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[1], dL_dcov3D[3],
                           0.5f * dL_dcov3D[4], 0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[4], dL_dcov3D[5]);
    m3 M2; /* 2.0f * M */
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) M2.c[j][i] = M.c[j][i] * 2.0f;
    m3 dL_dM = m3_mul(&M2, &dL_dSigma);
    m3 Rt = m3_t(&R);
    m3 dL_dMt = m3_t(&dL_dM);
    for (int k = 0; k < 3; k++)
        dL_dscale[k] = Rt.c[k][0] * dL_dMt.c[k][0] + Rt.c[k][1] * dL_dMt.c[k][1] + Rt.c[k][2] * dL_dMt.c[k][2];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) dL_dMt.c[k][i] *= s[k];
#define D(a, b) dL_dMt.c[a][b]
    dL_drot[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
    dL_drot[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
    dL_drot[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * y * (D(2, 2) + D(0, 0));
    dL_drot[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
#undef D
}

/*
 * DGR/backward.cu:144-274 computeCov2DCUDA followed by :346-412 preprocessCUDA (backward), per Gaussian.
 * Inputs dL_dmean2D[P][2], dL_dconic[P][3] (x,y,w), dL_dcolor[P][3], dL_ddepth[P] are fp32 (the caller
 * rounds the double sums of orc_render_backward). All outputs are dense and zero for invisible Gaussians
 * (DGR/../rasterize_points.cu:166-177 zero-initialises them). cov3Ds = cov3D_precomp or the forward's.
 * dL_dcov3D is always produced; dL_dsh needs shs; dL_dscale/dL_drot need scales.
 */
void orc_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations, float scale_modifier,
                             const float* cov3Ds, const float* view, const float* proj, const float* campos, int W, int H,
                             float tan_fovx, float tan_fovy, const float* dL_dmean2D, const float* dL_dconic,
                             const float* dL_dcolor, const float* dL_ddepth, float* dL_dmeans, float* dL_dcov3D,
                             float* dL_dsh, float* dL_dscale, float* dL_drot)
{
    const float h_y = H / (2.0f * tan_fovy);
    const float h_x = W / (2.0f * tan_fovx);
    memset(dL_dmeans, 0, (size_t)P * 3 * sizeof(float));
    memset(dL_dcov3D, 0, (size_t)P * 6 * sizeof(float));
    if (dL_dsh) memset(dL_dsh, 0, (size_t)P * M * 3 * sizeof(float));
    if (dL_dscale) memset(dL_dscale, 0, (size_t)P * 3 * sizeof(float));
    if (dL_drot) memset(dL_drot, 0, (size_t)P * 4 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* mean = means3D + 3 * (size_t)idx;
        /* ---- computeCov2DCUDA, backward.cu:159-273 ---- */
        const float* cov3D = cov3Ds + 6 * (size_t)idx;
        const float dLc[3] = {dL_dconic[3 * (size_t)idx], dL_dconic[3 * (size_t)idx + 1], dL_dconic[3 * (size_t)idx + 2]};
        float cv[3];
        cov2d_ctx cx;
        cov2d(mean, h_x, h_y, tan_fovx, tan_fovy, cov3D, view, cv, &cx);
        const float x_grad_mul = (cx.txtz < -cx.limx || cx.txtz > cx.limx) ? 0.f : 1.f;
        const float y_grad_mul = (cx.tytz < -cx.limy || cx.tytz > cx.limy) ? 0.f : 1.f;
        const float a = cv[0], b = cv[1], c = cv[2];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * (size_t)idx;
#define T_(i, j) cx.T.c[i][j]
#define V_(i, j) cx.Vrk.c[i][j]
#define W_(i, j) cx.W.c[i][j]
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dLc[0] + 2 * b * c * dLc[1] + (denom - a * c) * dLc[2]);
            dL_dc = denom2inv * (-a * a * dLc[2] + 2 * a * b * dLc[1] + (denom - a * c) * dLc[0]);
            dL_db = denom2inv * 2 * (b * c * dLc[0] - (denom + 2 * b * b) * dLc[1] + a * b * dLc[2]);
            dcov[0] = (T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(1, 0) * dL_db + T_(1, 0) * T_(1, 0) * dL_dc);
            dcov[3] = (T_(0, 1) * T_(0, 1) * dL_da + T_(0, 1) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc);
            dcov[5] = (T_(0, 2) * T_(0, 2) * dL_da + T_(0, 2) * T_(1, 2) * dL_db + T_(1, 2) * T_(1, 2) * dL_dc);
            dcov[1] = 2 * T_(0, 0) * T_(0, 1) * dL_da + (T_(0, 0) * T_(1, 1) + T_(0, 1) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 1) * dL_dc;
            dcov[2] = 2 * T_(0, 0) * T_(0, 2) * dL_da + (T_(0, 0) * T_(1, 2) + T_(0, 2) * T_(1, 0)) * dL_db + 2 * T_(1, 0) * T_(1, 2) * dL_dc;
            dcov[4] = 2 * T_(0, 2) * T_(0, 1) * dL_da + (T_(0, 1) * T_(1, 2) + T_(0, 2) * T_(1, 1)) * dL_db + 2 * T_(1, 1) * T_(1, 2) * dL_dc;
        } else {
            for (int i = 0; i < 6; i++) dcov[i] = 0;
        }
        float dL_dT00 = 2 * (T_(0, 0) * V_(0, 0) + T_(0, 1) * V_(0, 1) + T_(0, 2) * V_(0, 2)) * dL_da +
                        (T_(1, 0) * V_(0, 0) + T_(1, 1) * V_(0, 1) + T_(1, 2) * V_(0, 2)) * dL_db;
        float dL_dT01 = 2 * (T_(0, 0) * V_(1, 0) + T_(0, 1) * V_(1, 1) + T_(0, 2) * V_(1, 2)) * dL_da +
                        (T_(1, 0) * V_(1, 0) + T_(1, 1) * V_(1, 1) + T_(1, 2) * V_(1, 2)) * dL_db;
        float dL_dT02 = 2 * (T_(0, 0) * V_(2, 0) + T_(0, 1) * V_(2, 1) + T_(0, 2) * V_(2, 2)) * dL_da +
                        (T_(1, 0) * V_(2, 0) + T_(1, 1) * V_(2, 1) + T_(1, 2) * V_(2, 2)) * dL_db;
        float dL_dT10 = 2 * (T_(1, 0) * V_(0, 0) + T_(1, 1) * V_(0, 1) + T_(1, 2) * V_(0, 2)) * dL_dc +
                        (T_(0, 0) * V_(0, 0) + T_(0, 1) * V_(0, 1) + T_(0, 2) * V_(0, 2)) * dL_db;
        float dL_dT11 = 2 * (T_(1, 0) * V_(1, 0) + T_(1, 1) * V_(1, 1) + T_(1, 2) * V_(1, 2)) * dL_dc +
                        (T_(0, 0) * V_(1, 0) + T_(0, 1) * V_(1, 1) + T_(0, 2) * V_(1, 2)) * dL_db;
        float dL_dT12 = 2 * (T_(1, 0) * V_(2, 0) + T_(1, 1) * V_(2, 1) + T_(1, 2) * V_(2, 2)) * dL_dc +
                        (T_(0, 0) * V_(2, 0) + T_(0, 1) * V_(2, 1) + T_(0, 2) * V_(2, 2)) * dL_db;
        float dL_dJ00 = W_(0, 0) * dL_dT00 + W_(0, 1) * dL_dT01 + W_(0, 2) * dL_dT02;
        float dL_dJ02 = W_(2, 0) * dL_dT00 + W_(2, 1) * dL_dT01 + W_(2, 2) * dL_dT02;
        float dL_dJ11 = W_(1, 0) * dL_dT10 + W_(1, 1) * dL_dT11 + W_(1, 2) * dL_dT12;
        float dL_dJ12 = W_(2, 0) * dL_dT10 + W_(2, 1) * dL_dT11 + W_(2, 2) * dL_dT12;
#undef T_
#undef V_
#undef W_
        float tz = 1.f / cx.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * cx.tx) * tz3 * dL_dJ02 + (2 * h_y * cx.ty) * tz3 * dL_dJ12;
        /* transformVec4x3Transpose, auxiliary.h:89-97 */
        float dm[3] = {view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz, view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
                       view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz};
        float* dmean = dL_dmeans + 3 * (size_t)idx;
        dmean[0] = dm[0]; dmean[1] = dm[1]; dmean[2] = dm[2]; /* assignment, backward.cu:273 */

        /* ---- preprocessCUDA backward, backward.cu:372-411 ---- */
        float mh[4];
        xform4x4(mean, proj, mh);
        float m_w = 1.0f / (mh[3] + 0.0000001f);
        float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[2 * (size_t)idx], g2y = dL_dmean2D[2 * (size_t)idx + 1];
        float d1[3];
        d1[0] = (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        d1[1] = (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        d1[2] = (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        for (int i = 0; i < 3; i++) dmean[i] += d1[i];
        /* depth path, backward.cu:394-403 */
        float mul3 = view[2] * mean[0] + view[6] * mean[1] + view[10] * mean[2] + view[14];
        float d2[3] = {(view[2] - view[3] * mul3) * dL_ddepth[idx], (view[6] - view[7] * mul3) * dL_ddepth[idx],
                       (view[10] - view[11] * mul3) * dL_ddepth[idx]};
        for (int i = 0; i < 3; i++) dmean[i] += d2[i];
        if (shs)
            sh_backward(D, M, mean, campos, shs + (size_t)idx * M * 3, clamped + 3 * (size_t)idx, dL_dcolor + 3 * (size_t)idx,
                        dmean, dL_dsh + (size_t)idx * M * 3);
        if (scales)
            cov3d_backward(scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx, dcov, dL_dscale + 3 * (size_t)idx,
                           dL_drot + 4 * (size_t)idx);
    }
}

/* DGR/rasterizer_impl.cu:54-66 checkFrustum */
void orc_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present)
{
    (void)proj;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        float pv[3];
        xform4x3(means3D + 3 * (size_t)idx, view, pv);
        present[idx] = pv[2] > 0.2f;
    }
}

/* ------------------------------------------------------------------------------------ simple-knn */
/* KNN/simple_knn.cu:45-52 */
static uint32_t prep_morton(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
static uint32_t f2u(float f)
{
    if (!(f > 0.0f)) return 0; /* cvt.rzi.u32.f32 saturates; NaN -> 0 */
    if (f >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)f;
}
typedef struct { float mn[3], mx[3]; } box_t;

/* KNN/simple_knn.cu:119-129 distBoxPoint */
static float dist_box_point(const box_t* b, const float* p)
{
    float diff[3] = {0, 0, 0};
    for (int a = 0; a < 3; a++)
        if (p[a] < b->mn[a] || p[a] > b->mx[a]) diff[a] = fminf(fabsf(p[a] - b->mn[a]), fabsf(p[a] - b->mx[a]));
    return diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2];
}
/* KNN/simple_knn.cu:131-145 updateKBest<3> */
static void update3(const float* ref, const float* pt, float* knn)
{
    float d[3] = {pt[0] - ref[0], pt[1] - ref[1], pt[2] - ref[2]};
    float dist = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    for (int j = 0; j < 3; j++)
        if (knn[j] > dist) {
            float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
}

/*
 * KNN/simple_knn.cu:185-221 SimpleKNN::knn : AABB (init value (0,0,0) participates, :191-200), 30-bit Morton
 * codes (:54-70), stable sort by code (:206-213), boxes of 1024 sorted points (:78-117), box-pruned exact
 * 3-NN mean of squared distances (:147-183). Output indexed by original point id.
 */
#define KNN_BOX 1024
void orc_knn_dist2(int P, const float* points, float* mean_dists)
{
    if (P <= 0) return;
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = 0; i < P; i++)
        for (int a = 0; a < 3; a++) {
            mn[a] = fminf(mn[a], points[3 * (size_t)i + a]);
            mx[a] = fmaxf(mx[a], points[3 * (size_t)i + a]);
        }
    uint64_t* keys = (uint64_t*)malloc((size_t)P * sizeof(uint64_t));
    uint64_t* keys_s = (uint64_t*)malloc((size_t)P * sizeof(uint64_t));
    uint32_t* idx = (uint32_t*)malloc((size_t)P * sizeof(uint32_t));
    uint32_t* order = (uint32_t*)malloc((size_t)P * sizeof(uint32_t));
    for (int i = 0; i < P; i++) {
        uint32_t q[3];
        for (int a = 0; a < 3; a++)
            q[a] = prep_morton(f2u(((points[3 * (size_t)i + a] - mn[a]) / (mx[a] - mn[a])) * (float)((1 << 10) - 1)));
        keys[i] = q[0] | (q[1] << 1) | (q[2] << 2);
        idx[i] = (uint32_t)i;
    }
    orc_sort_pairs((size_t)P, keys, keys_s, idx, order, 32);
    const int nb = (P + KNN_BOX - 1) / KNN_BOX;
    box_t* boxes = (box_t*)malloc((size_t)nb * sizeof(box_t));
    for (int b = 0; b < nb; b++) {
        box_t bx = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
        for (int i = b * KNN_BOX; i < P && i < (b + 1) * KNN_BOX; i++)
            for (int a = 0; a < 3; a++) {
                float v = points[3 * (size_t)order[i] + a];
                bx.mn[a] = fminf(bx.mn[a], v);
                bx.mx[a] = fmaxf(bx.mx[a], v);
            }
        boxes[b] = bx;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < P; i++) {
        const float* pt = points + 3 * (size_t)order[i];
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        int lo = i - 3 < 0 ? 0 : i - 3, hi = i + 3 > P - 1 ? P - 1 : i + 3;
        for (int j = lo; j <= hi; j++)
            if (j != i) update3(pt, points + 3 * (size_t)order[j], best);
        float reject = best[2];
        best[0] = best[1] = best[2] = FLT_MAX;
        for (int b = 0; b < nb; b++) {
            float dist = dist_box_point(&boxes[b], pt);
            if (dist > reject || dist > best[2]) continue;
            int e = (b + 1) * KNN_BOX < P ? (b + 1) * KNN_BOX : P;
            for (int j = b * KNN_BOX; j < e; j++)
                if (j != i) update3(pt, points + 3 * (size_t)order[j], best);
        }
        mean_dists[order[i]] = (best[0] + best[1] + best[2]) / 3.0f;
    }
    free(keys); free(keys_s); free(idx); free(order); free(boxes);
}

int orc_abi_version(void) { return 1; }
