// extern "C" entry points of libgsr (include/gsr.h) and the host-side orchestration of the stages.
// Host logic mirrors CudaRasterizer::Rasterizer::forward/backward (cuda_rasterizer/rasterizer_impl.cu:198-458)
// and the torch glue of rasterize_points.cu:35-242, without torch: raw device pointers + a stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "gsr_common.cuh"

namespace gsr
{
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return GSR_ERR_CUDA;
}

int after_launch(cudaStream_t s, bool debug, const char* stage)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) after stage '%s'", (int)e, cudaGetErrorString(e), stage);
    return GSR_ERR_CUDA;
}

// ---- optional per-stage timing (bench.py roofline block) ----
struct StageTimer
{
    bool enabled = false;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    float ms[GSR_STAGE_COUNT] = {0};
    const char* out_names[GSR_STAGE_COUNT] = {nullptr};
    int count = 0;

    void begin(cudaStream_t s)
    {
        names.clear();
        if (!enabled) return;
        mark(s, "start");
    }
    void mark(cudaStream_t s, const char* name)
    {
        if (!enabled) return;
        if (names.size() >= ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        cudaEventRecord(ev[names.size()], s);
        names.push_back(name);
    }
    void finish(cudaStream_t s, int first_slot)
    {
        if (!enabled || names.size() < 2) return;
        cudaStreamSynchronize(s);
        for (size_t i = 1; i < names.size() && first_slot + (int)i - 1 < GSR_STAGE_COUNT; i++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[i - 1], ev[i]);
            ms[first_slot + i - 1] = t;
            out_names[first_slot + i - 1] = names[i];
            if (first_slot + (int)i > count) count = first_slot + (int)i;
        }
    }
};
// process-global on purpose: autograd runs gsr_backward on its own worker thread, bench.py reads the times from the main
// thread. Profiling mode is a single-stream diagnostic (bench.py), not meant for concurrent callers.
static StageTimer g_timer;
static const int kBwdFirstSlot = 10;

// Pinned landing zone + event for the forward's counter read-back, one per (thread, device).
struct HostSync
{
    uint32_t* pinned = nullptr;
    cudaEvent_t ev = nullptr;
    size_t last_bin_bytes = 0; // binning state of the previous forward on this (thread, device): the next one's speculative size
};
static HostSync* host_sync()
{
    static thread_local HostSync cache[64];
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice") || dev < 0 || dev >= 64) return nullptr;
    HostSync& h = cache[dev];
    if (!h.pinned) {
        if (check_cuda(cudaHostAlloc((void**)&h.pinned, 64, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
        if (check_cuda(cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    }
    return &h;
}

static thread_local uint32_t g_last_visible = 0;

// Side stream (+ fork/join events) for work that is independent of the main stream's current kernel, one per (thread, device).
struct SideStream
{
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr; // profiling mode: the fills' own duration on the side stream
};
static SideStream* side_stream()
{
    static thread_local SideStream cache[64];
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice") || dev < 0 || dev >= 64) return nullptr;
    SideStream& h = cache[dev];
    if (!h.stream) {
        if (check_cuda(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return nullptr;
        if (check_cuda(cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
        if (check_cuda(cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
        if (check_cuda(cudaEventCreate(&h.t0), "cudaEventCreate") || check_cuda(cudaEventCreate(&h.t1), "cudaEventCreate")) return nullptr;
    }
    return &h;
}

static int ceil_log2(uint32_t v)
{
    int b = 0;
    while (b < 32 && (1ull << b) < v) b++;
    return b;
}

// Fused sub-scenes (GsrGaussians.parts): check the part list and return, in `eff`, the Gaussians struct the rest of the host
// code reasons about (which optional members exist) -- the caller's own struct, or part 0's pointers with the fused P.
static int resolve_parts(const GsrGaussians* in, GsrGaussians& eff)
{
    eff = *in;
    if (in->num_parts <= 0) {
        eff.num_parts = 0;
        eff.parts = nullptr;
        return 0;
    }
    if (in->num_parts > GSR_MAX_PARTS || !in->parts) {
        set_error("num_parts=%d: at most %d parts, and `parts` must be set", in->num_parts, GSR_MAX_PARTS);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in->raw_params || in->subset) {
        set_error("fused sub-scenes (parts) are not combined with raw_params or subset");
        return GSR_ERR_UNSUPPORTED;
    }
    long long total = 0;
    const GsrGaussians& p0 = in->parts[0];
    for (int k = 0; k < in->num_parts; k++) {
        const GsrGaussians& pk = in->parts[k];
        if (pk.P <= 0 || pk.num_parts > 0 || pk.subset || pk.raw_params) {
            set_error("part %d: P must be > 0 and parts do not nest / carry subset / raw_params", k);
            return GSR_ERR_INVALID_ARGUMENT;
        }
        const bool same = (pk.means3D != nullptr) == (p0.means3D != nullptr) && (pk.shs != nullptr) == (p0.shs != nullptr) &&
                          (pk.colors_precomp != nullptr) == (p0.colors_precomp != nullptr) && (pk.segments != nullptr) == (p0.segments != nullptr) &&
                          (pk.opacities != nullptr) == (p0.opacities != nullptr) && (pk.scales != nullptr) == (p0.scales != nullptr) &&
                          (pk.rotations != nullptr) == (p0.rotations != nullptr) && (pk.cov3D_precomp != nullptr) == (p0.cov3D_precomp != nullptr);
        if (!same) {
            set_error("part %d provides a different set of members than part 0", k);
            return GSR_ERR_INVALID_ARGUMENT;
        }
        total += pk.P;
    }
    if (total != (long long)in->P) {
        set_error("P=%d is not the sum of the parts' sizes (%lld)", in->P, total);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    eff = p0;
    eff.P = in->P;
    eff.parts = in->parts;
    eff.num_parts = in->num_parts;
    return 0;
}

static int validate(const GsrView* view, const GsrGaussians* in)
{
    if (!view || !in) {
        set_error("null view/gaussians");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in->P < 0 || view->image_width <= 0 || view->image_height <= 0) {
        set_error("invalid sizes P=%d W=%d H=%d", in->P, view->image_width, view->image_height);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in->P == 0) return 0;
    if (!in->means3D || !in->opacities || !view->bg || !view->viewmatrix || !view->projmatrix) {
        set_error("means3D, opacities, bg, viewmatrix and projmatrix are required");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if ((in->shs == nullptr) == (in->colors_precomp == nullptr)) {
        set_error("Please provide excatly one of either SHs or precomputed colors!");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const bool sr = in->scales != nullptr && in->rotations != nullptr;
    if (sr == (in->cov3D_precomp != nullptr) || ((in->scales != nullptr) != (in->rotations != nullptr))) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in->shs) {
        if (!view->campos) {
            set_error("campos is required with SHs");
            return GSR_ERR_INVALID_ARGUMENT;
        }
        const int need = (view->sh_degree + 1) * (view->sh_degree + 1);
        if (view->sh_degree < 0 || view->sh_degree > 3 || view->sh_coeffs < need || view->sh_coeffs > 16) {
            set_error("sh_degree=%d needs %d <= sh_coeffs <= 16, got %d", view->sh_degree, need, view->sh_coeffs);
            return GSR_ERR_INVALID_ARGUMENT;
        }
    }
    if (in->subset && in->subset_count < 0) {
        set_error("subset_count must be >= 0");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in->raw_params) {
        if (in->cov3D_precomp || !sr) {
            set_error("raw_params needs scales + rotations (log-scales, un-normalised quaternions), not cov3D_precomp");
            return GSR_ERR_INVALID_ARGUMENT;
        }
        if (in->shs && view->sh_coeffs > 1 && !in->shs_rest) {
            set_error("raw_params with SHs needs shs (features_dc [P,1,3]) and shs_rest (features_rest [P,M-1,3])");
            return GSR_ERR_INVALID_ARGUMENT;
        }
    }
    if (view->num_class < 0 || view->num_class > GSR_MAX_NUM_CLASS) {
        set_error("num_class=%d: 0 (no segments) .. %d", view->num_class, GSR_MAX_NUM_CLASS);
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (view->num_class > 2 && in->num_parts > 0) {
        set_error("fused sub-scenes (parts) render at most 2 segment classes");
        return GSR_ERR_UNSUPPORTED;
    }
    const int gx = (view->image_width + TILE_X - 1) / TILE_X, gy = (view->image_height + TILE_Y - 1) / TILE_Y;
    if (gx > 65535 || gy > 65535) {
        set_error("image too large: tile grid %dx%d exceeds 65535", gx, gy);
        return GSR_ERR_UNSUPPORTED;
    }
    return 0;
}
} // namespace gsr

using namespace gsr;

extern "C" int gsr_abi_version(void) { return GSR_ABI_VERSION; }
extern "C" const char* gsr_last_error(void) { return g_err; }
extern "C" unsigned long long gsr_launch_count(void) { return gsr::launches(); }
extern "C" void gsr_set_profiling(int enable) { g_timer.enabled = enable != 0; }
extern "C" int gsr_get_stage_times(float* ms, const char** names)
{
    for (int i = 0; i < GSR_STAGE_COUNT; i++) {
        ms[i] = i < g_timer.count ? g_timer.ms[i] : 0.f;
        names[i] = i < g_timer.count && g_timer.out_names[i] ? g_timer.out_names[i] : "";
    }
    return g_timer.count;
}

extern "C" int gsr_forward(const GsrView* view, const GsrGaussians* in_, const GsrOutputs* out, gsr_alloc_fn alloc, void* alloc_user,
                           int32_t* num_rendered, gsr_stream_t stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (!view || !in_) {
        set_error("null view/gaussians");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    GsrGaussians eff;
    int rc = resolve_parts(in_, eff);
    if (rc) return rc;
    const GsrGaussians* in = &eff;
    rc = validate(view, in);
    if (rc) return rc;
    if (num_rendered) *num_rendered = 0;
    if (in->P == 0 || (in->subset && in->subset_count == 0)) return 0;
    if (!out || !out->color || !out->depth || !out->alpha || !out->radii || (view->num_class > 0 && !out->segment) || !alloc) {
        set_error("missing output pointers or allocator");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const bool debug = view->debug != 0;
    const int P = in->subset ? in->subset_count : in->P; // Gaussians rendered (= positions of the per-Gaussian state and of radii)
    const int W = view->image_width, H = view->image_height;
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const uint32_t T = (uint32_t)gx * (uint32_t)gy;
    const size_t N = (size_t)W * H;

    if (g_timer.enabled) {
        g_timer.count = 0;
        for (int i = 0; i < GSR_STAGE_COUNT; i++) { g_timer.ms[i] = 0.f; g_timer.out_names[i] = nullptr; }
    }
    g_timer.begin(s);

    // ---- state buffers whose size is known up front ----
    GeomState g;
    const size_t geom_bytes = geom_layout(nullptr, P, g, view->num_class);
    char* geom_base = (char*)alloc(alloc_user, GSR_BUF_GEOM, geom_bytes);
    ImgState img;
    const size_t img_bytes = img_layout(nullptr, N, T, img);
    char* img_base = (char*)alloc(alloc_user, GSR_BUF_IMG, img_bytes);
    if (!geom_base || !img_base) {
        set_error("state allocation failed (geom %zu B, img %zu B)", geom_bytes, img_bytes);
        return GSR_ERR_ALLOC;
    }
    geom_layout(geom_base, P, g, view->num_class);
    img_layout(img_base, N, T, img);

    GSR_CUDA(cudaMemsetAsync(g.counters, 0, CNT_WORDS * sizeof(uint32_t), s));

    // ---- per-Gaussian preprocess ----
    PreFwdArgs pa;
    pa.P = P; pa.D = view->sh_degree; pa.M = in->shs ? view->sh_coeffs : 0; pa.S = in->segments ? view->num_class : 0;
    pa.means3D = in->means3D; pa.scales = in->scales; pa.scale_modifier = view->scale_modifier; pa.rotations = in->rotations;
    pa.opacities = in->opacities; pa.shs = in->shs; pa.cov3D_precomp = in->cov3D_precomp; pa.colors_precomp = in->colors_precomp;
    pa.shs_rest = in->shs_rest; pa.raw = in->raw_params; pa.subset = in->subset;
    pa.segments = in->segments; pa.view = view->viewmatrix; pa.proj = view->projmatrix; pa.campos = view->campos;
    pa.W = W; pa.H = H; pa.tan_fovx = view->tanfovx; pa.tan_fovy = view->tanfovy;
    pa.focal_y = H / (2.0f * view->tanfovy); // rasterizer_impl.cu:226-227
    pa.focal_x = W / (2.0f * view->tanfovx);
    pa.grid_x = gx; pa.grid_y = gy; pa.prefiltered = view->prefiltered; pa.radii = out->radii; pa.g = g;
    pa.num_parts = in->num_parts;
    pa.part_start[0] = 0;
    for (int k = 0; k < in->num_parts; k++) {
        const GsrGaussians& pk = in->parts[k];
        pa.part[k] = {pk.means3D, pk.scales, pk.rotations, pk.opacities, pk.shs, pk.cov3D_precomp, pk.colors_precomp, pk.segments};
        pa.part_start[k + 1] = pa.part_start[k] + pk.P;
    }
    launch_preprocess_fwd(pa, s);
    GSR_LAUNCHED(s, debug, "preprocess_fwd");
    launch_block_offsets(g, s);
    GSR_LAUNCHED(s, debug, "block_offsets");
    g_timer.mark(s, "preprocess_fwd");

    // ---- the one host synchronisation: V and R size the binning state (reference: rasterizer_impl.cu:285). The counters are
    // copied to pinned memory behind an event, and the depth sort -- whose buffers live in the geometry state with capacity P
    // and whose element count V is read on the device -- is enqueued BEFORE waiting, so the GPU keeps working while the host
    // round-trips, allocates the binning state and enqueues the rest.
    HostSync* hs = host_sync();
    if (!hs) return GSR_ERR_CUDA;
    GSR_CUDA(cudaMemcpyAsync(hs->pinned, g.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    GSR_CUDA(cudaEventRecord(hs->ev, s));

    // depth sort workspace: digit totals, tickets and look-back words of its four single-kernel passes, zeroed in one memset
    GSR_CUDA(cudaMemsetAsync(g.dhist, 0, g.dhist_words * sizeof(uint32_t), s));
    launch_depth_keys(g, g.dhist, s);
    GSR_LAUNCHED(s, debug, "depth_keys");
    // the depth sort has few keys (V ~ 1.2 M at cfg3 = 294 tiles of 4096 keys, two per SM): 2048-key tiles give it twice as many CTAs
    // (0.125 -> 0.110 ms at cfg3; the tile sort, 10 M keys, is better off with 4096: 0.146 vs 0.179 ms). GSR_DEPTH_PT=16 for A/B.
    static const int depth_pt = getenv("GSR_DEPTH_PT") ? atoi(getenv("GSR_DEPTH_PT")) : 8;
    int dres = radix_sort_pairs_lookback(g.dkeys, g.dvals, g.slots, 32, g.dhist, g.dhist_words, true, s, g.counters + CNT_VISIBLE, depth_pt);
    if (dres < 0) return dres;
    GSR_LAUNCHED(s, debug, "depth_sort");
    const uint32_t* sorted_slots = g.dvals[dres];

    // The binning state is requested BEFORE the wait, with the previous forward's size + 12.5 % as a guess: the allocator callback
    // (Python, through ctypes: tens of microseconds) then runs while the GPU is busy instead of between the wait and the
    // remaining launches, where the depth sort is all that covers the host. The exact size is requested again only if the guess
    // was too small (the allocator's last answer for a buffer kind is the one that counts).
    char* bin_base = nullptr;
    size_t bin_cap = 0;
    if (hs->last_bin_bytes) {
        bin_cap = hs->last_bin_bytes + hs->last_bin_bytes / 8;
        bin_base = (char*)alloc(alloc_user, GSR_BUF_BINNING, bin_cap);
        if (!bin_base) bin_cap = 0;
    }
    GSR_CUDA(cudaEventSynchronize(hs->ev));
    const uint32_t* counters = hs->pinned;
    if (counters[CNT_ERROR] & 1u) {
        set_error("Point is filtered although prefiltered is set. This shouldn't happen!");
        return GSR_ERR_PREFILTERED;
    }
    const uint32_t V = counters[CNT_VISIBLE];
    g_last_visible = V;
    const unsigned long long R64 = (unsigned long long)counters[CNT_RENDERED_LO] | ((unsigned long long)counters[CNT_RENDERED_LO + 1] << 32);
    if (R64 > 0x7fffffffull) {
        set_error("%llu tile instances exceed the int32 limit", R64);
        return GSR_ERR_OVERFLOW;
    }
    const uint32_t R = (uint32_t)R64;
    if (num_rendered) *num_rendered = (int32_t)R;

    const int tile_bits = ceil_log2(T) < 1 ? 1 : ceil_log2(T);
    BinState b;
    const size_t bin_bytes = bin_layout(nullptr, V, R, tile_bits, b);
    hs->last_bin_bytes = bin_bytes;
    if (bin_bytes > bin_cap) bin_base = (char*)alloc(alloc_user, GSR_BUF_BINNING, bin_bytes);
    if (!bin_base) {
        set_error("binning state allocation failed (%zu B)", bin_bytes);
        return GSR_ERR_ALLOC;
    }
    bin_layout(bin_base, V, R, tile_bits, b);
    g_timer.mark(s, "depth_sort");

    // ---- instances in depth order, then stable sort by tile ----
    GSR_CUDA(cudaMemsetAsync(b.hist, 0, b.zero_bytes, s)); // tile-sort digit totals / tickets / look-back words + the scan's
    rc = launch_instance_offsets(g, b, V, R, sorted_slots, s);
    if (rc) return rc;
    GSR_LAUNCHED(s, debug, "instance_offsets");
    const int tpasses = radix_num_passes(tile_bits);
    // choose the starting buffers so that the final pass lands in point_list (offset 0 of the binning state)
    uint32_t* tvals[2];
    uint32_t* tkeys[2] = {b.tkeys[0], b.tkeys[1]};
    if (tpasses % 2 == 0) { tvals[0] = b.point_list; tvals[1] = b.vals_alt; }
    else { tvals[0] = b.vals_alt; tvals[1] = b.point_list; }
    rc = launch_emit(g, b, V, R, sorted_slots, gx, radix_digit_bits(tile_bits), tpasses, b.hist, tkeys[0], tvals[0], s);
    if (rc) return rc;
    GSR_LAUNCHED(s, debug, "emit");
    g_timer.mark(s, "emit");
    int tres = radix_sort_pairs_lookback(tkeys, tvals, R, tile_bits, b.hist, b.hist_words, true, s);
    if (tres < 0) return tres;
    GSR_LAUNCHED(s, debug, "tile_sort");
    if (R > 0 && tvals[tres] != b.point_list) {
        set_error("internal: tile sort ended in the wrong buffer");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    g_timer.mark(s, "tile_sort");
    static const bool use_tile_order = getenv("GSR_TILE_ORDER") && atoi(getenv("GSR_TILE_ORDER")) != 0;
    rc = launch_tile_ranges(tkeys[tres], R, img.ranges, use_tile_order ? img.tile_order : nullptr, T, s);
    if (rc) return rc;
    GSR_LAUNCHED(s, debug, "tile_ranges");
    g_timer.mark(s, "tile_ranges");

    // ---- compositing ----
    RenderArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.W = W; ra.H = H; ra.grid_x = gx; ra.grid_y = gy;
    ra.ranges = img.ranges; ra.tile_order = use_tile_order ? img.tile_order : nullptr; ra.point_list = b.point_list; ra.rec = g.rec; ra.bg = view->bg;
    ra.out_color = out->color; ra.out_segment = out->segment; ra.out_depth = out->depth; ra.out_alpha = out->alpha;
    ra.n_contrib = img.n_contrib;
    ra.seg_count = view->num_class >= 2 ? 2 : view->num_class;
    launch_render_fwd(ra, view->num_class > 0 ? 2 : 0, s);
    GSR_LAUNCHED(s, debug, "render_fwd");
    // runtime class count (the reference compiles NUM_CLASS = 2 in, config.h:16; its ModelParams default to 29 classes): the
    // record carries channels 0-1, every further pair of channels is composited by one more pass over the same lists
    for (uint32_t k = 0; k < g.extra_pairs; k++) {
        const int c0 = 2 * ((int)k + 1);
        if (!in->segments) {
            GSR_CUDA(cudaMemsetAsync(out->segment + (size_t)c0 * N, 0, (size_t)(view->num_class - c0) * N * sizeof(float), s));
            break;
        }
        RenderArgs rk = ra;
        rk.out_color = nullptr; rk.out_depth = nullptr; rk.out_alpha = nullptr; rk.n_contrib = nullptr;
        rk.out_segment = out->segment + (size_t)c0 * N;
        rk.seg_src = g.seg_extra + (size_t)k * g.slots;
        rk.seg_count = view->num_class - c0 >= 2 ? 2 : 1;
        launch_render_fwd(rk, 2, s);
        GSR_LAUNCHED(s, debug, "render_fwd (segment pair)");
    }
    g_timer.mark(s, "render_fwd");
    g_timer.finish(s, 0);
    return 0;
}

extern "C" size_t gsr_backward_scratch_bytes_n(int32_t P, int32_t num_class)
{
    if (P <= 0) return 0;
    const size_t nblk = ((size_t)P + PRE_BLOCK - 1) / PRE_BLOCK;
    const size_t extra = num_class > 2 ? (size_t)((num_class + 1) / 2 - 1) : 0; // dL/dsegment of the channel pairs beyond the record's
    return nblk * PRE_BLOCK * (GRAD_REC_FLOATS + 2 * extra) * sizeof(float) + 512;
}
extern "C" size_t gsr_backward_scratch_bytes(int32_t P) { return gsr_backward_scratch_bytes_n(P, 2); }

static int backward_impl(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                         const GsrPixelGrads* pix, const GsrParamGrads* grads, void* scratch, size_t scratch_bytes, cudaStream_t s,
                         uint32_t* packets, uint32_t capacity, uint32_t* count_dev, uint32_t* vis_index);

extern "C" int gsr_backward(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                            const GsrPixelGrads* pix, const GsrParamGrads* grads, void* scratch, size_t scratch_bytes, gsr_stream_t stream_)
{
    return backward_impl(view, in, radii, state, alpha, pix, grads, scratch, scratch_bytes, (cudaStream_t)stream_, nullptr, 0, nullptr, nullptr);
}

extern "C" uint32_t gsr_last_num_visible(void) { return g_last_visible; }

extern "C" int gsr_backward_packets(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                                    const GsrPixelGrads* pix, uint32_t* packets, uint32_t capacity, uint32_t* count_dev, float* dL_dmeans2D,
                                    uint32_t* vis_index, void* scratch, size_t scratch_bytes, gsr_stream_t stream_)
{
    if (!packets || !count_dev) {
        set_error("gsr_backward_packets: packets/count_dev missing");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (in && in->P > 0 && (!in->shs || !in->scales || !in->rotations)) {
        set_error("gsr_backward_packets needs shs and scales/rotations");
        return GSR_ERR_UNSUPPORTED;
    }
    if (state && state->num_visible > 0 && (uint32_t)state->num_visible > capacity) {
        set_error("gsr_backward_packets: %d visible Gaussians do not fit %u packets", state->num_visible, capacity);
        return GSR_ERR_OVERFLOW;
    }
    GsrParamGrads g;
    memset(&g, 0, sizeof(g));
    g.dL_dmeans2D = dL_dmeans2D;
    return backward_impl(view, in, radii, state, alpha, pix, &g, scratch, scratch_bytes, (cudaStream_t)stream_, packets, capacity, count_dev,
                         vis_index);
}

extern "C" size_t gsr_packet_index_words(int32_t P) { return P > 0 ? (2 * (size_t)((P + 31) / 32) + 31) / 32 * 32 : 0; }

extern "C" int gsr_gather_packets_v(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t num_class, const float* means3D,
                                    int32_t num_views, const float* campos, const uint32_t* const* view_ptrs, size_t packet_off_words,
                                    size_t index_off_words, uint32_t capacity, const GsrParamGrads* grads, gsr_stream_t stream_)
{
    if (P <= 0) return 0;
    if (!means3D || !campos || !view_ptrs || capacity < 1 || !grads || sh_coeffs > 16 || sh_degree < 0 || sh_degree > 3 || num_views < 1 ||
        num_views > GSR_MAX_GATHER_VIEWS) {
        set_error("gsr_gather_packets: invalid argument (1 <= num_views <= 64)");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    GatherPacketsArgs a;
    a.P = P; a.D = sh_degree; a.M = sh_coeffs; a.S = num_class; a.num_views = num_views; a.means3D = means3D; a.campos = campos;
    for (int v = 0; v < GSR_MAX_GATHER_VIEWS; v++) a.views[v] = view_ptrs[v < num_views ? v : 0];
    for (int v = 0; v < num_views; v++)
        if (!view_ptrs[v]) {
            set_error("gsr_gather_packets: null view blob");
            return GSR_ERR_INVALID_ARGUMENT;
        }
    a.packet_off = packet_off_words; a.index_off = index_off_words; a.capacity = capacity; a.out = *grads;
    launch_gather_packets(a, (cudaStream_t)stream_);
    GSR_LAUNCHED((cudaStream_t)stream_, false, "gather_packets");
    return 0;
}

extern "C" int gsr_gather_packets(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t num_class, const float* means3D, int32_t num_views,
                                  const float* campos, const uint32_t* blobs, size_t blob_stride_words, uint32_t capacity,
                                  const GsrParamGrads* grads, gsr_stream_t stream_)
{
    if (P <= 0) return 0;
    if (!blobs || num_views < 1 || num_views > GSR_MAX_GATHER_VIEWS) {
        set_error("gsr_gather_packets: invalid argument (1 <= num_views <= 64)");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const uint32_t* ptrs[GSR_MAX_GATHER_VIEWS];
    for (int v = 0; v < num_views; v++) ptrs[v] = blobs + (size_t)v * blob_stride_words;
    return gsr_gather_packets_v(P, sh_degree, sh_coeffs, num_class, means3D, num_views, campos, ptrs, gsr_packet_index_words(P), 0, capacity,
                                grads, stream_);
}

extern "C" int gsr_peer_alloc(size_t bytes, void** ptr, void* handle_out)
{
    if (!ptr || !handle_out || bytes == 0) {
        set_error("gsr_peer_alloc: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == GSR_PEER_HANDLE_BYTES, "handle size");
    GSR_CUDA(cudaMalloc(ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        set_error("gsr_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return GSR_ERR_CUDA;
    }
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
extern "C" int gsr_peer_open(const void* handle, void** ptr)
{
    if (!handle || !ptr) {
        set_error("gsr_peer_open: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GSR_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int gsr_peer_close(void* ptr)
{
    if (ptr) GSR_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
extern "C" int gsr_peer_free(void* ptr)
{
    if (ptr) GSR_CUDA(cudaFree(ptr));
    return 0;
}
extern "C" int gsr_peer_copy(void* dst, const void* src, size_t bytes, gsr_stream_t stream_)
{
    if (bytes == 0) return 0;
    if (!dst || !src) {
        set_error("gsr_peer_copy: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    GSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream_));
    return 0;
}

static int backward_impl(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                         const GsrPixelGrads* pix, const GsrParamGrads* grads, void* scratch, size_t scratch_bytes, cudaStream_t s,
                         uint32_t* packets, uint32_t capacity, uint32_t* count_dev, uint32_t* vis_index)
{
    if (in && in->num_parts > 0) {
        set_error("gsr_backward: fused sub-scenes (parts) are render-only");
        return GSR_ERR_UNSUPPORTED;
    }
    int rc = validate(view, in);
    if (rc) return rc;
    if (in->P == 0 || (in->subset && in->subset_count == 0)) return 0;
    if (in->subset && packets) {
        set_error("gsr_backward_packets does not take an index list (subset)");
        return GSR_ERR_UNSUPPORTED;
    }
    const int count = in->subset ? in->subset_count : in->P; // rendered Gaussians: sizes of the saved state
    if (!radii || !state || !state->geom || !state->img || !alpha || !pix || !pix->dL_dcolor || !grads || !scratch) {
        set_error("gsr_backward: missing argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (state->num_rendered > 0 && !state->binning) {
        set_error("gsr_backward: binning state missing");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (scratch_bytes < gsr_backward_scratch_bytes_n(count, view->num_class)) {
        set_error("gsr_backward: scratch too small (%zu < %zu)", scratch_bytes, gsr_backward_scratch_bytes_n(count, view->num_class));
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (packets && view->num_class != 0 && view->num_class != 2) {
        set_error("gsr_backward_packets carries 2 segment classes (num_class=%d)", view->num_class);
        return GSR_ERR_UNSUPPORTED;
    }
    const bool debug = view->debug != 0;
    const int P = in->P, W = view->image_width, H = view->image_height;
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const uint32_t T = (uint32_t)gx * (uint32_t)gy;

    GeomState g;
    geom_layout((char*)state->geom, count, g, view->num_class);
    ImgState img;
    img_layout((char*)state->img, (size_t)W * H, T, img);
    float* grad_rec = (float*)align_up((size_t)scratch, 256);
    float* grad_seg_extra = (float*)align_up((size_t)(grad_rec + (size_t)g.slots * GRAD_REC_FLOATS), 256);

    PreBwdArgs pb;
    pb.P = P; pb.D = view->sh_degree; pb.M = in->shs ? view->sh_coeffs : 0; pb.S = view->num_class;
    pb.means3D = in->means3D; pb.scales = in->scales; pb.scale_modifier = view->scale_modifier; pb.rotations = in->rotations;
    pb.shs_rest = in->shs_rest; pb.raw = in->raw_params; pb.opacities = in->opacities; pb.segments = in->segments;
    pb.shs = in->shs; pb.cov3D_precomp = in->cov3D_precomp; pb.view = view->viewmatrix; pb.proj = view->projmatrix; pb.campos = view->campos;
    pb.W = W; pb.H = H; pb.tan_fovx = view->tanfovx; pb.tan_fovy = view->tanfovy;
    pb.focal_y = H / (2.0f * view->tanfovy);
    pb.focal_x = W / (2.0f * view->tanfovx);
    pb.radii = radii; pb.g = g; pb.grad_rec = grad_rec; pb.grad_seg_extra = reinterpret_cast<const float2*>(grad_seg_extra); pb.out = *grads;
    pb.colors_precomp_given = in->colors_precomp != nullptr;
    pb.packets = packets; pb.packet_capacity = capacity; pb.packet_count = count_dev; pb.vis_index = vis_index;
    pb.has_subset = in->subset != nullptr;
    if (!in->shs) { pb.out.dL_dsh = nullptr; pb.out.dL_dsh_rest = nullptr; }
    if (!in->raw_params) pb.out.dL_dsh_rest = nullptr;
    if (in->raw_params && pb.out.dL_dsh && view->sh_coeffs > 1 && !pb.out.dL_dsh_rest && !packets) {
        set_error("raw_params: dL_dsh ([P,1,3]) needs dL_dsh_rest ([P,M-1,3])");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (!in->scales) { pb.out.dL_dscales = nullptr; pb.out.dL_drotations = nullptr; }

    g_timer.begin(s);
    // The dense zero fills of the output gradients (HBM-write bound, ~1.5 GB at cfg3) are independent of the compositing backward
    // (issue bound). GSR_FILL_STREAM=side (default) forks them onto a side stream so that they overlap it; =main runs them on the
    // launching stream after it (A/B: profiles/).
    static const bool fills_on_side = !(getenv("GSR_FILL_STREAM") && !strcmp(getenv("GSR_FILL_STREAM"), "main"));
    SideStream* ss = side_stream();
    if (!ss) return GSR_ERR_CUDA;
    if (fills_on_side) {
        GSR_CUDA(cudaEventRecord(ss->fork, s));
        GSR_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
        if (g_timer.enabled) GSR_CUDA(cudaEventRecord(ss->t0, ss->stream));
        rc = launch_grad_fills(pb, ss->stream);
        if (rc) return rc;
        if (g_timer.enabled) GSR_CUDA(cudaEventRecord(ss->t1, ss->stream));
        GSR_CUDA(cudaEventRecord(ss->join, ss->stream));
    }

    // per-slot gradient records: only the live slots [256 b, 256 b + blk_count[b]) are read back, so only those are zeroed
    // (58 MB instead of 288 MB at cfg3)
    rc = launch_zero_grad_rec(g, grad_rec, s);
    if (rc) return rc;

    RenderArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.W = W; ra.H = H; ra.grid_x = gx; ra.grid_y = gy;
    static const bool use_tile_order_b = getenv("GSR_TILE_ORDER") && atoi(getenv("GSR_TILE_ORDER")) != 0;
    ra.ranges = img.ranges; ra.tile_order = use_tile_order_b ? img.tile_order : nullptr; ra.point_list = (const uint32_t*)state->binning; ra.rec = g.rec; ra.bg = view->bg;
    ra.n_contrib = img.n_contrib; ra.alphas = alpha;
    ra.dL_dcolor = pix->dL_dcolor; ra.dL_dsegment = pix->dL_dsegment; ra.dL_ddepth = pix->dL_ddepth; ra.dL_dalpha = pix->dL_dalpha;
    ra.grad_rec = grad_rec;
    ra.seg_count = view->num_class >= 2 ? 2 : view->num_class;
    launch_render_bwd(ra, view->num_class > 0 ? 2 : 0, s);
    GSR_LAUNCHED(s, debug, "render_bwd");
    if (g.extra_pairs && in->segments && pix->dL_dsegment) { // one more pass per further pair of segment channels (linear: grad_rec accumulates)
        GSR_CUDA(cudaMemsetAsync(grad_seg_extra, 0, (size_t)g.extra_pairs * g.slots * 2 * sizeof(float), s));
        for (uint32_t k = 0; k < g.extra_pairs; k++) {
            const int c0 = 2 * ((int)k + 1);
            RenderArgs rk = ra;
            rk.dL_dcolor = nullptr; rk.dL_ddepth = nullptr; rk.dL_dalpha = nullptr;
            rk.dL_dsegment = pix->dL_dsegment + (size_t)c0 * W * H;
            rk.seg_src = g.seg_extra + (size_t)k * g.slots;
            rk.seg_count = view->num_class - c0 >= 2 ? 2 : 1;
            rk.grad_seg = grad_seg_extra + (size_t)k * g.slots * 2;
            launch_render_bwd(rk, 2, s);
            GSR_LAUNCHED(s, debug, "render_bwd (segment pair)");
        }
    } else if (g.extra_pairs) {
        GSR_CUDA(cudaMemsetAsync(grad_seg_extra, 0, (size_t)g.extra_pairs * g.slots * 2 * sizeof(float), s));
    }
    g_timer.mark(s, "render_bwd");

    if (fills_on_side) {
        GSR_CUDA(cudaStreamWaitEvent(s, ss->join, 0)); // join
    } else {
        if (g_timer.enabled) GSR_CUDA(cudaEventRecord(ss->t0, s));
        rc = launch_grad_fills(pb, s);
        if (rc) return rc;
        if (g_timer.enabled) GSR_CUDA(cudaEventRecord(ss->t1, s));
    }
    launch_preprocess_bwd(pb, s);
    GSR_LAUNCHED(s, debug, "preprocess_bwd");
    g_timer.mark(s, "preprocess_bwd");
    g_timer.finish(s, kBwdFirstSlot);
    if (g_timer.enabled && g_timer.count < GSR_STAGE_COUNT) { // the stream was synchronised by finish(): the side stream's events are done
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ss->t0, ss->t1) == cudaSuccess) {
            g_timer.ms[g_timer.count] = t;
            g_timer.out_names[g_timer.count] = "grad_fills";
            g_timer.count++;
        } else {
            (void)cudaGetLastError();
        }
    }
    return 0;
}

extern "C" int gsr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                                gsr_stream_t stream_)
{
    (void)projmatrix; // the reference computes the projected point but only tests view-space z (auxiliary.h:154)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("gsr_mark_visible: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_);
    GSR_LAUNCHED((cudaStream_t)stream_, false, "mark_visible");
    return 0;
}

extern "C" size_t gsr_knn_workspace_bytes(int32_t P) { return knn_workspace_bytes(P); }

extern "C" int gsr_knn_dist2(int32_t P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes, gsr_stream_t stream_)
{
    if (P < 0 || (P > 0 && (!points || !mean_dist2 || !workspace))) {
        set_error("gsr_knn_dist2: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return 0;
    if (workspace_bytes < knn_workspace_bytes(P)) {
        set_error("gsr_knn_dist2: workspace too small");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    return knn_run(P, points, mean_dist2, workspace, workspace_bytes, (cudaStream_t)stream_);
}

// ------------------------------------------------------------------------------------------ state export
namespace gsr
{
__global__ void export_gaussians_kernel(GeomState g, uint32_t nslots_blocks, GsrStateExport o)
{
    const uint32_t b = blockIdx.x;
    const uint32_t cnt = g.blk_count[b];
    if (threadIdx.x >= cnt) return;
    const uint32_t slot = b * PRE_BLOCK + threadIdx.x;
    const uint32_t id = g.slot_gid[slot];
    const float4 A = g.rec[3 * (size_t)slot], B = g.rec[3 * (size_t)slot + 1], C = g.rec[3 * (size_t)slot + 2];
    if (o.depths) o.depths[id] = C.y;
    if (o.means2D) { o.means2D[2 * (size_t)id] = A.x; o.means2D[2 * (size_t)id + 1] = A.y; }
    if (o.conic_opacity) {
        o.conic_opacity[4 * (size_t)id + 0] = A.z; o.conic_opacity[4 * (size_t)id + 1] = A.w;
        o.conic_opacity[4 * (size_t)id + 2] = B.x; o.conic_opacity[4 * (size_t)id + 3] = B.y;
    }
    if (o.rgb) { o.rgb[3 * (size_t)id] = B.z; o.rgb[3 * (size_t)id + 1] = B.w; o.rgb[3 * (size_t)id + 2] = C.x; }
    if (o.clamped) {
        const uint8_t cb = g.clamped[slot];
        o.clamped[3 * (size_t)id] = cb & 1; o.clamped[3 * (size_t)id + 1] = (cb >> 1) & 1; o.clamped[3 * (size_t)id + 2] = (cb >> 2) & 1;
    }
    if (o.tiles_touched) {
        const ushort4 r = g.rect[slot];
        o.tiles_touched[id] = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
    }
}

__global__ void export_lists_kernel(GeomState g, const uint32_t* point_list, const uint2* ranges, uint32_t T, GsrStateExport o)
{
    // one CTA per tile: its range of the sorted list
    const uint32_t t = blockIdx.x;
    if (t >= T) return;
    const uint2 r = ranges[t];
    for (uint32_t k = r.x + threadIdx.x; k < r.y; k += blockDim.x) {
        const uint32_t slot = point_list[k];
        if (o.point_list) o.point_list[k] = g.slot_gid[slot];
        if (o.point_keys) o.point_keys[k] = ((uint64_t)t << 32) | (uint64_t)__float_as_uint(g.rec[3 * (size_t)slot + 2].y);
    }
}
} // namespace gsr

extern "C" int gsr_export_state(int32_t P, int32_t W, int32_t H, const GsrState* state, const GsrStateExport* out, gsr_stream_t stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (P <= 0 || !state || !out || !state->geom || !state->img) {
        set_error("gsr_export_state: invalid argument");
        return GSR_ERR_INVALID_ARGUMENT;
    }
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const uint32_t T = (uint32_t)gx * (uint32_t)gy;
    const size_t N = (size_t)W * H;
    GeomState g;
    geom_layout((char*)state->geom, P, g);
    ImgState img;
    img_layout((char*)state->img, N, T, img);
    if (out->depths) GSR_CUDA(cudaMemsetAsync(out->depths, 0, (size_t)P * 4, s));
    if (out->means2D) GSR_CUDA(cudaMemsetAsync(out->means2D, 0, (size_t)P * 8, s));
    if (out->conic_opacity) GSR_CUDA(cudaMemsetAsync(out->conic_opacity, 0, (size_t)P * 16, s));
    if (out->rgb) GSR_CUDA(cudaMemsetAsync(out->rgb, 0, (size_t)P * 12, s));
    if (out->clamped) GSR_CUDA(cudaMemsetAsync(out->clamped, 0, (size_t)P * 3, s));
    if (out->tiles_touched) GSR_CUDA(cudaMemsetAsync(out->tiles_touched, 0, (size_t)P * 4, s));
    export_gaussians_kernel<<<g.nblk, PRE_BLOCK, 0, s>>>(g, g.nblk, *out); count_launches(1);
    GSR_LAUNCHED(s, false, "export_gaussians");
    if ((out->point_list || out->point_keys) && state->num_rendered > 0) {
        export_lists_kernel<<<T, 256, 0, s>>>(g, (const uint32_t*)state->binning, img.ranges, T, *out); count_launches(1);
        GSR_LAUNCHED(s, false, "export_lists");
    }
    if (out->ranges) GSR_CUDA(cudaMemcpyAsync(out->ranges, img.ranges, (size_t)T * 8, cudaMemcpyDeviceToDevice, s));
    if (out->n_contrib) GSR_CUDA(cudaMemcpyAsync(out->n_contrib, img.n_contrib, N * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}
