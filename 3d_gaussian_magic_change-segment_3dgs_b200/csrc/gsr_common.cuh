// Shared declarations of libgsr: state layouts, launch-side structs, error helpers.
// sm_100a only. No torch types anywhere below the C ABI (include/gsr.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gsr.h"

namespace gsr
{
constexpr int TILE_X = 16;            // reference BLOCK_X / BLOCK_Y (cuda_rasterizer/config.h:17-18)
constexpr int TILE_Y = 16;
constexpr int TILE_PIXELS = TILE_X * TILE_Y;
constexpr int PRE_BLOCK = 256;        // Gaussians per preprocess CTA == slots per slot-block
constexpr int GRAD_REC_FLOATS = 12;   // per-Gaussian record accumulated by the compositing backward

// ---- header counters at the start of the geometry state ----
enum { CNT_VISIBLE = 0, CNT_RENDERED_LO = 2, CNT_ERROR = 4, CNT_WORDS = 64 };

// Per-visible-Gaussian record, 48 B, in a "slot": slot = block*256 + rank of the Gaussian among the visible
// ones of its preprocess block. Slots of a block are contiguous, so the state stays dense in 128-B lines
// without a global compaction pass. Fields (what the reference keeps as means2D/conic_opacity/rgb/depths):
//   A = { mean2D.x, mean2D.y, conic.x, conic.y }
//   B = { conic.z,  opacity,  rgb.r,   rgb.g   }
//   C = { rgb.b,    depth,    seg0,    seg1    }
struct GeomState
{
    uint32_t* counters;  // [CNT_WORDS]
    float4* rec;         // [3 * slots]
    ushort4* rect;       // [slots] tile rectangle {xmin, ymin, xmax, ymax} (auxiliary.h:46-56)
    uint32_t* slot_gid;  // [slots] Gaussian id of the slot
    uint8_t* clamped;    // [slots] bit c set <=> SH colour channel c was clamped (forward.cu:67-69)
    uint32_t* blk_count; // [nblk] visible Gaussians per slot-block
    uint32_t* blk_offset;// [nblk] exclusive scan of blk_count
    uint32_t* dkeys[2];  // [slots] depth bits of the visible Gaussians (ping-pong), first V used
    uint32_t* dvals[2];  // [slots] their slots (ping-pong)
    uint32_t* dhist;     // depth-sort histograms + scan partials
    size_t dhist_words;
    uint32_t nblk;
    uint32_t slots;      // nblk * PRE_BLOCK
    // num_class > 2: the record carries segment channels 0 and 1; channel pair k >= 1 of a slot lives in seg_extra[(k-1)*slots + slot]
    // and is composited by one extra pass per pair (runtime class count; the reference compiles NUM_CLASS in, config.h:16)
    float2* seg_extra;
    uint32_t extra_pairs; // ceil(num_class / 2) - 1, or 0
};

struct BinState
{
    uint32_t* point_list; // [R] final per-tile lists (slots), depth-sorted; MUST be first (gsr_backward relies on offset 0)
    uint32_t* vals_alt;   // [R]
    uint32_t* tkeys[2];   // [R] tile ids (ping-pong)
    uint32_t* soff;       // [V+1] exclusive scan of tiles_touched in depth order
    uint32_t* hist;       // tile sort workspace: [passes][256] digit totals | 32 tickets | [passes][tiles][256] look-back words
    size_t hist_words;
    unsigned long long* scan_status; // look-back words of the instance-offset scan, one per 2048 Gaussians (+ its ticket)
    size_t scan_status_words;
    uint32_t* chunk_owner; // [chunks + 1]: depth-order index of the Gaussian that owns instance 2048 * c
    size_t zero_bytes;     // everything from `hist` on is zeroed by one memset before the binning kernels
};

struct ImgState
{
    uint32_t* n_contrib;  // [H*W]
    uint2* ranges;        // [T]
    uint32_t* tile_order; // [T] tile ids by descending list length: CTA i of the compositing kernels takes tile tile_order[i]
};

constexpr int RADIX_ITEMS = 2048;  // keys per radix CTA (256 threads x 8)
constexpr int SCAN_ITEMS = 2048;   // items per scan CTA (256 threads x 8)

size_t radix_lookback_ws_words(uint32_t n_cap, int nbits); // gsr_scan_sort.cu

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
inline void carve(char*& p, T*& out, size_t count)
{
    p = (char*)align_up((size_t)p, 256);
    out = (T*)p;
    p += count * sizeof(T);
}

inline size_t geom_layout(char* base, int P, GeomState& g, int num_class = 2)
{
    char* p = base;
    g.nblk = (uint32_t)((P + PRE_BLOCK - 1) / PRE_BLOCK);
    g.slots = g.nblk * PRE_BLOCK;
    carve(p, g.counters, (size_t)CNT_WORDS);
    carve(p, g.rec, (size_t)3 * g.slots);
    carve(p, g.rect, (size_t)g.slots);
    carve(p, g.slot_gid, (size_t)g.slots);
    carve(p, g.clamped, (size_t)g.slots);
    carve(p, g.blk_count, (size_t)g.nblk);
    carve(p, g.blk_offset, (size_t)g.nblk + 1);
    for (int i = 0; i < 2; i++) {
        carve(p, g.dkeys[i], (size_t)g.slots);
        carve(p, g.dvals[i], (size_t)g.slots);
    }
    // depth sort (32 bits = 4 passes, one kernel each): [4][256] digit totals | 32 tickets | [4][tiles][256] look-back words
    g.dhist_words = radix_lookback_ws_words(g.slots, 32);
    carve(p, g.dhist, g.dhist_words);
    g.extra_pairs = num_class > 2 ? (uint32_t)((num_class + 1) / 2 - 1) : 0u; // carved LAST: the layout before it does not depend on num_class
    carve(p, g.seg_extra, (size_t)g.extra_pairs * g.slots);
    return (size_t)(p - base) + 256;
}

inline size_t bin_layout(char* base, size_t V, size_t R, int tile_bits, BinState& b)
{
    char* p = base;
    carve(p, b.point_list, R);
    carve(p, b.vals_alt, R);
    carve(p, b.tkeys[0], R);
    carve(p, b.tkeys[1], R);
    carve(p, b.soff, V + 1);
    const size_t nb = (R + RADIX_ITEMS - 1) / RADIX_ITEMS;
    b.hist_words = radix_lookback_ws_words((uint32_t)R, tile_bits);
    carve(p, b.hist, b.hist_words);
    char* zero_from = (char*)b.hist;
    b.scan_status_words = (V + SCAN_ITEMS - 1) / SCAN_ITEMS + 8;
    carve(p, b.scan_status, b.scan_status_words);
    b.zero_bytes = (size_t)(p - zero_from);
    carve(p, b.chunk_owner, nb + 2);
    return (size_t)(p - base) + 256;
}

inline size_t img_layout(char* base, size_t N, size_t T, ImgState& s)
{
    char* p = base;
    carve(p, s.n_contrib, N);
    carve(p, s.ranges, T);
    carve(p, s.tile_order, T);
    return (size_t)(p - base) + 256;
}

// ---- error plumbing (thread-local message, C-ABI return codes) ----
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define GSR_CUDA(call)                                      \
    do {                                                    \
        int _rc = gsr::check_cuda((call), #call);           \
        if (_rc) return _rc;                                \
    } while (0)

// After a kernel launch: always catch launch-configuration errors; in debug mode also synchronise
// (the reference's CHECK_CUDA, auxiliary.h:166-173).
int after_launch(cudaStream_t s, bool debug, const char* stage);
void count_launches(int n); // kernels launched by this library (gsr_launch_count)
#define GSR_LAUNCHED(stream, debug, stage)                           \
    do {                                                             \
        int _rc = gsr::after_launch((stream), (debug), (stage));     \
        if (_rc) return _rc;                                         \
    } while (0)

// ---- launch-side argument blocks ----
struct PartPtrs // one sub-scene of a fused scene (GsrGaussians.parts)
{
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    const float* segments;
};

struct PreFwdArgs
{
    int P, D, M, S;
    const float* means3D;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    const float* segments;
    const float* shs_rest; // raw-parameter mode: shs = [P,1,3], shs_rest = [P,M-1,3]
    int raw;               // activations inside the kernel (GsrGaussians.raw_params)
    const int32_t* subset; // index-list rendering: position idx < P reads input row subset[idx] (P = the list length)
    const float* view;
    const float* proj;
    const float* campos;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int grid_x, grid_y;
    int prefiltered;
    int32_t* radii;
    GeomState g;
    int num_parts;                      // > 0: Gaussian idx lives in part k with part_start[k] <= idx < part_start[k + 1]
    int part_start[GSR_MAX_PARTS + 1];
    PartPtrs part[GSR_MAX_PARTS];
};

struct PreBwdArgs
{
    int P, D, M, S;
    const float* means3D;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* shs;
    const float* shs_rest;  // raw-parameter mode (see PreFwdArgs)
    const float* opacities; // raw-parameter mode only (logits)
    const float* segments;  // raw-parameter mode only (logits)
    int raw;
    const float* cov3D_precomp;
    const float* view;
    const float* proj;
    const float* campos;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    const int32_t* radii;
    GeomState g;
    const float* grad_rec; // [slots][12]
    const float2* grad_seg_extra; // num_class > 2: [extra_pairs][slots] dL/dsegment of channel pairs 1.. (compositing backward, extra passes)
    GsrParamGrads out;
    bool colors_precomp_given;
    // packet mode (multi-GPU gradient exchange): one GSR_PACKET_WORDS record per visible Gaussian instead of dense rows
    uint32_t* packets;
    uint32_t packet_capacity;
    uint32_t* packet_count;
    uint32_t* vis_index;
    int fill; // set by launch_preprocess_bwd: dense mode, the kernel zero-fills the rows its CTA owns
    int has_subset; // index-list rendering: a CTA's slots are not the gradient rows [256 b, 256 b + 256)
};

struct GatherPacketsArgs
{
    int P, D, M, S, num_views;
    const float* means3D;
    const float* campos;
    const uint32_t* views[GSR_MAX_GATHER_VIEWS]; // one blob per view; local or PEER memory (NVLink loads)
    size_t packet_off, index_off;                 // word offsets of the packets / of the visibility index inside a blob
    uint32_t capacity;
    GsrParamGrads out;
};
int launch_gather_packets(const GatherPacketsArgs& a, cudaStream_t s);

struct RenderArgs
{
    int W, H, grid_x, grid_y;
    const uint2* ranges;
    const uint32_t* tile_order; // longest lists first (launch_tile_ranges)
    const uint32_t* point_list;
    const float4* rec;
    const float* bg;
    // forward outputs
    float* out_color;
    float* out_segment;
    float* out_depth;
    float* out_alpha;
    uint32_t* n_contrib;
    // backward inputs
    const float* alphas;
    const float* dL_dcolor;
    const float* dL_dsegment;
    const float* dL_ddepth;
    const float* dL_dalpha;
    float* grad_rec;
    // extra segment-pair passes (num_class > 2): the pair's per-slot values, how many of its two channels exist, and where the
    // backward accumulates the pair's dL/dsegment (instead of columns 4-5 of grad_rec); out_* / dL_* other than segment are null
    const float2* seg_src;
    int seg_count;
    float* grad_seg;
};

// ---- stage launchers (each defined next to its kernels) ----
int launch_preprocess_fwd(const PreFwdArgs& a, cudaStream_t s);
int launch_preprocess_bwd(const PreBwdArgs& a, cudaStream_t s);
int launch_grad_fills(const PreBwdArgs& a, cudaStream_t s);
int launch_zero_grad_rec(const GeomState& g, float* grad_rec, cudaStream_t s);
int launch_block_offsets(const GeomState& g, cudaStream_t s);
// depth (key, slot) pairs of the visible Gaussians + the digit totals of the four depth-sort passes (hist: [4][256], zeroed)
int launch_depth_keys(const GeomState& g, uint32_t* hist, cudaStream_t s);
// tiles_touched in depth order -> exclusive scan (soff[0..V]) + owner of every 2048th instance, one kernel (b.scan_status zeroed)
int launch_instance_offsets(const GeomState& g, const BinState& b, uint32_t V, uint32_t R, const uint32_t* sorted_slots, cudaStream_t s);
// (tile id, slot) per instance in depth order + digit totals of the tile sort's passes (hist: [passes][256], zeroed)
int launch_emit(const GeomState& g, const BinState& b, uint32_t V, uint32_t R, const uint32_t* sorted_slots, int grid_x, int digit_bits, int passes,
                uint32_t* hist, uint32_t* out_keys, uint32_t* out_vals, cudaStream_t s);
int launch_tile_ranges(const uint32_t* sorted_tile_keys, uint32_t R, uint2* ranges, uint32_t* tile_order, uint32_t T, cudaStream_t s);
int launch_render_fwd(const RenderArgs& a, int S, cudaStream_t s);
int launch_render_bwd(const RenderArgs& a, int S, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s);

// exclusive scan of n u32 values (in may alias out); writes the grand total to out[n] when write_total.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, bool write_total, uint32_t* partials, cudaStream_t s);
// stable LSD radix sort of (key,val) pairs on key bits [0,nbits). Result lands in keys[res]/vals[res]; returns res (0/1) or <0.
// With n_dev != nullptr, n is a CAPACITY (it sizes the grid and the histogram) and the element count is read on the device.
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int nbits, uint32_t* hist, size_t hist_words, cudaStream_t s,
                     const uint32_t* n_dev = nullptr);
int radix_num_passes(int nbits);
int radix_digit_bits(int nbits);
size_t radix_lookback_ws_words(uint32_t n_cap, int nbits);
// one kernel per pass (decoupled look-back); see gsr_scan_sort.cu
int radix_sort_pairs_lookback(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int nbits, uint32_t* ws, size_t ws_words, bool hist_ready,
                              cudaStream_t s, const uint32_t* n_dev = nullptr, int keys_per_thread = 16);

int knn_run(int P, const float* points, float* out, void* ws, size_t ws_bytes, cudaStream_t s);
size_t knn_workspace_bytes(int P);

} // namespace gsr
