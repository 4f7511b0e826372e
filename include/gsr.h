/*
 * libgsr -- B200-native (sm_100a) differentiable Gaussian-splatting rasterizer + simple-knn.
 *
 * C ABI of the drop-in boundary. Every entry point replaces one native entry point of the reference
 * (paths relative to /root/reference/submodules_local/):
 *
 *   gsr_forward        <- RasterizeGaussiansCUDA            diff-gaussian-rasterization/rasterize_points.cu:35-125
 *                         CudaRasterizer::Rasterizer::forward   cuda_rasterizer/rasterizer_impl.cu:198-344
 *   gsr_backward       <- RasterizeGaussiansBackwardCUDA    rasterize_points.cu:127-221
 *                         Rasterizer::backward                  cuda_rasterizer/rasterizer_impl.cu:348-458
 *   gsr_mark_visible   <- markVisible                       rasterize_points.cu:223-242 (rasterizer_impl.cu:141-153)
 *   gsr_knn_dist2      <- distCUDA2 / SimpleKNN::knn        simple-knn/spatial.cu:15-26, simple_knn.cu:185-221
 *   gsr_alloc_fn       <- std::function<char*(size_t)> resizeFunctional   rasterize_points.cu:27-33
 *
 * Plain pointers and sizes only: no torch types. All data pointers are DEVICE pointers on the device that is
 * current on the calling thread; the structs themselves live in host memory. All work is enqueued on `stream`
 * (a cudaStream_t). gsr_forward synchronises `stream` once (to learn the number of tile instances, like the
 * reference's blocking 4-byte read at rasterizer_impl.cu:285); nothing else blocks unless view->debug != 0.
 * Functions return 0 on success or a negative GSR_ERR_* code; gsr_last_error() describes the last failure on
 * the calling thread. There is no CPU fallback anywhere in this library.
 */
#ifndef GSR_H_INCLUDED
#define GSR_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 4
#define GSR_MAX_NUM_CLASS 64

#define GSR_OK 0
#define GSR_ERR_INVALID_ARGUMENT (-1)
#define GSR_ERR_CUDA (-2)
#define GSR_ERR_ALLOC (-3)
#define GSR_ERR_PREFILTERED (-4) /* prefiltered=1 but a Gaussian was culled (reference: printf + __trap) */
#define GSR_ERR_UNSUPPORTED (-5)
#define GSR_ERR_OVERFLOW (-6) /* more than 2^31-1 tile instances */

typedef void* gsr_stream_t; /* cudaStream_t */

/* State buffers, same three as the reference (geomBuffer, binningBuffer, imgBuffer). Contents are opaque and
 * differ from the reference's layout; they must be handed back unchanged to gsr_backward. */
enum { GSR_BUF_GEOM = 0, GSR_BUF_BINNING = 1, GSR_BUF_IMG = 2 };

/* Called by gsr_forward to obtain each state buffer once its size is known. Must return a device pointer
 * aligned to >= 256 bytes (NULL = failure). bytes may be 0. GSR_BUF_BINNING may be requested twice in one forward: first
 * speculatively (the previous forward's size plus a margin, before the instance count has come back from the device, so that the
 * allocator runs while the GPU is busy) and again with the exact size if that was too small; the LAST answer for a buffer kind is
 * the buffer of this forward, an earlier one is unused and may be released. */
typedef void* (*gsr_alloc_fn)(void* user, int which, size_t bytes);

/* Per-view constants: GaussianRasterizationSettings (diff_gaussian_rasterization/__init__.py:168-180). */
typedef struct GsrView {
    int32_t image_width, image_height;
    float tanfovx, tanfovy;
    float scale_modifier;
    int32_t sh_degree;   /* D: active degree 0..3 */
    int32_t sh_coeffs;   /* M: coefficients per Gaussian in `shs` (0 when shs == NULL) */
    int32_t num_class;   /* channels of `segments` / out segment, 0 .. GSR_MAX_NUM_CLASS at run time (the reference compiles NUM_CLASS = 2 in,
                            config.h:16, while its ModelParams default to 29): channels 0-1 ride in the splat record, every further pair is
                            composited by one more pass over the same sorted lists, forward and backward */
    int32_t prefiltered;
    int32_t debug;       /* != 0: synchronise and check after every stage (auxiliary.h:166-173) */
    const float* bg;         /* [3] */
    const float* viewmatrix; /* [16], as the reference receives it (world_view_transform, column-major for the kernels) */
    const float* projmatrix; /* [16] full_proj_transform */
    const float* campos;     /* [3] */
} GsrView;

/* Gaussian inputs, row-major contiguous fp32. NULL = "not provided" (the reference's empty-tensor sentinel). */
typedef struct GsrGaussians {
    int32_t P;
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* segments;       /* [P,num_class] or NULL (then the segment output is zero) */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] (w,x,y,z; used un-normalised) or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    /* Fused-activation entry (replaces the five torch kernels + the 1.15 GB torch.cat of
     * scene/gaussian_model.py:100-124 in front of the rasterizer). raw_params != 0: the tensors above are the model's RAW
     * parameters and the activations run inside the preprocess kernels, forward and backward:
     *   opacities, segments = logits (sigmoid), scales = log-scales (exp), rotations = un-normalised quaternions
     *   (x / max(|x|, 1e-12), torch.nn.functional.normalize), shs = _features_dc [P,1,3] and shs_rest = _features_rest
     *   [P,M-1,3] (the torch.cat is never materialised). Gradients are then w.r.t. the raw parameters, and gsr_backward
     *   needs `opacities` (raw) as well. cov3D_precomp must be NULL. raw_params == 0: classic behaviour, shs_rest ignored. */
    const float* shs_rest;
    int32_t raw_params;
    /* Index-list rendering: subset != NULL renders only the Gaussians subset[0..subset_count) -- row numbers
     * into the P-row tensors above, STRICTLY ASCENDING -- without materialising masked copies of every tensor (the viewer's
     * bbox mask, gaussian_renderer/__init__.py:239-268, and sub-scene selection). Per-Gaussian outputs (radii) then have
     * subset_count entries, in list order; gradients stay P rows (rows outside the list are zero). State buffers are sized by
     * subset_count. Not combined with gsr_backward_packets. */
    const int32_t* subset;
    int32_t subset_count;
    /* Sub-scene fusion without concatenation. The reference's viewer merges sub-scenes by concatenating every
     * attribute array (visualizer.py:196-226, _merge_scenes: np.concatenate of xyz, features, opacities, segments, rots, scales).
     * num_parts > 0 renders parts[0], parts[1], ... as ONE scene in that order -- Gaussian i of the fused scene is row
     * i - start(k) of part k -- straight from the parts' own tensors: P must equal the sum of parts[k].P, the top-level data
     * pointers are ignored, every part provides the same set of members, and the result is bit-identical to rendering the
     * concatenated tensors. Render-only (gsr_backward rejects it); not combined with raw_params or subset. */
    const struct GsrGaussians* parts;
    int32_t num_parts;
} GsrGaussians;
#define GSR_MAX_PARTS 16

/* Forward outputs; every element is written by the kernels (no pre-zeroing needed) when P > 0. */
typedef struct GsrOutputs {
    float* color;   /* [3,H,W] */
    float* segment; /* [num_class,H,W] (may be NULL when num_class == 0) */
    float* depth;   /* [1,H,W] */
    float* alpha;   /* [1,H,W] = sum_i alpha_i T_i */
    int32_t* radii; /* [P] */
} GsrOutputs;

typedef struct GsrState {
    void* geom;
    void* binning;
    void* img;
    int32_t num_rendered; /* R, as returned by gsr_forward */
    int32_t num_visible;  /* V of THAT forward (gsr_last_num_visible() right after it), or 0 when unknown: lets
                             gsr_backward_packets refuse a packet buffer that is too small instead of dropping packets */
} GsrState;

typedef struct GsrPixelGrads {
    const float* dL_dcolor;   /* [3,H,W] */
    const float* dL_dsegment; /* [num_class,H,W] or NULL (= zeros) */
    const float* dL_ddepth;   /* [1,H,W] or NULL (= zeros) */
    const float* dL_dalpha;   /* [1,H,W] or NULL (= zeros) */
} GsrPixelGrads;

/* Dense gradients. accumulate == 0: every row is written (zeros for invisible Gaussians). accumulate != 0: the rows of
 * visible Gaussians are ADDED to what the buffers hold and nothing else is touched -- the multi-view path, where several
 * views of one step sum into one flat gradient buffer without a dense zero-fill or a torch add per view.
 * NULL members are skipped. */
typedef struct GsrParamGrads {
    float* dL_dmeans3D;   /* [P,3] */
    float* dL_dmeans2D;   /* [P,3] (third column 0) */
    float* dL_dsh;        /* [P,M,3] (needs shs) */
    float* dL_dcolors;    /* [P,3]  gradient w.r.t. colors_precomp (or the internal SH colour) */
    float* dL_dsegments;  /* [P,num_class] */
    float* dL_dopacity;   /* [P] */
    float* dL_dscales;    /* [P,3] (needs scales) */
    float* dL_drotations; /* [P,4] (needs rotations) */
    float* dL_dcov3D;     /* [P,6] */
    int32_t accumulate;
    float* dL_dsh_rest;   /* raw_params only: dL_dsh is then [P,1,3] (features_dc) and this is [P,M-1,3] (features_rest) */
} GsrParamGrads;

int gsr_abi_version(void);
const char* gsr_last_error(void);

/* Returns R (>= 0) in *num_rendered. `out->radii` etc. must be valid for P > 0. P == 0 is a no-op returning R = 0
 * (the host side returns zero images, as rasterize_points.cu:87 does). */
int gsr_forward(const GsrView* view, const GsrGaussians* in, const GsrOutputs* out, gsr_alloc_fn alloc, void* alloc_user,
                int32_t* num_rendered, gsr_stream_t stream);

/* Bytes of scratch gsr_backward needs (per-Gaussian gradient records accumulated by the compositing backward); the _n form for
 * num_class > 2. P = rendered Gaussians (subset_count with an index list). */
size_t gsr_backward_scratch_bytes(int32_t P);
size_t gsr_backward_scratch_bytes_n(int32_t P, int32_t num_class);

int gsr_backward(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                 const GsrPixelGrads* pix, const GsrParamGrads* grads, void* scratch, size_t scratch_bytes, gsr_stream_t stream);

/* ---- multi-GPU gradient exchange (multi-view data parallelism) ----
 * Instead of dense rows, the backward can emit one compact PACKET per visible Gaussian of the view, 16 floats = 64 bytes,
 * 64-byte aligned (one 16-byte-vector copy per quarter):
 *   words 0-2   dL/dRGB of the SH colour, clamp mask applied      words 3-5   dL/dmean3D (all paths, incl. the SH view direction)
 *   word 6      dL/dopacity     words 7-8  dL/dsegment     words 9-11  dL/dscale     words 12-15  dL/drotation
 * instead of the 244-B dense row: the 48-float SH gradient row is rank one, basis(view direction) x dL/dRGB, and is rebuilt by
 * gsr_gather_packets on the receiving rank from the Gaussian's position and that view's camera centre. Packets carry no id:
 * they are ordered by Gaussian id and the view's visibility index (below) maps a Gaussian to its packet.
 * *count_dev receives the number of visible Gaussians. `capacity` must be >= the number of
 * visible Gaussians of the forward the state belongs to: with state->num_visible set, a smaller buffer fails with
 * GSR_ERR_OVERFLOW before anything is launched (with num_visible == 0 = unknown, packets beyond `capacity` are dropped and
 * *count_dev tells). dL_dmeans2D (optional, dense [P,3]) is overwritten for the
 * densification statistics. Requires shs + scales/rotations (the training configuration). */
#define GSR_PACKET_WORDS 16
int gsr_backward_packets(const GsrView* view, const GsrGaussians* in, const int32_t* radii, const GsrState* state, const float* alpha,
                         const GsrPixelGrads* pix, uint32_t* packets, uint32_t capacity, uint32_t* count_dev, float* dL_dmeans2D,
                         uint32_t* vis_index, void* scratch, size_t scratch_bytes, gsr_stream_t stream);
/* vis_index (optional, [W][2] words, W = ceil(P / 32); gsr_packet_index_words(P) = 2 * W rounded up to a multiple of 32 words is
 * the room to reserve for it; must be 8-byte aligned): the view's
 * visibility index, written in full. Pair w = { bits, ~first }: bit b of `bits` is set when Gaussian 32 * w + b is visible;
 * `first` is the packet index of the group's first visible Gaussian (packets are in ascending Gaussian order), so the packet
 * of Gaussian i is  first[i / 32] + popcount(bits[i / 32] & ((1 << (i % 32)) - 1)). A group without visible Gaussians is {0, 0}. */
size_t gsr_packet_index_words(int32_t P);
/* One pass over all Gaussians that SUMS the packets of `num_views` views (all-gathered from all ranks) into dense gradient
 * rows and WRITES every row (zeros where no view saw the Gaussian): no zero fill, no read-modify-write per view.
 * blobs: num_views view blobs, `blob_stride_words` apart; one blob = the view's vis_index (gsr_packet_index_words(P) words)
 * followed by [capacity][16] packet words -- exactly one all-gather payload per view. campos: [num_views][3] (device). Views are
 * summed in index order on every rank, so replicas end up bitwise identical. dL_dmeans2D / dL_dcolors / dL_dcov3D are ignored. */
int gsr_gather_packets(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t num_class, const float* means3D, int32_t num_views,
                       const float* campos, const uint32_t* blobs, size_t blob_stride_words, uint32_t capacity,
                       const GsrParamGrads* grads, gsr_stream_t stream);
/* The same pass over blobs that need not be adjacent and need not be local: view_ptrs[v] (host array of num_views DEVICE
 * pointers) is the blob of view v, in this GPU's memory or in a PEER GPU's (a pointer from gsr_peer_open). With peer pointers
 * the gather pulls every rank's packets straight over NVLink: no all-gather, no staging copy. Inside a blob the packets start
 * at word `packet_off_words` and the visibility index at word `index_off_words`. The caller orders the kernel after the
 * producers (a stream-ordered barrier across ranks) and keeps the blobs unchanged until every reader is done. */
#define GSR_MAX_GATHER_VIEWS 64
int gsr_gather_packets_v(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t num_class, const float* means3D, int32_t num_views,
                         const float* campos, const uint32_t* const* view_ptrs, size_t packet_off_words, size_t index_off_words,
                         uint32_t capacity, const GsrParamGrads* grads, gsr_stream_t stream);

/* Peer-visible device buffers for gsr_gather_packets_v (one process per GPU, same node). gsr_peer_alloc: cudaMalloc on the
 * current device + its 64-byte inter-process handle (GSR_PEER_HANDLE_BYTES), which the host ships to the other ranks by any
 * means (torch.distributed all-gather); gsr_peer_open maps a peer's handle into this process (peer access over NVLink);
 * gsr_peer_close unmaps it; gsr_peer_free releases the owner's allocation. */
#define GSR_PEER_HANDLE_BYTES 64
int gsr_peer_alloc(size_t bytes, void** ptr, void* handle_out);
int gsr_peer_open(const void* handle, void** ptr);
int gsr_peer_close(void* ptr);
int gsr_peer_free(void* ptr);
/* Stream-ordered copy of `bytes` between two device buffers of which either may be a peer's (gsr_peer_open): the PUSH form of the
 * exchange -- a rank writes its view blob into every peer's receive slot with the copy engines (posted NVLink writes) and the
 * peers' gather kernels then read local memory only. */
int gsr_peer_copy(void* dst, const void* src, size_t bytes, gsr_stream_t stream);

/* Fused multi-tensor Adam over flat buffers (replaces torch.optim.Adam(l, lr=0.0, eps=1e-15) and its seven
 * parameter groups, scene/gaussian_model.py:166-177, following the default foreach path op by op in fp32). params / grads /
 * exp_avg / exp_avg_sq are flat fp32 buffers of the same layout; group g covers elements [offset, offset + count) with
 * learning rate lr. step = the 1-based step number (bias corrections are evaluated in double on the host, as Python does).
 * One launch; 28 bytes of HBM traffic per parameter. No weight decay, no amsgrad (the reference uses neither). */
typedef struct GsrAdamGroup {
    uint64_t offset;
    uint64_t count;
    float lr;
    int32_t step; /* > 0: this group's own 1-based step number (torch.optim.Adam counts steps per parameter); 0: use `step` */
} GsrAdamGroup;
int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const GsrAdamGroup* groups, int32_t num_groups,
                  double beta1, double beta2, double eps, int32_t step, gsr_stream_t stream);

/* Fused photometric loss of the training step (train.py:110-111 with utils/loss_utils.py:104-150):
 *   loss = (1 - lambda_dssim) * mean|image - gt| + lambda_dssim * (1 - mean(SSIM_11x11,sigma=1.5(image, gt)))
 * loss_out[3] = {L1, SSIM, loss} (device); dL_dimage [C,H,W] = grad_scale * dloss/dimage (NULL: forward only). Two launches,
 * deterministic (no float atomics). image / gt are [C,H,W] fp32; zero padding like F.conv2d(padding=5). */
size_t gsr_image_loss_scratch_bytes(int32_t C, int32_t H, int32_t W);
int gsr_image_loss(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim, float grad_scale,
                   float* loss_out, float* dL_dimage, void* scratch, size_t scratch_bytes, gsr_stream_t stream);

/* Depth supervision of the training step (train.py:118-121 with depth_loss_choice 'localrf',
 * utils/loss_utils.py:88-102 compute_depth_loss, and the normalisation of gaussian_renderer/__init__.py:375):
 *   mode 0: in = compute_depth_loss's dyn_depth x;   mode 1: in = the rasterizer's raw depth d, and x = 1 / max(d / (max(d) + 1e-5), 1e-6)
 *   t = median(x) (lower middle element), s = mean|x - t|, same for gt;  a = ((x - t)/s - (gt - tg)/sg)^2;
 *   loss = lambda * mean(a where a <= quantile(a, 0.8) else 0)            (quantile with torch's float32 rank and lerp)
 * loss_out[1] (device). grad_out [n] = grad_scale * dloss/d(in) (NULL: value only), including the paths through the median element,
 * the mean absolute deviation and (mode 1) the arg-max pixel. Order statistics by radix select, no sort, no host sync; n < 2^31. */
size_t gsr_depth_loss_scratch_bytes(int64_t n);
int gsr_depth_loss(const float* in, const float* gt, int64_t n, float lambda, float grad_scale, int32_t mode, float* loss_out, float* grad_out,
                   void* scratch, size_t scratch_bytes, gsr_stream_t stream);

/* Densify / prune as one index list (replaces the per-tensor boolean masks and torch.cat of
 * scene/gaussian_model.py:377-441 over seven parameters and their two Adam moments). src and dst are flat buffers of
 * num_blocks blocks; block k starts at float src_offsets[k] (resp. dst_offsets[k]) and holds n_src (resp. n_out) rows of
 * row_floats[k] floats. Row j of every dst block = row index[j] of the matching src block; index[j] == -1 gives a row of
 * zeros (the moments of a freshly cloned / split Gaussian). One launch. */
int gsr_select_rows(const float* src, float* dst, const int64_t* index, int64_t n_out, int64_t n_src, const int32_t* row_floats,
                    const uint64_t* src_offsets, const uint64_t* dst_offsets, int32_t num_blocks, gsr_stream_t stream);

/* Number of visible Gaussians of the most recent gsr_forward on the calling thread. */
uint32_t gsr_last_num_visible(void);

int gsr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     gsr_stream_t stream);

size_t gsr_knn_workspace_bytes(int32_t P);
int gsr_knn_dist2(int32_t P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes, gsr_stream_t stream);

/* Test/diagnostic support: unpack the opaque state of a forward pass into reference-style, Gaussian-id-indexed
 * arrays so it can be compared bit for bit with the reference's GeometryState/BinningState/ImageState
 * (rasterizer_impl.cu:155-194). Every member may be NULL. */
typedef struct GsrStateExport {
    float* depths;           /* [P]   geomState.depths (0 for invisible) */
    float* means2D;          /* [P,2] */
    float* conic_opacity;    /* [P,4] */
    float* rgb;              /* [P,3] */
    uint8_t* clamped;        /* [P,3] */
    uint32_t* tiles_touched; /* [P] */
    uint64_t* point_keys;    /* [R] sorted keys  (tile << 32) | depth bits */
    uint32_t* point_list;    /* [R] sorted Gaussian ids */
    uint32_t* ranges;        /* [T,2] */
    uint32_t* n_contrib;     /* [H*W] */
} GsrStateExport;

int gsr_export_state(int32_t P, int32_t image_width, int32_t image_height, const GsrState* state, const GsrStateExport* out,
                     gsr_stream_t stream);

/* Per-stage device timings (ms) of the most recent gsr_forward / gsr_backward on this thread when
 * gsr_set_profiling(1) is active (adds cudaEvent records; used by bench.py for the roofline block). */
/* Number of CUDA kernels this library has launched so far in this process (all threads). */
unsigned long long gsr_launch_count(void);

/* ---- measurement support (bench.py, tests; not on the product path) ----
 * gsr_microbench: achievable rates of the instruction classes that bound the compositing kernels, measured on the current device
 * with CUDA events on `stream` (blocks until done, ~20 ms): dependent-chain-free FFMA (3-register form), packed FFMA2
 * (fma.rn.f32x2), MUFU.EX2, SHFL, and the backward's atomic pattern (12 consecutive floats of one 48-byte record per warp
 * instruction, records spread over an L2-resident table). SURVEY.md 8d item 2. */
typedef struct GsrMicrobench {
    float ffma_tflops;   /* 2 flop per lane-op */
    float ffma2_tflops;  /* 4 flop per lane-op */
    float ex2_gops;      /* 1e9 lane-ops / s */
    float shfl_gops;     /* 1e9 lane-ops / s */
    float red_gops;      /* 1e9 float atomics / s */
    float sm_clock_mhz_nominal;
    int32_t sm_count;
} GsrMicrobench;
int gsr_microbench(GsrMicrobench* result, gsr_stream_t stream);
/* gsr_count_work: algorithmic work of a rendered frame, counted by replaying the reference's per-pixel loop (forward.cu:314-377)
 * over the saved state: counters_dev[0] = E (list entries evaluated until each pixel is done), [1] = Cc (entries that
 * contributed), [2] = E_b (entries the backward re-traverses = sum of n_contrib). P = rendered Gaussians (subset_count with an
 * index list). */
int gsr_count_work(int32_t P, int32_t image_width, int32_t image_height, const GsrState* state, uint64_t* counters_dev, gsr_stream_t stream);

#define GSR_STAGE_COUNT 16
void gsr_set_profiling(int enable);
int gsr_get_stage_times(float* ms /*[GSR_STAGE_COUNT]*/, const char** names /*[GSR_STAGE_COUNT]*/);

#ifdef __cplusplus
}
#endif
#endif
