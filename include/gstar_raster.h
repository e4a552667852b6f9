/*
 * gstar_raster.h -- C ABI of the B200-native differentiable surface-Gaussian rasterizer.
 *
 * This is the drop-in boundary for the ONE hot path of eth-ait/GauSTAR: the operator behind
 * diff_gaussian_rasterization.GaussianRasterizer.  Each entry point replaces one static method of
 * CudaRasterizer::Rasterizer in the reference
 * (DGR = gaussian_splatting/submodules/diff-gaussian-rasterization):
 *
 *   gstar_raster_forward   <->  Rasterizer::forward    DGR/cuda_rasterizer/rasterizer.h:31-56
 *                                                      (impl. rasterizer_impl.cu:198-336)
 *   gstar_raster_backward  <->  Rasterizer::backward   DGR/cuda_rasterizer/rasterizer.h:58-84
 *                                                      (impl. rasterizer_impl.cu:340-434)
 *   gstar_mark_visible     <->  Rasterizer::markVisible DGR/cuda_rasterizer/rasterizer.h:24-29
 *                                                      (impl. rasterizer_impl.cu:141-153)
 *
 * Beyond the reference (no counterpart there): gstar_raster_reblend -- a second feature pass over the same
 * Gaussians and camera that reuses the first pass's sorted record stream -- with gstar_bwd_args.blend_only for running
 * the per-Gaussian backward stage once for all passes of a view; and CUDA-graph capture of forward / backward / re-blend.
 *
 * Plain pointers and sizes only: no torch types, no C++ types, no exceptions cross this boundary.
 * All data pointers are DEVICE pointers on the current CUDA device unless stated otherwise; all
 * work is enqueued on `stream`.  The reference binding that a maintainer would write against this
 * header is shown in INTEGRATION.md; the torch binding shipped here is gaustar_b200/csrc/torch_ext.cpp.
 *
 * Return convention: >= 0 success, < 0 error (message from gstar_last_error(), thread local).
 */
#ifndef GSTAR_RASTER_H
#define GSTAR_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSTAR_ABI_VERSION 5

#if defined(__GNUC__)
#define GSTAR_API __attribute__((visibility("default")))
#else
#define GSTAR_API
#endif

/* error codes */
#define GSTAR_ERR_INVALID  (-1) /* bad argument combination */
#define GSTAR_ERR_CUDA     (-2) /* a CUDA call failed (debug mode: after a device synchronize) */
#define GSTAR_ERR_ALLOC    (-3) /* a buffer callback returned NULL */
#define GSTAR_ERR_NONRGB   (-4) /* "For non-RGB, provide precomputed Gaussian colors!" rasterizer_impl.cu:242-245 */
#define GSTAR_ERR_NOLOG    (-5) /* a two-pass forward (colors2) that must serve a backward cannot get its hit log: use gstar_raster_reblend */

/* Resizable-buffer callback: mirrors the three std::function<char*(size_t)> of rasterizer.h:32-34
 * (created by resizeFunctional(), DGR/rasterize_points.cu:27-33).  Must return a device pointer to
 * at least nbytes bytes, 128-byte aligned, owned by the caller.  It may be called more than once per
 * forward for the binning buffer (a second time, with a larger size, only if the instance count
 * exceeded the first provision). */
typedef char* (*gstar_alloc_fn)(void* user, size_t nbytes);

/* Arguments of Rasterizer::forward, rasterizer.h:35-56, in the same order and meaning.
 * Absent optional inputs are NULL (the reference passes data_ptr() of an empty tensor). */
typedef struct gstar_fwd_args {
    int P, D, M;                 /* gaussians, active SH degree, SH coeffs per gaussian */
    const float* background;     /* [3] */
    int width, height;
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3]   or NULL */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3]   or NULL */
    float scale_modifier;
    const float* rotations;      /* [P,4]   or NULL */
    const float* cov3D_precomp;  /* [P,6]   or NULL */
    const float* viewmatrix;     /* [16] row-vector convention, auxiliary.h:58-77 */
    const float* projmatrix;     /* [16] */
    const float* cam_pos;        /* [3] */
    float tan_fovx, tan_fovy;
    int prefiltered;
    float* out_color;            /* [3,H,W] fully written */
    int* radii;                  /* [P] fully written (0 = culled) */
    int debug;                   /* !=0: synchronize + check after every stage (auxiliary.h:166-173) */
    int forward_only;            /* !=0: no backward will follow (inference): the forward skips the hit log; a backward on
                                  * these buffers is still correct (it takes the walk-back kernel) */
    /* Two feature passes in ONE blend (no counterpart in the reference, whose NUM_CHANNELS is a compile-time 3: config.h:15;
     * GauSTAR renders RGB and then depth from the same Gaussians and camera, refine.py:552-564 / :607-616).  With colors2 != NULL
     * the call also renders what a second call with colors_precomp = colors2 and background = background2 would: the passes
     * share every pair's alpha and transmittance, so the second image costs three multiply-adds per blended pair instead of a
     * blend of its own -- bit-identical to that second call.  All three pointers or none.  Unless forward_only is set the call
     * must end with a hit log (gstar_set_hit_log): it waits for the sort instead of the scan, grows the log and re-blends if
     * the view needs more than was provisioned, and fails with GSTAR_ERR_NOLOG when the log is switched off or capped below
     * the need -- the caller then renders the second pass with gstar_raster_reblend.  Not available while capturing.
     * The second "pass" has channels2 = 1..4 channels (0 means 3): depth as ONE channel and a normal as three, say -- with RGB
     * that is a seven-channel blend.  colors2 is always a float4 per Gaussian (one 16-byte gather; channels >= channels2 must be
     * finite, e.g. zero); background2 and out_color2 have channels2 entries / planes. */
    const float* colors2;        /* [P,4] or NULL */
    const float* background2;    /* [channels2] */
    float* out_color2;           /* [channels2,H,W] fully written */
    int channels2;               /* 1..4; 0 = 3 */
} gstar_fwd_args;

/* Forward pass.  Returns num_rendered (sum of tiles touched), exactly like Rasterizer::forward.
 * CUDA graphs: if `stream` is being captured, nothing is waited for or read back on the host; the binning
 * buffer gets the capacity that earlier un-captured calls of this thread provisioned (at least one such call is
 * required; it also performs the one-time setup that is illegal inside a capture), the return value is that capacity
 * -- an upper bound of num_rendered, valid as `R` of the matching backward -- and a replay whose view needs more
 * instances leaves the overflow flag of its header set (gstar_debug_header) instead of growing the buffer.  debug != 0
 * is rejected while capturing.  gstar_raster_backward and gstar_raster_reblend never touch the host and capture as is. */
GSTAR_API int gstar_raster_forward(const gstar_fwd_args* args,
                         gstar_alloc_fn geom_alloc, void* geom_user,
                         gstar_alloc_fn binning_alloc, void* binning_user,
                         gstar_alloc_fn image_alloc, void* image_user,
                         void* stream /* cudaStream_t */);

/* Arguments of Rasterizer::backward, rasterizer.h:58-84. */
typedef struct gstar_bwd_args {
    int P, D, M, R;
    const float* background;
    int width, height;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy;
    const int* radii;
    char* geom_buffer;           /* the three opaque buffers of the matching forward call */
    char* binning_buffer;
    char* image_buffer;
    const float* dL_dpix;        /* [3,H,W] */
    /* Outputs.  Unlike the reference (which accumulates with atomics into caller-zeroed arrays,
     * rasterize_points.cu:150-158) every output below is FULLY OVERWRITTEN; no zero-fill needed. */
    float* dL_dmean2D;           /* [P,3] (x,y in NDC-scaled pixels, z = 0) backward.cu:545-546 */
    float* dL_dconic;            /* [P,4] (.x,.y,.w used)                    backward.cu:549-551; may be NULL (intermediate) */
    float* dL_dopacity;          /* [P]                                      backward.cu:554 */
    float* dL_dcolor;            /* [P,3]                                    backward.cu:523; may be NULL unless colors_precomp */
    float* dL_dmean3D;           /* [P,3] */
    float* dL_dcov3D;            /* [P,6]; may be NULL unless cov3D_precomp (an intermediate otherwise) */
    float* dL_dsh;               /* [P,M,3] (NULL if M == 0) */
    float* dL_dscale;            /* [P,3] */
    float* dL_drot;              /* [P,4] */
    /* Scratch: P * GSTAR_GRAD_SCRATCH_FLOATS floats, ZERO-FILLED by the caller (or pre-loaded, see blend_only). The blend
     * backward reduces per-(gaussian,tile) partial sums into it with one vector of atomics per
     * warp instead of the reference's nine atomics per (gaussian,pixel). */
    float* blend_grad_scratch;
    int debug;
    /* Multi-view steps: if non-zero, the five PARAMETER gradients (dL_dmean3D, dL_dscale, dL_drot, dL_dsh,
     * dL_dopacity) are ACCUMULATED (+=) into the given arrays instead of overwritten -- a step's views sum
     * straight into the flat buffer that is all-reduced once per step, with no separate
     * accumulation pass.  Invisible Gaussians are then not touched at all.  The other outputs are unaffected.
     * 1: plain read-modify-write -- one backward at a time per array (calls on one stream).  2: reductions at L2
     * (red.global.add) -- backward calls running concurrently on different streams may add into the SAME arrays, so a step that
     * keeps several views in flight needs one gradient buffer, not one per stream plus a merge. */
    int accumulate_param_grads;
    /* Several feature passes over one geometry (gstar_raster_reblend): if non-zero, only the blend stage runs -- the nine
     * raw moments of this pass ([S, S dx, S dy, S dx2, S dxdy, S dy2] and the three colour moments = this pass's
     * dL_dcolor) are ADDED into blend_grad_scratch and no gradient output is written (all may be NULL).  The six
     * geometric moments of the passes of one view add, so the caller copies floats 6..8 of every row out as the pass's
     * dL_dcolor, zeroes them, and hands the same scratch -- now pre-loaded -- to the full backward of the last pass: the
     * per-Gaussian stage then runs ONCE for all passes.  (A scratch that is not zero on entry is simply added to.) */
    int blend_only;
    /* Backward of a two-pass forward (gstar_fwd_args::colors2): the second image's upstream gradient [channels2,H,W], its
     * background and its colours ([P,4], as in the forward).  dL/dalpha of a pair sums both passes; the second pass's dL_dcolors
     * (colour moments) are left in floats 9..11 of every row of blend_grad_scratch -- and, for a fourth channel, ADDED into
     * blend_grad_scratch2 [P] (caller-zeroed) -- the first pass's go where they always do.  Must match the forward.  The
     * deterministic mode covers up to three channels. */
    const float* dL_dpix2;
    const float* background2;
    const float* colors2;
    int channels2;               /* 1..4; 0 = 3 */
    float* blend_grad_scratch2;  /* [P], needed iff channels2 == 4 */
} gstar_bwd_args;
#define GSTAR_GRAD_SCRATCH_FLOATS 12

GSTAR_API int gstar_raster_backward(const gstar_bwd_args* args, void* stream);

/* ---- shared-geometry re-blend (no counterpart in the reference, which repeats the whole forward) ----
 * GauSTAR rasterizes the same Gaussians from the same camera twice per training step: RGB, then depth as three equal
 * channels through colors_precomp (gaustar_trainers/refine.py:552-564 and :607-616; refined_mesh.py:733-774 does it three
 * times).  Preprocess, binning and the sort depend on neither colour nor background, so the second call can start from
 * the first call's sorted record stream: gstar_raster_reblend copies it with the colour fields replaced by
 * colors_precomp[gid] into a binning buffer of its own and runs the blend kernel.  out_color and the image buffer are
 * what gstar_raster_forward would produce for (the source call's geometry inputs, colors_precomp, background),
 * bit for bit.
 *
 * The source call is identified by its image buffer and must be one of the last 16 forward / re-blend calls made by THIS
 * host thread on this device (their layouts are remembered on the host, so that nothing is read back from the device);
 * otherwise GSTAR_ERR_INVALID.  The source buffers must still be alive and the work is enqueued on `stream`, which must
 * be ordered behind the source call.  The backward of a re-blend is gstar_raster_backward with
 *     geom_buffer    = the SOURCE call's geometry buffer (shared, read-only: positions, conics, opacities, radii),
 *     binning_buffer / image_buffer = the ones allocated here,  colors_precomp = the colours given here,  shs = NULL,
 *     R = the value returned here (the source call's num_rendered).
 * forward_only != 0: no backward will follow; the new binning buffer then holds no hit log. */
typedef struct gstar_reblend_args {
    int P;                          /* gaussians of the source call */
    int width, height;              /* as in the source call */
    const float* background;        /* [3] */
    const float* colors_precomp;    /* [P,3] the new per-gaussian colours */
    const char* src_binning_buffer; /* the source call's binning buffer */
    const char* src_image_buffer;   /* the source call's image buffer */
    float* out_color;               /* [3,H,W] fully written */
    int debug;
    int forward_only;
    /* Optional guard (all four or none; device pointers to 16 floats): the source call's camera and this call's.  They
     * are compared bit for bit ON THE DEVICE (no host round trip); if they differ the re-blend is refused there: the
     * blend kernels skip the call, out_color is filled with NaN and the new header's overflow word reads 2
     * (gstar_debug_header) -- a re-blend through another camera must not return a plausible image of the wrong view. */
    const float* src_viewmatrix;
    const float* src_projmatrix;
    const float* viewmatrix;
    const float* projmatrix;
} gstar_reblend_args;
GSTAR_API int gstar_raster_reblend(const gstar_reblend_args* args,
                         gstar_alloc_fn binning_alloc, void* binning_user,
                         gstar_alloc_fn image_alloc, void* image_user,
                         void* stream /* cudaStream_t */);

/* checkFrustum: present[i] = (view-space z > 0.2).  present is a byte array (C++ bool). */
GSTAR_API int gstar_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                       unsigned char* present, void* stream);

/* Thread-local message of the last error returned on this thread. */
GSTAR_API const char* gstar_last_error(void);
GSTAR_API int gstar_abi_version(void);

/* ---- sizes and views of the opaque buffers (the layout is private; these are for callers that
 * pre-provision memory and for the parity tests that compare intermediates) ---- */
GSTAR_API size_t gstar_geom_bytes(int P);
GSTAR_API size_t gstar_image_bytes(int width, int height);
GSTAR_API size_t gstar_binning_bytes(size_t capacity_instances);

/* Copies the per-gaussian intermediates out of a geometry buffer into reference-layout arrays
 * (any may be NULL): depths[P], means2D[P,2], conic_opacity[P,4], rgb[P,3], tiles_touched[P],
 * clamped[P,3] (bytes).  Mirrors GeometryState, rasterizer_impl.h:33-47. */
GSTAR_API int gstar_geom_unpack(const char* geom_buffer, int P, float* depths, float* means2D, float* conic_opacity, float* rgb,
                      uint32_t* tiles_touched, unsigned char* clamped, void* stream);
/* Views into the image buffer (ImageState, rasterizer_impl.h:49-56): final_T[H*W], n_contrib[H*W],
 * ranges[T,2] with T = ceil(W/16)*ceil(H/16). */
GSTAR_API int gstar_image_views(char* image_buffer, int width, int height, float** final_T, uint32_t** n_contrib, uint32_t** ranges);
/* View of the sorted instance list (BinningState::point_list, rasterizer_impl.h:58-68).  Synchronous (reads the
 * provisioned capacity from the image buffer's header); meant for tests. */
GSTAR_API int gstar_binning_views(char* binning_buffer, char* image_buffer, uint32_t** point_list, uint64_t* capacity);

/* ---- hit log (backward strategy) ----
 * The forward blend can record, for every blended (instance, pixel) pair, the transmittance in front of it and the
 * colour accumulated so far ("hit log", 16 bytes per pixel of every instance's footprint, carved out of the binning
 * buffer).  The backward blend then needs no per-pixel back-to-front walk (backward.cu:472-556 in closed form) and
 * runs instance-parallel.  The log is sized from the previous call on this thread/device; a view whose log does not
 * fit -- or any view when the log is switched off -- takes the walk-back kernel instead (decided on the device, same
 * results within fp32 rounding).  Environment: GSTAR_HIT_LOG=0 disables it, GSTAR_HIT_LOG_MAX_MB caps its size
 * (default 8192).  gstar_set_hit_log(0|1) overrides the environment (returns the previous mode; other values only
 * query).  gstar_hit_log_state() is a synchronous test helper reading one forward call's header. */
GSTAR_API int gstar_set_hit_log(int mode);
/* ---- deterministic backward (test mode) ----
 * The reference's backward is not reproducible run to run: every (Gaussian, pixel) pair adds into the Gaussian's gradients with
 * fp32 atomics in whatever order the hardware schedules them (backward.cu:523-554), and so does this library's blend backward per
 * (Gaussian, tile).  gstar_set_deterministic(1) -- or GSTAR_DETERMINISTIC=1 in the environment -- makes gstar_raster_backward
 * leave one row of moments per record in a scratch of R x 48 bytes (cudaMallocAsync on the call's stream) and add every
 * Gaussian's rows up in a fixed order: all gradients are then BIT-IDENTICAL from run to run (the forward always is), at roughly
 * twice the blend-backward time.  Needs the forward's hit log (a view whose log did not fit gets NaN gradients, not silently
 * unordered ones) and cannot be captured into a CUDA graph.  Returns the previous mode; other values only query. */
GSTAR_API int gstar_set_deterministic(int on);
/* Raw copy of one forward call's device header (24 32-bit words; private layout, gstar_common.cuh) -- diagnostics. */
GSTAR_API int gstar_debug_header(char* image_buffer, uint32_t* words24);
GSTAR_API int gstar_hit_log_state(char* image_buffer, uint64_t* slots_needed, uint64_t* slots_capacity, int* in_use);

/* ---- co-requisite of the drop-in: simple_knn._C.distCUDA2 (SimpleKNN::knn, simple-knn/simple_knn.cu:188-220) ----
 * out[i] = mean squared distance of point i to its three nearest OTHER points.  The caller bins the points into a
 * uniform grid (nx*ny*nz cells of edge `cell` starting at (ox,oy,oz)): `order` lists the point indices cell by cell,
 * cell_start[c] .. cell_start[c+1] is cell c's range in it (cell index = (z*ny + y)*nx + x).  See simple_knn/_C.py. */
GSTAR_API int gstar_knn3_mean_dist2(int P, const float* points, const int* order, const int* cell_start, int nx, int ny, int nz,
                                    float ox, float oy, float oz, float cell, float* out, void* stream);

/* ---- SURVEY 8f-4 (beyond the reference's operator boundary): the caller-side prologue of a mesh-bound SuGaR model --------
 * What gaustar_scene/sugar_model.py computes per render call with a few dozen torch kernels: `points` (:417-435), `scaling`
 * (:457-476), `quaternions` (:478-508) and `strengths` (:443-447) of P = F*K Gaussians bound K per face to a triangle mesh.
 * All pointers are device pointers; faces32 / faces64: exactly one non-NULL ([F,3] int32 or int64).  Backward: g_* are the
 * upstream gradients (NULL = zero); d_scales / d_cplx / d_dens are written, d_verts [Nv,3] is ACCUMULATED into (zero it first);
 * any d_* may be NULL.  gaustar_b200/sugar.py is the host side (autograd function + a patch for SuGaR instances). */
typedef struct gstar_sugar_args {
    int P, K;
    const float* verts;
    const int* faces32;
    const long long* faces64;
    const float* bary;      /* [K,3] barycentric coordinates of the K Gaussians of a face (sugar_model.py:176-226) */
    const float* scales;    /* [P,2] */
    const float* cplx;      /* [P,2] */
    const float* dens;      /* [P]   */
    float thickness, min_scale, max_scale;
    int has_min, has_max;
    float *points, *scaling, *quats, *opac;
    const float *g_points, *g_scaling, *g_quats, *g_opac;
    float *d_verts, *d_scales, *d_cplx, *d_dens;
} gstar_sugar_args;
GSTAR_API int gstar_sugar_prologue_forward(const gstar_sugar_args* a, void* stream);
GSTAR_API int gstar_sugar_prologue_backward(const gstar_sugar_args* a, void* stream);

/* ---- measurement hook: record `start`/`stop` (cudaEvent_t) around kernel stage `stage` of every
 * subsequent call on this thread (stage < 0 disables).  Stages: see gstar_stage_name(). ---- */
GSTAR_API int gstar_profile_stage(int stage, void* start_event, void* stop_event);
GSTAR_API const char* gstar_stage_name(int stage);
#define GSTAR_STAGE_PREPROCESS_FWD 0
#define GSTAR_STAGE_TILE_SCAN      1
#define GSTAR_STAGE_EMIT           2
#define GSTAR_STAGE_TILE_SORT      3
#define GSTAR_STAGE_BLEND_FWD      4
#define GSTAR_STAGE_BLEND_BWD      5
#define GSTAR_STAGE_PREPROCESS_BWD 6
#define GSTAR_NUM_STAGES           7

#ifdef __cplusplus
}
#endif
#endif /* GSTAR_RASTER_H */
