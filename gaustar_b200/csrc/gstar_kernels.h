// gstar_kernels.h -- host-visible launch interfaces of the kernels (internal to libgstar_raster.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct GRec;
struct GAux;
struct GHeader;

namespace gstar {

struct PreFwdParams {
    int P, D, M, W, H, gx, gy;
    const float* means3D;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int prefiltered;
    GRec* recs;
    GAux* aux;
    int* radii;
    uint32_t* tile_count;
};

struct PreBwdParams {
    int P, D, M, W, H;
    const float* means3D;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* shs;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    const int* radii;
    const GRec* recs;
    const GAux* aux;
    const float* gacc;  // [P][12] blend-stage gradient accumulator
    float* dL_dmean2D;
    float* dL_dconic;
    float* dL_dopacity;
    float* dL_dcolor;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
    int accumulate;  // += into the parameter gradients (mean3D, scale, rot, sh, opacity): 1 plain read-modify-write, 2 atomic (concurrent calls)
};

struct BinParams {
    int P, gx, gy, num_tiles, W, H;
    const GRec* recs;
    const GAux* aux;
    GHeader* hdr;          // device header (image buffer)
    uint32_t* tile_count;  // [T] histogram from preprocess
    uint32_t* tile_cursor; // [T] running write position per tile
    uint32_t* ranges;      // [T][2]
    uint32_t* tile_order;  // [T] all tiles, longest list first (work order of the blend kernels)
    unsigned char* tile_lanes;  // [T] lanes per instance of the gather backward (1, 2, 4 or 8), chosen by tile_sort
    uint2* entries;        // [capacity] (depth bits, idx), tile-segmented
    uint32_t* point_list;  // [capacity] sorted gaussian ids
    unsigned char* packed; // [capacity][48] tile-contiguous packed records (GRec[0:44] + first hit-log slot), sorted order
    uint32_t capacity;
    unsigned long long log_capacity;    // hit-log slots available behind `entries` (0: log disabled)
    unsigned long long off_point_list;  // byte offsets inside the binning buffer, recorded in the header
    unsigned long long off_log;
    unsigned long long off_pixstate2 = 0;  // (two feature passes in one: see GHeader)
    uint32_t log_row_bytes = 16;
    volatile uint32_t* host_counts;  // mapped pinned host memory: [0]=R, [1]=overflow, [2]=max tile, [4..5]=hit-log slots needed
};

struct BlendParams {
    int W, H, gx, gy;
    const GRec* recs;
    const GHeader* hdr;
    const uint32_t* ranges;
    const unsigned char* packed;  // [R][48] tile-contiguous packed records in blend order (= start of the binning buffer)
    const uint32_t* tile_order;
    const unsigned char* tile_lanes;  // [T] lanes per instance for k_blend_bwd_gather (NULL: 1)
    float4* pixstate;             // [H*W] (C_r, C_g, C_b, T) per pixel as the forward left it (colour without background)
    volatile uint32_t* host_counts;
    const float* bg;
    // forward
    float* out_color;
    float* final_T;
    uint32_t* n_contrib;
    // backward
    const float* dL_dpix;
    float* gacc;
    // deterministic test mode (gstar_set_deterministic): [R][12] one row of raw moments per record, written with plain stores by
    // k_blend_bwd_gather instead of its reductions into gacc; launch_det_reduce then sums every Gaussian's rows in a fixed order
    float* det_partial = nullptr;
    const GAux* aux = nullptr;
    int P = 0;
    // two feature passes in one blend (blend.cu, NP == 2): per-Gaussian colours [P,3] and background of the second pass, its image,
    // and in the backward its upstream gradient; NULL: single pass
    const float* colors2 = nullptr;   // [P,4] (a float4 per Gaussian; channels >= ch2 must be finite, e.g. zero)
    const float* bg2 = nullptr;       // [ch2]
    float* out_color2 = nullptr;      // [ch2,H,W]
    const float* dL_dpix2 = nullptr;  // [ch2,H,W]
    int ch2 = 3;                      // channels of the second pass (1..4)
    float* gacc2 = nullptr;           // backward, ch2 == 4: [P] colour moment of the fourth channel (the first three: gacc slots 9..11)
};

void launch_preprocess_fwd(const PreFwdParams& p, cudaStream_t s);
void launch_preprocess_bwd(const PreBwdParams& p, cudaStream_t s);
void launch_mark_visible(int P, const float* means3D, const float* vm, unsigned char* present, cudaStream_t s);
void launch_geom_unpack(const GRec* recs, const GAux* aux, int P, float* depths, float* means2D, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                        unsigned char* clamped, cudaStream_t s);

void launch_tile_scan(const BinParams& p, cudaStream_t s);
void launch_emit(const BinParams& p, cudaStream_t s);
void launch_tile_sort(const BinParams& p, cudaStream_t s);
void launch_publish_log(const GHeader* hdr, volatile uint32_t* host_counts, cudaStream_t s);  // host_counts[6..7] = hit-log slots the view needs
int  tile_sort_setup();  // one-time cudaFuncSetAttribute calls; returns cudaError_t
int  preprocess_setup();
int  blend_setup();

void launch_blend_fwd(const BlendParams& p, cudaStream_t s);
void launch_blend_bwd(const BlendParams& p, cudaStream_t s);         // walk-back path (no hit log)
void launch_blend_bwd_gather(const BlendParams& p, cudaStream_t s);  // instance-parallel path over the hit log
void launch_det_reduce(const BlendParams& p, cudaStream_t s);        // deterministic mode: gacc[g] += sum of g's rows of det_partial, tile by tile
// re-blend (shared geometry): copy the packed stream with the colour fields replaced by colors[gid] and rebuild the sorted
// value list from the records' ids; optionally switch the new header's hit log off
// the cameras of the source call and of the re-blend (device pointers to 16 floats each; src_view == NULL: not checked)
struct CameraCheck {
    const float* src_view;
    const float* src_proj;
    const float* view;
    const float* proj;
};
void launch_recolor(const unsigned char* src_packed, unsigned char* dst_packed, uint32_t* dst_point_list, uint32_t R, const float* colors, GHeader* hdr,
                    int disable_log, const CameraCheck& cam, cudaStream_t s);
// fills `out` with NaN if the header says the re-blend was refused on the device (overflow == 2) or, with any_overflow, if
// the call found its instance capacity too small (a replayed CUDA graph cannot grow the buffer)
void launch_poison(const GHeader* hdr, float* out, size_t n, cudaStream_t s, int any_overflow = 0);
// two-pass backward: NaN into the moment scratch if the forward ended without a hit log after all (there is no walk-back kernel over
// two passes; gstar_raster_forward guarantees the log, this makes a violated guarantee loud instead of a silent zero gradient)
void launch_poison_no_log(const GHeader* hdr, float* gacc, size_t n, cudaStream_t s);

// mesh-bound SuGaR prologue (sugar_prologue.cu).  Forward fills points / scaling / quats / opac; backward reads the
// g_* upstream gradients (NULL: zero) and fills d_scales / d_cplx / d_dens and ACCUMULATES into d_verts (zeroed by the caller).
struct SugarParams {
    int P, K;                   // Gaussians, Gaussians per face
    const float* verts;         // [Nv,3]
    const int* faces32;         // [F,3] (one of the two is non-NULL)
    const long long* faces64;
    const float* bary;          // [K,3]
    const float* scales;        // [P,2]  log of the in-plane scales
    const float* cplx;          // [P,2]  in-plane rotation as a complex number (normalised in-kernel)
    const float* dens;          // [P]    logit of the opacity
    float thickness, min_scale, max_scale;
    int has_min, has_max;
    float *points, *scaling, *quats, *opac;                       // forward outputs [P,3] [P,3] [P,4] [P]
    const float *g_points, *g_scaling, *g_quats, *g_opac;         // backward inputs
    float *d_verts, *d_scales, *d_cplx, *d_dens;                  // backward outputs
};
void launch_sugar_prologue_fwd(const SugarParams& s, cudaStream_t st);
void launch_sugar_prologue_bwd(const SugarParams& s, cudaStream_t st);

}  // namespace gstar
