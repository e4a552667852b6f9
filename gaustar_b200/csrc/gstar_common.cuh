// gstar_common.cuh -- shared device-side definitions for the sm_100a surface-Gaussian rasterizer.
//
// Data layout in HBM:
//   GRec[P]       48-byte packed per-Gaussian record written by preprocess_fwd (what blending needs of a Gaussian);
//   GAux[P]       16 bytes per Gaussian for binning (depth, tile rect, radius).
//   entries[R]    unsorted per-tile segments of (depth bits, gaussian idx) pairs.
//   point_list[R] per-tile depth-sorted gaussian indices == the reference's sorted value list
//                 (DGR/cuda_rasterizer/rasterizer_impl.cu:303-308).
//   ranges[T]     [start,end) of every tile in point_list (rasterizer_impl.cu:116-138).
//   packed[R]     48-byte records in blend order: x y A B | C o r g | foot gid b slot  (foot = the alpha-bounds clipped to
//                 the tile, packed; gid = Gaussian index; slot = the instance's first hit-log slot).
//   hitlog[A]     16-byte GHit per (instance, footprint pixel), written by blend_fwd for every blended pair.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GSTAR_TILE 16          // DGR/cuda_rasterizer/config.h:16-17 (BLOCK_X, BLOCK_Y)
#define GSTAR_BATCH 256        // records staged in shared memory per pipeline stage
#define GSTAR_REC_BYTES 48     // stride of GRec in HBM
#define GSTAR_REC_SMEM 48      // bytes of a record the blend kernels need (bulk-copied)
#define GSTAR_GACC 12          // floats per Gaussian in the blend-gradient accumulator

struct __align__(16) GRec {
    float x, y;        // pixel centre (means2D)            forward.cu:233,252
    float A, B;        // conic.x, conic.y                  forward.cu:223
    float C, o;        // conic.z, opacity                  forward.cu:254
    float r, g;        // colour (SH result or colors_precomp)
    uint32_t bbox_x;   // int16 xmin | int16 xmax << 16 : exact-conservative alpha>=1/255 pixel bounds
    uint32_t bbox_y;   // int16 ymin | int16 ymax << 16
    float b;           // colour, third channel
    uint32_t flags;    // bit0..2: SH clamp mask (forward.cu:67-69)
};
// What binning needs of a Gaussian, in its own array: emit streams 16 bytes per Gaussian instead of whole records.
struct __align__(16) GAux {
    float depth;       // view-space z                      forward.cu:250
    uint32_t rect_min; // tile rect min x | y << 16         auxiliary.h:46-56
    uint32_t rect_max; // tile rect max x | y << 16 (exclusive)
    int32_t radius;    // forward.cu:251 (0 = culled)
};
static_assert(sizeof(GRec) == GSTAR_REC_BYTES && sizeof(GAux) == 16, "GRec / GAux layout");

// Small header kept at the start of the image buffer (device) -- counters of one forward call.
struct GHeader {
    uint32_t num_rendered;   // R = sum of tiles_touched
    uint32_t capacity;       // instances the binning buffer can hold
    uint32_t overflow;       // 1 if R > capacity (binning/blend skipped, host retries)
    uint32_t max_tile;       // longest tile list
    uint32_t log_overflow;   // 1: the hit log is disabled or too small for this view -> blend_bwd walks the lists instead
    uint32_t pad0[3];
    unsigned long long log_cursor;      // hit-log slots handed out by tile_sort (= sum of clipped footprint areas)
    unsigned long long log_capacity;    // hit-log slots the binning buffer holds
    unsigned long long off_point_list;  // byte offsets inside the binning buffer
    unsigned long long off_log;
    // tile_order lists the tiles longest first; class c = positions [cls_end[c-1], cls_end[c]) of it:
    // 0: >= 8192 instances, 1: 4096..8191, 2: 2048..4095, 3: 1..2047 (cls_end[3] = number of non-empty tiles).  The sort
    // kernel is persistent and pulls the tiles of a class through cls_cursor.
    uint32_t cls_end[4];
    uint32_t cls_cursor[4];
    // two feature passes blended in one (gstar_fwd_args::colors2): bytes of a hit-log row (16, or 32 with the second pass's
    // colour) and the byte offset, inside the binning buffer, of the second pass's per-pixel final colour (0: single pass)
    unsigned long long off_pixstate2;
    uint32_t log_row_bytes;
    uint32_t pad1;
};
static_assert(sizeof(GHeader) == 112, "GHeader layout");

// One hit-log slot per (instance, pixel of its clipped footprint): what the forward blend knew when it blended the
// pair -- transmittance in front of it and the colour accumulated up to and including it (blend.cu).
struct __align__(16) GHit {
    float c0, c1, c2, T;  // the order of blend_fwd's per-pixel state registers: one 128-bit store, no moves
};

namespace gstar {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// an mbarrier object must be invalidated before its memory is initialised again (re-init of a live object is undefined)
__device__ __forceinline__ void mbar_inval(uint64_t* bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the warp is parked by the hardware for up to `ns` instead of re-issuing the poll
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;  // the common case: the stage is already there
    for (uint32_t spins = 0; !mbar_try_wait_hint(bar, parity, 4000u); ++spins)
        if (spins > (1u << 22)) __trap();
}
// Waiting side of a long-latency hand-off (the producer waiting for its consumers): back off with nanosleep so
// that the spinning warp does not steal issue slots from the warps doing the work.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait_hint(bar, parity, 20000u); ++spins) {
        __nanosleep(512);
        if (spins > (1u << 22)) __trap();
    }
}
// TMA bulk copy global -> shared (linear, 16-byte granularity), completion on an mbarrier. SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Footprint of a record inside one tile: its alpha-bounds (GRec::bbox_x/y) clipped to the tile and the image, in
// tile-local pixel coordinates.  w <= 0 or h <= 0: empty.  tile_sort hands every instance w*h consecutive hit-log
// slots; slot of local pixel (lx, ly) = base + (ly - y0) * w + (lx - x0).  lim_x = min(15, W-1-tile_x0), same for y.
struct Foot {
    int x0, y0, w, h;
};
__device__ __forceinline__ Foot clip_foot(uint32_t bbx, uint32_t bby, int tile_x0, int tile_y0, int lim_x, int lim_y)
{
    Foot f;
    f.x0 = max((int)(short)(bbx & 0xffffu) - tile_x0, 0);
    f.y0 = max((int)(short)(bby & 0xffffu) - tile_y0, 0);
    f.w = min((int)(short)(bbx >> 16) - tile_x0, lim_x) - f.x0 + 1;
    f.h = min((int)(short)(bby >> 16) - tile_y0, lim_y) - f.y0 + 1;
    return f;
}

// The packed records carry the footprint already clipped (tile_sort computes it once per instance):
// x0 | y0 << 4 | w << 8 | h << 13, all tile-local, w and h in 0..16, 0 = the alpha-bounds miss the tile.
__device__ __forceinline__ uint32_t pack_foot(const Foot& f)
{
    if (f.w <= 0 || f.h <= 0) return 0u;
    return (uint32_t)f.x0 | ((uint32_t)f.y0 << 4) | ((uint32_t)f.w << 8) | ((uint32_t)f.h << 13);
}
__device__ __forceinline__ Foot unpack_foot(uint32_t v)
{
    Foot f;
    f.x0 = (int)(v & 15u); f.y0 = (int)((v >> 4) & 15u); f.w = (int)((v >> 8) & 31u); f.h = (int)((v >> 13) & 31u);
    return f;
}

// 128-bit vector reduction to global memory (SASS: REDG.E.ADD.F32x4): four fp32 adds in one instruction
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace gstar
