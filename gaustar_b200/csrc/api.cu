// api.cu -- the C ABI of libgstar_raster.so (include/gstar_raster.h): orchestration only.
//
// Forward pipeline (replaces CudaRasterizer::Rasterizer::forward, rasterizer_impl.cu:198-336):
//   memset(tile histogram) -> K1 preprocess_fwd -> K2 tile_scan -> K3 emit -> K4 tile_sort -> K6 blend_fwd
// The reference blocks the host on a D2H copy of num_rendered right after its prefix sum
// (rasterizer_impl.cu:281) and only then sizes the binning buffer.  Here the binning buffer is
// provisioned from the previous call's instance count, K2 publishes R and an overflow flag through
// mapped pinned memory, all remaining kernels are enqueued immediately, and the host waits on an
// event recorded right behind K2 only AFTER it has enqueued everything -- the GPU never idles, and
// the exact R is still returned.  If R exceeded the provision (rare) the kernels behind K2 no-op'd
// on the device flag and binning + blend are re-enqueued with an exact-size buffer.
//
// Beyond the reference: gstar_raster_reblend (shared-geometry second pass: k_recolor -> K6 on the first pass's sorted
// record stream), gstar_bwd_args.blend_only (K7 alone, moments left in the caller's scratch), and a forward that stays
// on the device while its stream is being captured into a CUDA graph.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: the ranges below cost a pointer test when no tool is attached
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>

#include "../../include/gstar_raster.h"
#include "gstar_common.cuh"
#include "gstar_kernels.h"
#include "recent_calls.h"

namespace {

thread_local std::string t_err;

int fail(int code, const std::string& msg)
{
    t_err = msg;
    return code;
}

#define CU_OK(expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return fail(GSTAR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int MAX_DEV = 64;
struct DevCtx {
    bool inited = false;
    uint32_t* host_counts = nullptr;  // pinned + mapped: [0]=R [1]=overflow [2]=max tile
    uint32_t* host_counts_dev = nullptr;
    cudaEvent_t scan_done = nullptr;
    double estimate = 0.0;  // running provision for R (instances)
    int small_streak = 0;
    bool have_estimate = false;
    double log_estimate = 0.0;  // running provision for the hit log (slots); 0 with log_have: log off for this workload
    int log_small_streak = 0;
    bool log_have = false;
    RecentCalls calls;  // layouts of the most recent forward / re-blend calls (recent_calls.h)
};
int g_hit_log_mode = -1;  // -1: read GSTAR_HIT_LOG on first use; 0 off; 1 auto
double g_hit_log_max_slots = 0.0;

bool hit_log_enabled()
{
    if (g_hit_log_mode < 0) {
        const char* e = getenv("GSTAR_HIT_LOG");
        g_hit_log_mode = (e && e[0] == '0') ? 0 : 1;
    }
    if (g_hit_log_max_slots == 0.0) {
        const char* e = getenv("GSTAR_HIT_LOG_MAX_MB");  // ceiling of the log (default 8 GiB of the 180 GB of HBM)
        const double mb = e ? atof(e) : 8192.0;
        g_hit_log_max_slots = std::max(mb, 1.0) * 1048576.0 / sizeof(GHit);
    }
    return g_hit_log_mode != 0;
}
constexpr double LOG_QUANTUM = 4194304.0;  // slots (64 MiB): keeps the binning buffer at one size in steady state
thread_local DevCtx t_ctx[MAX_DEV];

struct Profile {
    int stage = -1;
    cudaEvent_t start = nullptr, stop = nullptr;
};
thread_local Profile t_prof;

// NVTX ranges (SURVEY section 5 "tracing"): one per entry point ("gstar_raster_forward", ...) and, nested inside it, one per pipeline
// stage ("gstar::blend_fwd", ...), so that a timeline groups the kernels of a view and `ncu --nvtx --nvtx-include "gstar::blend_fwd/"`
// selects one stage's kernels.  The reference has no ranges.
const char* const STAGE_RANGE[GSTAR_NUM_STAGES] = {"gstar::preprocess_fwd", "gstar::tile_scan", "gstar::emit", "gstar::tile_sort", "gstar::blend_fwd",
                                                   "gstar::blend_bwd", "gstar::preprocess_bwd"};
struct CallRange {
    explicit CallRange(const char* name) { nvtxRangePushA(name); }
    ~CallRange() { nvtxRangePop(); }
};

struct StageScope {
    int stage;
    cudaStream_t s;
    StageScope(int st, cudaStream_t stream) : stage(st), s(stream)
    {
        nvtxRangePushA(STAGE_RANGE[stage]);
        if (t_prof.stage == stage && t_prof.start) cudaEventRecord(t_prof.start, s);
    }
    ~StageScope()
    {
        if (t_prof.stage == stage && t_prof.stop) cudaEventRecord(t_prof.stop, s);
        nvtxRangePop();
    }
};

int g_deterministic = -1;  // -1: read GSTAR_DETERMINISTIC on first use
bool deterministic_mode()
{
    if (g_deterministic < 0) {
        const char* e = getenv("GSTAR_DETERMINISTIC");
        g_deterministic = (e && e[0] == '1') ? 1 : 0;
    }
    return g_deterministic != 0;
}

int get_ctx(DevCtx** out, bool capturing = false)
{
    int dev = 0;
    CU_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail(GSTAR_ERR_INVALID, "device index out of range");
    DevCtx& c = t_ctx[dev];
    if (!c.inited) {
        if (capturing)  // pinned allocation / event creation / function attributes are not legal inside a capture
            return fail(GSTAR_ERR_INVALID, "the first call of a thread on a device must not be captured into a CUDA graph");
        CU_OK(cudaHostAlloc((void**)&c.host_counts, 64, cudaHostAllocMapped));
        memset(c.host_counts, 0, 64);
        CU_OK(cudaHostGetDevicePointer((void**)&c.host_counts_dev, c.host_counts, 0));
        CU_OK(cudaEventCreateWithFlags(&c.scan_done, cudaEventDisableTiming));
        CU_OK((cudaError_t)gstar::tile_sort_setup());
        CU_OK((cudaError_t)gstar::preprocess_setup());
        CU_OK((cudaError_t)gstar::blend_setup());
        c.inited = true;
    }
    *out = &c;
    return 0;
}

// ---- private layouts of the three opaque buffers ----
struct ImgLayout {
    size_t hdr, final_T, n_contrib, pixstate, ranges, tile_count, tile_cursor, tile_order, tile_lanes, total;
};
ImgLayout img_layout(int W, int H)
{
    const size_t npix = (size_t)W * H;
    const size_t T = (size_t)((W + GSTAR_TILE - 1) / GSTAR_TILE) * ((H + GSTAR_TILE - 1) / GSTAR_TILE);
    ImgLayout L;
    size_t o = 0;
    L.hdr = o; o = align_up(o + sizeof(GHeader), 128);
    L.final_T = o; o = align_up(o + npix * 4, 128);
    L.n_contrib = o; o = align_up(o + npix * 4, 128);
    L.pixstate = o; o = align_up(o + npix * 16, 128);
    L.ranges = o; o = align_up(o + T * 8, 128);
    L.tile_count = o; o = align_up(o + T * 4, 128);
    L.tile_cursor = o; o = align_up(o + T * 4, 128);
    L.tile_order = o; o = align_up(o + T * 4, 128);
    L.tile_lanes = o; o = align_up(o + T, 128);
    L.total = o;
    return L;
}
struct BinLayout {
    size_t packed, point_list, entries, log, pixstate2, total;
};
// row_bytes: 16, or 32 for a two-pass forward, which also keeps the second pass's final colour per pixel (npix2 pixels) behind the log
BinLayout bin_layout(size_t cap, size_t log_slots, size_t row_bytes = sizeof(GHit), size_t npix2 = 0)
{
    BinLayout L;
    L.packed = 0;  // first, so that backward needs no capacity to find it (the other offsets travel in the header)
    L.point_list = align_up(cap * GSTAR_REC_SMEM, 128);
    L.entries = align_up(L.point_list + cap * 4, 128);
    L.log = align_up(L.entries + cap * 8, 128);
    L.pixstate2 = align_up(L.log + log_slots * row_bytes, 128);
    L.total = align_up(L.pixstate2 + npix2 * 16, 128);
    return L;
}

int debug_sync(int debug, const char* what)
{
    if (!debug) {
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return 0;
    }
    cudaError_t e = cudaDeviceSynchronize();  // auxiliary.h:166-173 CHECK_CUDA
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, std::string("[CUDA ERROR] in ") + what + ": " + cudaGetErrorString(e));
    return 0;
}
#define STAGE_CHECK(what)                        \
    do {                                         \
        int rc_ = debug_sync(a->debug, what);    \
        if (rc_ < 0) return rc_;                 \
    } while (0)

}  // namespace

extern "C" {

const char* gstar_last_error(void) { return t_err.c_str(); }
int gstar_abi_version(void) { return GSTAR_ABI_VERSION; }

static size_t geom_aux_offset(int P) { return align_up((size_t)std::max(P, 0) * sizeof(GRec), 128); }
size_t gstar_geom_bytes(int P) { return geom_aux_offset(P) + align_up((size_t)std::max(P, 0) * sizeof(GAux), 128) + 128; }
size_t gstar_image_bytes(int width, int height) { return img_layout(width, height).total + 128; }
size_t gstar_binning_bytes(size_t cap) { return bin_layout(cap, 0).total + 128; }

int gstar_set_hit_log(int mode)
{
    hit_log_enabled();
    const int old = g_hit_log_mode;
    if (mode == 0 || mode == 1) g_hit_log_mode = mode;
    return old;
}

int gstar_set_deterministic(int on)
{
    const int old = deterministic_mode() ? 1 : 0;
    if (on == 0 || on == 1) g_deterministic = on;
    return old;
}

const char* gstar_stage_name(int stage)
{
    static const char* names[GSTAR_NUM_STAGES] = {"preprocess_fwd", "tile_scan", "emit", "tile_sort", "blend_fwd", "blend_bwd", "preprocess_bwd"};
    return (stage >= 0 && stage < GSTAR_NUM_STAGES) ? names[stage] : "";
}

int gstar_profile_stage(int stage, void* start_event, void* stop_event)
{
    t_prof.stage = stage;
    t_prof.start = (cudaEvent_t)start_event;
    t_prof.stop = (cudaEvent_t)stop_event;
    return 0;
}

static inline char* aligned128(char* p) { return (char*)align_up((size_t)p, 128); }

int gstar_raster_forward(const gstar_fwd_args* a, gstar_alloc_fn geom_alloc, void* geom_user, gstar_alloc_fn binning_alloc,
                         void* binning_user, gstar_alloc_fn image_alloc, void* image_user, void* stream_)
{
    using namespace gstar;
    CallRange call_range("gstar_raster_forward");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || !geom_alloc || !binning_alloc || !image_alloc) return fail(GSTAR_ERR_INVALID, "null argument");
    if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(GSTAR_ERR_INVALID, "bad sizes");
    if (a->P == 0) return 0;
    if (!a->colors_precomp && !a->shs) return fail(GSTAR_ERR_NONRGB, "For non-RGB, provide precomputed Gaussian colors!");
    if (!a->cov3D_precomp && (!a->scales || !a->rotations))
        return fail(GSTAR_ERR_INVALID, "provide scales+rotations or cov3D_precomp");
    if (a->width > 32767 || a->height > 32767) return fail(GSTAR_ERR_INVALID, "image larger than 32767 pixels");
    const bool dual = a->colors2 != nullptr;
    if (dual != (a->background2 != nullptr) || dual != (a->out_color2 != nullptr))
        return fail(GSTAR_ERR_INVALID, "two-pass forward: colors2, background2 and out_color2 go together");
    const int ch2 = dual ? (a->channels2 == 0 ? 3 : a->channels2) : 3;
    if (ch2 < 1 || ch2 > 4) return fail(GSTAR_ERR_INVALID, "two-pass forward: channels2 must be 1..4");
    if (dual && !a->forward_only && !hit_log_enabled())
        return fail(GSTAR_ERR_NOLOG, "two-pass forward: the hit log is switched off (gstar_set_hit_log / GSTAR_HIT_LOG); re-blend the second pass instead");
    // CUDA-graph capture: while `stream` is being captured the forward stays entirely on the device -- no
    // event wait, no read of the instance count.  The binning buffer then has the size the earlier un-captured calls of
    // this thread provisioned, the return value is that capacity (an upper bound of num_rendered, good for the matching
    // backward), and a replay whose view needs more leaves the header's overflow flag set (gstar_debug_header) and the
    // outputs untouched by the later kernels.
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    CU_OK(cudaStreamIsCapturing(stream, &cap_status));
    const bool capturing = cap_status != cudaStreamCaptureStatusNone;
    if (capturing && a->debug) return fail(GSTAR_ERR_INVALID, "debug mode synchronizes the device: not available while capturing a CUDA graph");
    if (capturing && dual) return fail(GSTAR_ERR_INVALID, "a two-pass forward reads the hit-log need back on the host: not available while capturing a CUDA graph");
    DevCtx* ctx;
    int rc = get_ctx(&ctx, capturing);
    if (rc < 0) return rc;
    if (capturing && !(ctx->have_estimate && ctx->estimate >= 1.0))
        return fail(GSTAR_ERR_INVALID, "a captured forward takes its instance capacity from earlier un-captured calls of this thread: run one first");

    const int W = a->width, H = a->height;
    const int gx = (W + GSTAR_TILE - 1) / GSTAR_TILE, gy = (H + GSTAR_TILE - 1) / GSTAR_TILE;
    const int T = gx * gy;
    char* geom = geom_alloc(geom_user, gstar_geom_bytes(a->P));
    char* img = image_alloc(image_user, gstar_image_bytes(W, H));
    if (!geom || !img) return fail(GSTAR_ERR_ALLOC, "buffer callback returned NULL");
    geom = aligned128(geom);
    img = aligned128(img);
    const ImgLayout IL = img_layout(W, H);
    GHeader* hdr = (GHeader*)(img + IL.hdr);
    uint32_t* tile_count = (uint32_t*)(img + IL.tile_count);

    CU_OK(cudaMemsetAsync(tile_count, 0, (size_t)T * 4, stream));

    PreFwdParams pp;
    pp.P = a->P; pp.D = a->D; pp.M = a->M; pp.W = W; pp.H = H; pp.gx = gx; pp.gy = gy;
    pp.means3D = a->means3D; pp.scales = a->scales; pp.scale_modifier = a->scale_modifier; pp.rotations = a->rotations;
    pp.opacities = a->opacities; pp.shs = a->shs; pp.cov3D_precomp = a->cov3D_precomp; pp.colors_precomp = a->colors_precomp;
    pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.campos = a->cam_pos;
    pp.tan_fovx = a->tan_fovx; pp.tan_fovy = a->tan_fovy;
    pp.focal_y = H / (2.0f * a->tan_fovy);  // rasterizer_impl.cu:222-223
    pp.focal_x = W / (2.0f * a->tan_fovx);
    pp.prefiltered = a->prefiltered;
    pp.recs = (GRec*)geom; pp.aux = (GAux*)(geom + geom_aux_offset(a->P)); pp.radii = a->radii; pp.tile_count = tile_count;
    {
        StageScope sc(GSTAR_STAGE_PREPROCESS_FWD, stream);
        launch_preprocess_fwd(pp, stream);
    }
    STAGE_CHECK("preprocess_fwd");

    BinParams bp;
    bp.P = a->P; bp.gx = gx; bp.gy = gy; bp.num_tiles = T; bp.W = W; bp.H = H;
    bp.recs = (const GRec*)geom; bp.aux = (const GAux*)(geom + geom_aux_offset(a->P)); bp.hdr = hdr; bp.tile_count = tile_count;
    bp.tile_cursor = (uint32_t*)(img + IL.tile_cursor);
    bp.ranges = (uint32_t*)(img + IL.ranges);
    bp.tile_order = (uint32_t*)(img + IL.tile_order);
    bp.tile_lanes = (unsigned char*)(img + IL.tile_lanes);
    // a captured call must not touch the thread's pinned counters: its replays would race with un-captured calls reading them
    bp.host_counts = capturing ? nullptr : ctx->host_counts_dev;

    BlendParams bl;
    bl.W = W; bl.H = H; bl.gx = gx; bl.gy = gy;
    bl.recs = (const GRec*)geom; bl.hdr = hdr; bl.ranges = bp.ranges; bl.tile_order = bp.tile_order; bl.bg = a->background;
    bl.out_color = a->out_color; bl.final_T = (float*)(img + IL.final_T); bl.n_contrib = (uint32_t*)(img + IL.n_contrib);
    bl.dL_dpix = nullptr; bl.gacc = nullptr; bl.tile_lanes = bp.tile_lanes;
    bl.pixstate = (float4*)(img + IL.pixstate); bl.host_counts = capturing ? nullptr : ctx->host_counts_dev;
    bl.colors2 = a->colors2; bl.bg2 = a->background2; bl.out_color2 = a->out_color2; bl.ch2 = ch2;
    const size_t row_bytes = dual ? 2 * sizeof(GHit) : sizeof(GHit);
    hit_log_enabled();  // (reads the environment on first use: g_hit_log_max_slots)
    const double max_slots = g_hit_log_max_slots * (double)sizeof(GHit) / (double)row_bytes;
    const size_t npix2 = dual ? (size_t)W * H : 0;

    // Hit-log provision: the slots the previous view needed (published by its blend_fwd; a hint, it may lag) with the
    // same grow-at-once / shrink-slowly / quantised policy as the instance count.  A view whose log does not fit simply
    // takes the walk-back backward (device-side flag), so a wrong guess costs speed, never correctness.
    const bool use_log = hit_log_enabled() && !a->forward_only;
    if (use_log) {
        const double need = (double)(((unsigned long long)ctx->host_counts[5] << 32) | ctx->host_counts[4]);
        if (need > 0.0) {
            const double want = need * 1.25 > max_slots ? 0.0 : std::ceil((need * 1.25 + 65536.0) / LOG_QUANTUM) * LOG_QUANTUM;
            if (!ctx->log_have || want > ctx->log_estimate || (want == 0.0 && need > max_slots)) {
                ctx->log_estimate = want;
                ctx->log_small_streak = 0;
            } else if (want * 2.0 < ctx->log_estimate) {
                if (++ctx->log_small_streak >= 32) {
                    ctx->log_estimate = want;
                    ctx->log_small_streak = 0;
                }
            } else {
                ctx->log_small_streak = 0;
            }
            ctx->log_have = true;
        }
    }

    size_t cap = ctx->have_estimate ? (size_t)ctx->estimate : 0;
    uint32_t R = 0;
    size_t used_cap = 0, used_log_slots = 0;
    // A two-pass forward that a backward will follow MUST end with a hit log (there is no walk-back kernel over two passes): it
    // waits for the sort's slot count and, if the provision was too small, grows it and goes round once more.
    const bool need_log = dual && use_log;
    double log_force = 0.0;  // slots the next attempt must provide
    for (int attempt = 0; attempt < 3; attempt++) {
        char* bin = nullptr;
        if (cap > 0) {
            if (cap > 0xfffffff0ull) return fail(GSTAR_ERR_INVALID, "more than 2^32 instances");
        }
        size_t log_slots = 0;
        if (use_log && cap > 0) {
            double guess = ctx->log_have ? ctx->log_estimate : std::ceil((double)cap * 24.0 / LOG_QUANTUM) * LOG_QUANTUM;
            if (log_force > guess) guess = log_force;
            log_slots = (guess <= max_slots && guess < 4.0e9) ? (size_t)guess : 0;
            if (need_log && log_slots == 0)
                return fail(GSTAR_ERR_NOLOG, "two-pass forward: this view's hit log exceeds GSTAR_HIT_LOG_MAX_MB; re-blend the second pass instead");
        }
        const BinLayout BL = bin_layout(cap, log_slots, row_bytes, npix2);
        used_cap = cap; used_log_slots = log_slots;
        bp.capacity = (uint32_t)cap;
        bp.log_capacity = log_slots; bp.off_point_list = BL.point_list; bp.off_log = BL.log;
        bp.off_pixstate2 = dual ? BL.pixstate2 : 0; bp.log_row_bytes = (uint32_t)row_bytes;
        if (cap > 0) {
            bin = binning_alloc(binning_user, BL.total + 128);
            if (!bin) return fail(GSTAR_ERR_ALLOC, "binning buffer callback returned NULL");
            bin = aligned128(bin);
        }
        bp.point_list = bin ? (uint32_t*)(bin + BL.point_list) : nullptr;
        bp.packed = bin ? (unsigned char*)(bin + BL.packed) : nullptr;
        bp.entries = bin ? (uint2*)(bin + BL.entries) : nullptr;
        bl.packed = bp.packed;
        {
            StageScope sc(GSTAR_STAGE_TILE_SCAN, stream);
            launch_tile_scan(bp, stream);
        }
        if (!capturing) CU_OK(cudaEventRecord(ctx->scan_done, stream));
        STAGE_CHECK("tile_scan");
        if (cap > 0) {
            {
                StageScope sc(GSTAR_STAGE_EMIT, stream);
                launch_emit(bp, stream);
            }
            STAGE_CHECK("emit");
            {
                StageScope sc(GSTAR_STAGE_TILE_SORT, stream);
                launch_tile_sort(bp, stream);
                if (need_log) {  // the slots this view needs, to the host, before the blend starts
                    launch_publish_log(hdr, ctx->host_counts_dev, stream);
                    CU_OK(cudaEventRecord(ctx->scan_done, stream));
                }
            }
            STAGE_CHECK("tile_sort");
            {
                StageScope sc(GSTAR_STAGE_BLEND_FWD, stream);
                launch_blend_fwd(bl, stream);
            }
            STAGE_CHECK("blend_fwd");
        }
        if (capturing) {
            // a replay whose view needs more instances than `cap` leaves the header's overflow flag set and blends nothing:
            // the image is then filled with NaN instead of being returned uninitialised
            if (cap > 0) launch_poison(hdr, a->out_color, (size_t)3 * W * H, stream, 1);
            R = (uint32_t)cap;
            break;
        }
        // everything is enqueued; only now wait for the scan result (the GPU keeps working)
        CU_OK(cudaEventSynchronize(ctx->scan_done));
        R = ctx->host_counts[0];
        const bool overflow = R > cap;
        // Provision for the next call: 25 % headroom over the high-water mark, quantised to 256 Ki instances so that
        // the binning buffer keeps ONE size in steady state (a size that changes a little every call defeats the
        // caller's caching allocator: every call would cudaMalloc a new block).  Shrinks only after 32 consecutive
        // calls that needed less than half of it.
        {
            const double want = std::ceil(((double)R * 1.25 + 65536.0) / 262144.0) * 262144.0;
            if (want > ctx->estimate) {
                ctx->estimate = want;
                ctx->small_streak = 0;
            } else if (want * 2.0 < ctx->estimate) {
                if (++ctx->small_streak >= 32) {
                    ctx->estimate = want;
                    ctx->small_streak = 0;
                }
            } else {
                ctx->small_streak = 0;
            }
            ctx->have_estimate = true;
        }
        if (!overflow) {
            if (need_log && cap > 0) {  // (the event waited for above was recorded behind the sort)
                const double need = (double)(((unsigned long long)ctx->host_counts[7] << 32) | ctx->host_counts[6]);  // (k_publish_log's own words)
                if (need > (double)log_slots) {
                    if (attempt == 2) return fail(GSTAR_ERR_INVALID, "hit-log need changed between attempts");
                    log_force = std::ceil((need * 1.25 + 65536.0) / LOG_QUANTUM) * LOG_QUANTUM;
                    if (log_force > max_slots) log_force = std::ceil(need / 65536.0) * 65536.0;  // without headroom, if that still fits under the cap
                    ctx->log_estimate = std::max(ctx->log_estimate, log_force);
                    ctx->log_have = true;
                    continue;  // everything behind the scan no-op'd its log writes; binning + blend are re-enqueued with a log that fits
                }
            }
            break;
        }
        if (attempt == 2) return fail(GSTAR_ERR_INVALID, "instance count changed between attempts");
        cap = (size_t)R;  // exact size; tile_scan reruns to reset cursors and the device flag
    }
    if (R == 0) {
        // no instance anywhere: the image is the background. (with cap == 0 no blend was launched)
        if (cap == 0) {
            // cheapest correct path: run blend with empty ranges
            char* bin = binning_alloc(binning_user, gstar_binning_bytes(1));
            if (!bin) return fail(GSTAR_ERR_ALLOC, "binning buffer callback returned NULL");
            bl.packed = (unsigned char*)aligned128(bin);
            bp.capacity = 1; bp.log_capacity = 0; bp.off_point_list = 0; bp.off_log = 0; bp.off_pixstate2 = 0;
            launch_tile_scan(bp, stream);
            launch_blend_fwd(bl, stream);
            STAGE_CHECK("blend_fwd(empty)");
            used_cap = 1; used_log_slots = 0;
        }
    }
    ctx->calls.remember(img, used_cap, used_log_slots, R, a->P, W, H);
    return (int)R;
}

// Shared-geometry re-blend.  GauSTAR rasterizes the same Gaussians from the same camera twice per training
// step -- RGB, then depth as three equal channels (refine.py:552-564, :607-616) -- and the reference repeats preprocess,
// duplicateWithKeys and the radix sort for the second call although only `colors_precomp` changed.  Here the second call
// takes the first call's sorted, tile-contiguous record stream, rewrites its colour fields (k_recolor) into a binning
// buffer of its own and blends.  The result is what gstar_raster_forward would return for (same geometry inputs,
// colors_precomp): same records in the same order through the same blend kernel.
int gstar_raster_reblend(const gstar_reblend_args* a, gstar_alloc_fn binning_alloc, void* binning_user, gstar_alloc_fn image_alloc, void* image_user,
                         void* stream_)
{
    using namespace gstar;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || !binning_alloc || !image_alloc) return fail(GSTAR_ERR_INVALID, "null argument");
    if (a->P <= 0 || a->width <= 0 || a->height <= 0) return fail(GSTAR_ERR_INVALID, "bad sizes");
    if (!a->colors_precomp || !a->background || !a->out_color) return fail(GSTAR_ERR_INVALID, "re-blend needs colors_precomp, background and out_color");
    if (!a->src_image_buffer || !a->src_binning_buffer) return fail(GSTAR_ERR_INVALID, "re-blend needs the source call's binning and image buffers");
    DevCtx* ctx;
    int rc = get_ctx(&ctx);
    if (rc < 0) return rc;
    const char* src_img = aligned128((char*)a->src_image_buffer);
    const RecentCalls::Entry* src = ctx->calls.find(src_img);
    if (!src || src->P != a->P || src->W != a->width || src->H != a->height)
        return fail(GSTAR_ERR_INVALID, "re-blend: the source buffers do not belong to a recent forward call of this thread with the same P, width and height");
    const int W = a->width, H = a->height;
    const int gx = (W + GSTAR_TILE - 1) / GSTAR_TILE, gy = (H + GSTAR_TILE - 1) / GSTAR_TILE;
    const ImgLayout IL = img_layout(W, H);
    // same layout as the source call (the header's offsets stay valid); an inference re-blend carries no hit log
    const size_t cap = src->cap, R = src->R;
    const size_t log_slots = a->forward_only ? 0 : src->log_slots;
    const BinLayout BL = bin_layout(cap, log_slots);  // the offsets do not depend on the log's size
    const size_t bin_bytes = BL.total + 128;
    char* img = image_alloc(image_user, gstar_image_bytes(W, H));
    char* bin = binning_alloc(binning_user, bin_bytes);
    if (!img || !bin) return fail(GSTAR_ERR_ALLOC, "buffer callback returned NULL");
    img = aligned128(img);
    bin = aligned128(bin);
    const char* src_bin = aligned128((char*)a->src_binning_buffer);
    // header + everything binning left in the image buffer (ranges, tile order, gather lanes): a few hundred KB
    CU_OK(cudaMemcpyAsync(img + IL.hdr, src_img + IL.hdr, sizeof(GHeader), cudaMemcpyDeviceToDevice, stream));
    CU_OK(cudaMemcpyAsync(img + IL.ranges, src_img + IL.ranges, IL.total - IL.ranges, cudaMemcpyDeviceToDevice, stream));
    GHeader* hdr = (GHeader*)(img + IL.hdr);
    {
        StageScope sc(GSTAR_STAGE_TILE_SORT, stream);  // takes the place of preprocess .. sort
        CameraCheck cam = {nullptr, nullptr, nullptr, nullptr};
        if (a->src_viewmatrix && a->src_projmatrix && a->viewmatrix && a->projmatrix)
            cam = CameraCheck{a->src_viewmatrix, a->src_projmatrix, a->viewmatrix, a->projmatrix};
        launch_recolor((const unsigned char*)src_bin + BL.packed, (unsigned char*)bin + BL.packed, (uint32_t*)(bin + BL.point_list), (uint32_t)R,
                       a->colors_precomp, hdr, a->forward_only ? 1 : 0, cam, stream);
    }
    STAGE_CHECK("recolor");
    BlendParams bl;
    bl.W = W; bl.H = H; bl.gx = gx; bl.gy = gy;
    bl.recs = nullptr; bl.hdr = hdr; bl.ranges = (const uint32_t*)(img + IL.ranges); bl.tile_order = (const uint32_t*)(img + IL.tile_order);
    bl.bg = a->background; bl.out_color = a->out_color; bl.final_T = (float*)(img + IL.final_T); bl.n_contrib = (uint32_t*)(img + IL.n_contrib);
    bl.dL_dpix = nullptr; bl.gacc = nullptr; bl.tile_lanes = (const unsigned char*)(img + IL.tile_lanes);
    bl.pixstate = (float4*)(img + IL.pixstate); bl.host_counts = ctx->host_counts_dev;
    bl.packed = (const unsigned char*)bin + BL.packed;
    {
        StageScope sc(GSTAR_STAGE_BLEND_FWD, stream);
        launch_blend_fwd(bl, stream);
    }
    launch_poison(hdr, a->out_color, (size_t)3 * W * H, stream);  // no-op unless this call (or the one it re-blends) was refused on the device
    STAGE_CHECK("blend_fwd");
    ctx->calls.remember(img, cap, log_slots, (uint32_t)R, a->P, W, H);  // a re-blend can itself be re-blended
    return (int)R;
}

int gstar_raster_backward(const gstar_bwd_args* a, void* stream_)
{
    using namespace gstar;
    CallRange call_range("gstar_raster_backward");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a) return fail(GSTAR_ERR_INVALID, "null argument");
    if (a->P == 0) return 0;
    if (!a->geom_buffer || !a->image_buffer || !a->blend_grad_scratch || !a->dL_dpix)
        return fail(GSTAR_ERR_INVALID, "missing buffer");
    if (a->R > 0 && !a->binning_buffer) return fail(GSTAR_ERR_INVALID, "missing binning buffer");
    const int W = a->width, H = a->height;
    const int gx = (W + GSTAR_TILE - 1) / GSTAR_TILE, gy = (H + GSTAR_TILE - 1) / GSTAR_TILE;
    char* geom = aligned128(a->geom_buffer);
    char* img = aligned128(a->image_buffer);
    const ImgLayout IL = img_layout(W, H);

    if (a->R > 0) {
        BlendParams bl;
        bl.W = W; bl.H = H; bl.gx = gx; bl.gy = gy;
        bl.recs = (const GRec*)geom; bl.hdr = (const GHeader*)(img + IL.hdr);
        bl.ranges = (const uint32_t*)(img + IL.ranges);
        bl.tile_order = (const uint32_t*)(img + IL.tile_order);
        bl.tile_lanes = (const unsigned char*)(img + IL.tile_lanes);
        bl.packed = (const unsigned char*)aligned128(a->binning_buffer);
        bl.bg = a->background;
        bl.out_color = nullptr; bl.final_T = (float*)(img + IL.final_T); bl.n_contrib = (uint32_t*)(img + IL.n_contrib);
        bl.dL_dpix = a->dL_dpix; bl.gacc = a->blend_grad_scratch;
        bl.pixstate = (float4*)(img + IL.pixstate); bl.host_counts = nullptr;
        const bool dual = a->dL_dpix2 != nullptr;
        if (dual != (a->background2 != nullptr) || dual != (a->colors2 != nullptr))
            return fail(GSTAR_ERR_INVALID, "two-pass backward: dL_dpix2, background2 and colors2 go together");
        const int ch2 = dual ? (a->channels2 == 0 ? 3 : a->channels2) : 3;
        if (ch2 < 1 || ch2 > 4) return fail(GSTAR_ERR_INVALID, "two-pass backward: channels2 must be 1..4");
        if (dual && ch2 == 4 && !a->blend_grad_scratch2) return fail(GSTAR_ERR_INVALID, "two-pass backward with four channels: blend_grad_scratch2 [P] is missing");
        if (dual && ch2 == 4 && deterministic_mode()) return fail(GSTAR_ERR_INVALID, "the deterministic backward covers up to three channels of the second pass");
        bl.dL_dpix2 = a->dL_dpix2; bl.bg2 = a->background2; bl.colors2 = a->colors2; bl.ch2 = ch2;
        bl.gacc2 = (dual && ch2 == 4) ? a->blend_grad_scratch2 : nullptr;
        if (deterministic_mode()) {
            // test mode: one row of moments per record, then a fixed-order sum per Gaussian (k_det_reduce); needs the hit log
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            CU_OK(cudaStreamIsCapturing(stream, &cs));
            if (cs != cudaStreamCaptureStatusNone)
                return fail(GSTAR_ERR_INVALID, "the deterministic backward allocates scratch: not available while capturing a CUDA graph");
            float* rows = nullptr;
            const size_t bytes = (size_t)a->R * GSTAR_GACC * sizeof(float);
            CU_OK(cudaMallocAsync((void**)&rows, bytes, stream));
            CU_OK(cudaMemsetAsync(rows, 0, bytes, stream));
            bl.det_partial = rows; bl.aux = (const GAux*)(geom + geom_aux_offset(a->P)); bl.P = a->P;
            {
                StageScope sc(GSTAR_STAGE_BLEND_BWD, stream);
                launch_blend_bwd_gather(bl, stream);
                launch_det_reduce(bl, stream);
            }
            CU_OK(cudaFreeAsync(rows, stream));
        } else {
            // exactly one of the two does the work, decided on the device by the forward's log_overflow flag
            // (a two-pass forward always ends with a log: there is no walk-back kernel over two passes)
            StageScope sc(GSTAR_STAGE_BLEND_BWD, stream);
            launch_blend_bwd_gather(bl, stream);
            if (!dual) launch_blend_bwd(bl, stream);
            else launch_poison_no_log(bl.hdr, bl.gacc, (size_t)a->P * GSTAR_GACC, stream);
        }
        STAGE_CHECK("blend_bwd");
    }
    if (a->blend_only) return 0;  // the moments stay in the scratch for the full backward of another pass over this geometry
    PreBwdParams pb;
    pb.P = a->P; pb.D = a->D; pb.M = a->M; pb.W = W; pb.H = H;
    pb.means3D = a->means3D; pb.scales = a->scales; pb.scale_modifier = a->scale_modifier; pb.rotations = a->rotations;
    pb.shs = a->shs; pb.cov3D_precomp = a->cov3D_precomp; pb.viewmatrix = a->viewmatrix; pb.projmatrix = a->projmatrix;
    pb.campos = a->campos; pb.tan_fovx = a->tan_fovx; pb.tan_fovy = a->tan_fovy;
    pb.focal_y = H / (2.0f * a->tan_fovy);
    pb.focal_x = W / (2.0f * a->tan_fovx);
    pb.radii = a->radii; pb.recs = (const GRec*)geom; pb.aux = (const GAux*)(geom + geom_aux_offset(a->P)); pb.gacc = a->blend_grad_scratch;
    pb.dL_dmean2D = a->dL_dmean2D; pb.dL_dconic = a->dL_dconic; pb.dL_dopacity = a->dL_dopacity; pb.dL_dcolor = a->dL_dcolor;
    pb.dL_dmean3D = a->dL_dmean3D; pb.dL_dcov3D = a->dL_dcov3D; pb.dL_dsh = (a->M > 0) ? a->dL_dsh : nullptr;
    pb.dL_dscale = a->dL_dscale; pb.dL_drot = a->dL_drot;
    pb.accumulate = a->accumulate_param_grads;
    if (!pb.dL_dmean2D || !pb.dL_dopacity || !pb.dL_dmean3D) return fail(GSTAR_ERR_INVALID, "missing gradient output");
    if (a->colors_precomp && !pb.dL_dcolor) return fail(GSTAR_ERR_INVALID, "missing dL_dcolor (colors_precomp was given)");
    if (a->cov3D_precomp && !pb.dL_dcov3D) return fail(GSTAR_ERR_INVALID, "missing dL_dcov3D (cov3D_precomp was given)");
    if (a->shs && a->M > 0 && !a->dL_dsh) return fail(GSTAR_ERR_INVALID, "missing dL_dsh");
    {
        StageScope sc(GSTAR_STAGE_PREPROCESS_BWD, stream);
        launch_preprocess_bwd(pb, stream);
    }
    STAGE_CHECK("preprocess_bwd");
    return 0;
}

int gstar_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, unsigned char* present, void* stream)
{
    (void)projmatrix;  // the reference computes p_proj but only tests view z (auxiliary.h:154)
    if (P <= 0) return 0;
    if (!means3D || !viewmatrix || !present) return fail(GSTAR_ERR_INVALID, "null argument");
    gstar::launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

static int sugar_call(const gstar_sugar_args* a, void* stream, bool backward)
{
    if (!a) return fail(GSTAR_ERR_INVALID, "null argument");
    if (a->P <= 0) return 0;
    if (a->K <= 0 || a->P % a->K != 0) return fail(GSTAR_ERR_INVALID, "P must be a multiple of the Gaussians per face");
    if (!a->verts || !a->bary || !a->scales || !a->cplx || !a->dens || (!a->faces32 == !a->faces64))
        return fail(GSTAR_ERR_INVALID, "sugar prologue: missing input (exactly one of faces32 / faces64)");
    if (!backward && (!a->points || !a->scaling || !a->quats || !a->opac)) return fail(GSTAR_ERR_INVALID, "sugar prologue: missing output");
    gstar::SugarParams s;
    s.P = a->P; s.K = a->K; s.verts = a->verts; s.faces32 = a->faces32; s.faces64 = a->faces64; s.bary = a->bary; s.scales = a->scales;
    s.cplx = a->cplx; s.dens = a->dens; s.thickness = a->thickness; s.min_scale = a->min_scale; s.max_scale = a->max_scale;
    s.has_min = a->has_min; s.has_max = a->has_max; s.points = a->points; s.scaling = a->scaling; s.quats = a->quats; s.opac = a->opac;
    s.g_points = a->g_points; s.g_scaling = a->g_scaling; s.g_quats = a->g_quats; s.g_opac = a->g_opac;
    s.d_verts = a->d_verts; s.d_scales = a->d_scales; s.d_cplx = a->d_cplx; s.d_dens = a->d_dens;
    if (backward) gstar::launch_sugar_prologue_bwd(s, (cudaStream_t)stream);
    else gstar::launch_sugar_prologue_fwd(s, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}
int gstar_sugar_prologue_forward(const gstar_sugar_args* a, void* stream) { return sugar_call(a, stream, false); }
int gstar_sugar_prologue_backward(const gstar_sugar_args* a, void* stream) { return sugar_call(a, stream, true); }

int gstar_geom_unpack(const char* geom_buffer, int P, float* depths, float* means2D, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                      unsigned char* clamped, void* stream)
{
    if (P <= 0) return 0;
    if (!geom_buffer) return fail(GSTAR_ERR_INVALID, "null geometry buffer");
    const char* gb = aligned128((char*)geom_buffer);
    gstar::launch_geom_unpack((const GRec*)gb, (const GAux*)(gb + geom_aux_offset(P)), P, depths, means2D, conic_opacity, rgb, tiles_touched, clamped,
                              (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

int gstar_image_views(char* image_buffer, int width, int height, float** final_T, uint32_t** n_contrib, uint32_t** ranges)
{
    if (!image_buffer) return fail(GSTAR_ERR_INVALID, "null image buffer");
    char* img = aligned128(image_buffer);
    const ImgLayout IL = img_layout(width, height);
    if (final_T) *final_T = (float*)(img + IL.final_T);
    if (n_contrib) *n_contrib = (uint32_t*)(img + IL.n_contrib);
    if (ranges) *ranges = (uint32_t*)(img + IL.ranges);
    return 0;
}

int gstar_binning_views(char* binning_buffer, char* image_buffer, uint32_t** point_list, uint64_t* capacity)
{
    if (!binning_buffer || !image_buffer) return fail(GSTAR_ERR_INVALID, "null buffer");
    GHeader h;
    cudaError_t e = cudaMemcpy(&h, aligned128(image_buffer), sizeof(GHeader), cudaMemcpyDeviceToHost);  // test helper: synchronous
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    if (point_list) *point_list = (uint32_t*)(aligned128(binning_buffer) + h.off_point_list);
    if (capacity) *capacity = h.capacity;
    return 0;
}

int gstar_debug_header(char* image_buffer, uint32_t* words24)
{
    if (!image_buffer || !words24) return fail(GSTAR_ERR_INVALID, "null argument");
    cudaError_t e = cudaMemcpy(words24, aligned128(image_buffer), 24 * sizeof(uint32_t), cudaMemcpyDeviceToHost);  // test helper: synchronous
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

int gstar_hit_log_state(char* image_buffer, uint64_t* slots_needed, uint64_t* slots_capacity, int* in_use)
{
    if (!image_buffer) return fail(GSTAR_ERR_INVALID, "null image buffer");
    GHeader h;
    cudaError_t e = cudaMemcpy(&h, aligned128(image_buffer), sizeof(GHeader), cudaMemcpyDeviceToHost);  // test helper: synchronous
    if (e != cudaSuccess) return fail(GSTAR_ERR_CUDA, cudaGetErrorString(e));
    if (slots_needed) *slots_needed = h.log_cursor;
    if (slots_capacity) *slots_capacity = h.log_capacity;
    if (in_use) *in_use = h.log_overflow ? 0 : 1;
    return 0;
}

}  // extern "C"
