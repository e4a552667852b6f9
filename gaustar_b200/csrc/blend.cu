// blend.cu -- K6 blend_fwd and K7 blend_bwd: per-tile alpha compositing.
//
// Replace renderCUDA forward (DGR/cuda_rasterizer/forward.cu:261-374) and backward
// (DGR/cuda_rasterizer/backward.cu:399-557).  Same per-pixel arithmetic and thresholds
// (power>0 skip, alpha=min(0.99,o*exp(power)), alpha<1/255 skip, T<1e-4 stop), re-mapped for B200:
//
//  * forward: one CTA per 16x16 tile, eight warps each owning an 8x4 pixel block and streaming the tile's packed
//    records through its own shared-memory ring (TMA bulk copies, cp.async.bulk -> UBLKCP, completing on mbarriers);
//  * every record carries exact-conservative pixel bounds of {alpha >= 1/255}; per round every lane turns one record's
//    bounds into a mask of the block's pixels and a 32x32 bit transpose gives every pixel its own hit queue, so no lane
//    evaluates a record whose bounds exclude its pixel -- bit-identical, because such a pair can never pass the
//    reference's own alpha test;
//  * the forward leaves (T_i, C_i) of every blended pair in a hit log; the backward (k_blend_bwd_gather) is then
//    instance-parallel with no pixel-serial walk, no cross-lane reduction and three reductions per instance;
//  * without a hit log (switched off, or too small for the view) the backward walks the lists back to front
//    (k_blend_bwd): warp-specialised producer/consumer ring, ballot culling against the live pixels' bounding box,
//    one multi-value butterfly warp reduction and 9 RED ops per (Gaussian, warp) instead of 9 atomics per pair.
#include "gstar_common.cuh"
#include "gstar_kernels.h"
#ifdef GSTAR_FWD_DEBUG_TIME
#include <cstdio>
#endif

namespace gstar {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RS = GSTAR_REC_SMEM;
constexpr int NSTAGE = 3;                 // ring depth (3 x 12 KB stays within static shared memory)
constexpr int NCONS = 8;                  // consumer warps (8x4 pixel blocks of a 16x16 tile)
constexpr int BLEND_THREADS = (NCONS + 1) * 32;  // + one producer warp

struct WarpGeom {
    int rx0, ry0, rx1, ry1, px, py;
    int bx, by;  // the block's origin inside the tile
    bool inside;
};

__device__ __forceinline__ WarpGeom warp_geom(int tile, int gx, int W, int H)
{
    WarpGeom g;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    g.bx = (warp & 1) * 8;
    g.by = ((warp >> 1) & 3) * 4;
    g.rx0 = tx * GSTAR_TILE + g.bx;
    g.ry0 = ty * GSTAR_TILE + g.by;
    g.rx1 = g.rx0 + 7;
    g.ry1 = g.ry0 + 3;
    g.px = g.rx0 + (lane & 7);
    g.py = g.ry0 + (lane >> 3);
    g.inside = g.px < W && g.py < H;
    return g;
}

// Bounding box (tile-local pixel coordinates) of the lanes set in `live` (lane = 8*row + col of the warp's 8x4 block).
// Culling against the box of the pixels that can still change -- instead of the whole block -- is what keeps
// silhouette tiles cheap: there a handful of never-saturating background pixels would otherwise drag the whole
// warp through tens of thousands of records that only touch finished pixels.
struct LiveBox {
    int x0, x1, y0, y1;
};
__device__ __forceinline__ LiveBox live_box(unsigned live, const WarpGeom& g)
{
    const unsigned cols = (live | (live >> 8) | (live >> 16) | (live >> 24)) & 0xffu;
    LiveBox b;
    b.x0 = g.bx + __ffs(cols) - 1;
    b.x1 = g.bx + 31 - __clz(cols);
    b.y0 = g.by + ((__ffs(live) - 1) >> 3);
    b.y1 = g.by + ((31 - __clz(live)) >> 3);
    return b;
}
__device__ __forceinline__ bool bbox_overlaps_box(const unsigned char* rec, const LiveBox& b)
{
    const Foot f = unpack_foot(*reinterpret_cast<const uint32_t*>(rec + 32));
    return f.w > 0 && f.x0 <= b.x1 && f.x0 + f.w - 1 >= b.x0 && f.y0 <= b.y1 && f.y0 + f.h - 1 >= b.y0;
}

// 32-bit mask (bit = 8*row + col of the warp's 8x4 block) of the block pixels inside a record's (clipped) alpha-bounds.
__device__ __forceinline__ unsigned block_pixel_mask(const unsigned char* rec, const WarpGeom& g)
{
    const Foot f = unpack_foot(*reinterpret_cast<const uint32_t*>(rec + 32));
    const int cx0 = max(f.x0 - g.bx, 0), cx1 = min(f.x0 + f.w - 1 - g.bx, 7);
    const int cy0 = max(f.y0 - g.by, 0), cy1 = min(f.y0 + f.h - 1 - g.by, 3);
    if (cx0 > cx1 || cy0 > cy1) return 0u;  // also the empty footprint (w = h = 0)
    const unsigned cols = ((2u << cx1) - 1u) & ~((1u << cx0) - 1u);                       // 8 bits
    const unsigned rows = (0xffffffffu >> (8 * (3 - cy1))) & (0xffffffffu << (8 * cy0));  // whole bytes
    return (cols * 0x01010101u) & rows;
}

// Transpose a 32x32 bit matrix held one row per lane: afterwards lane p holds column p (bit j = row j's bit p).
__device__ __forceinline__ unsigned transpose32(unsigned x, unsigned lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
    }
    return x;
}

// forward.cu:332-335 in the operation order of the reference SASS
__device__ __forceinline__ float eval_power(float gx, float gy, float A, float B, float C, float pxf, float pyf, float& dx, float& dy)
{
    dx = __fsub_rn(gx, pxf);
    dy = __fsub_rn(gy, pyf);
    const float q = __fmaf_rn(dx, __fmul_rn(dx, A), __fmul_rn(dy, __fmul_rn(dy, C)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, B)));
}


// ---- producer: one TMA bulk copy per batch -----------------------------------------------------------------------
// The sort kernel left the tile's records contiguous and in blend order (BinParams::packed), so a batch of up to 256
// records is ONE cp.async.bulk of <= 12 KB issued by one lane (SASS: UBLKCP) that completes on the stage's `full` mbarrier.
struct Ring {
    unsigned char* rec;   // [NSTAGE][GSTAR_BATCH * RS]
    uint64_t* full;       // [NSTAGE]
    uint64_t* empty;      // [NSTAGE]
};

// first_pos = list position (relative to the tile's range start) of the batch's lowest entry, cnt = entries
__device__ __forceinline__ void produce_batch(const Ring& r, int b, int first_pos, int cnt, const unsigned char* __restrict__ tile_packed, int lane)
{
    const int s = b % NSTAGE;
    if (lane == 0) {
        if (b >= NSTAGE) mbar_wait_sleep(&r.empty[s], (uint32_t)((b / NSTAGE) - 1) & 1u);  // all 8 consumers released the stage
        mbar_arrive_expect_tx(&r.full[s], (uint32_t)cnt * RS);
        bulk_g2s(r.rec + (size_t)s * GSTAR_BATCH * RS, tile_packed + (size_t)first_pos * RS, (uint32_t)cnt * RS, &r.full[s]);
    }
    __syncwarp();
}

// Two feature passes in one blend (NP == 2; SURVEY 8f-1 "generalise the blend to more channels"): the passes share alpha and T, so
// the second one costs three more multiply-adds per blended pair instead of a whole K6 + K7.  Its per-Gaussian colours are
// gathered from BlendParams::colors2 by id (a float4 per Gaussian: one 16-byte load), its image goes to out_color2 with background
// bg2, a hit-log row grows from 16 to 32 bytes (C_rgb, T | C2_0123), and the second pass's final colour per pixel lives behind the
// log (GHeader::off_pixstate2).  The second "pass" has up to FOUR channels (BlendParams::ch2): depth as one channel plus a normal,
// say -- with RGB that is the seven-channel blend of SURVEY 8f-1.  Planes >= ch2 of bg2 / out_color2 / dL_dpix2 do not exist.
__device__ __forceinline__ float4* pixstate2_of(const BlendParams& p)
{
    return reinterpret_cast<float4*>(const_cast<unsigned char*>(p.packed) + p.hdr->off_pixstate2);
}

// ---- K6, fat tiles ---------------------------------------------------------------------------------------------------
#ifndef GSTAR_FAT_SB
#define GSTAR_FAT_SB 64
#define GSTAR_FAT_NST 2
#endif
constexpr int FAT_SB = GSTAR_FAT_SB;   // records per stage
constexpr int FAT_NW = FAT_SB / 32;
constexpr int FAT_NST = GSTAR_FAT_NST;    // stages per warp
constexpr int FAT_DYN_SMEM = NCONS * FAT_NST * FAT_SB * RS;  // 48 KB

__device__ __forceinline__ void fat_fetch(unsigned char* stage, uint64_t* bar, const unsigned char* tile_packed, int b, int n)
{
    const uint32_t bytes = (uint32_t)min(FAT_SB, n - b * FAT_SB) * RS;
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(stage, tile_packed + (size_t)b * FAT_SB * RS, bytes, bar);
}

// (round 1's forward blend, kept for the tiles whose splats are fat: footprints that cover a large part of the tile make
// lanes = pixels dense, and every warp streams the list on its own -- no CTA-wide phases)
template <int NP>
__device__ __forceinline__ void blend_fwd_fat_tile(const BlendParams& p, const int tile, unsigned char* s_dyn)
{
    __shared__ __align__(8) uint64_t s_full[NCONS][FAT_NST];
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
    const int n = (int)(re - rs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpGeom g = warp_geom(tile, p.gx, p.W, p.H);
    const float pxf = (float)g.px, pyf = (float)g.py;
    // hit log (see k_blend_bwd_gather): every blended (instance, pixel) pair records the transmittance in front of it
    // and the colour accumulated up to and including it in the instance's slot for this pixel
    const bool log_on = p.hdr->log_overflow == 0u;
    char* const hitlog = reinterpret_cast<char*>(const_cast<unsigned char*>(p.packed) + p.hdr->off_log);  // rows of NP x 16 bytes
    const int tile_x0 = (tile % p.gx) * GSTAR_TILE, tile_y0 = (tile / p.gx) * GSTAR_TILE;
    const int lx = g.px - tile_x0, ly = g.py - tile_y0;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    float D0 = 0.f, D1 = 0.f, D2 = 0.f, D3 = 0.f;  // NP == 2: the second pass's colour (same alpha, same T)
    uint32_t last = 0;
    bool done = !g.inside;

    if (n > 0 && __any_sync(FULL, !done)) {
        unsigned char* const ring = s_dyn + (size_t)warp * FAT_NST * FAT_SB * RS;
        uint64_t* const full = s_full[warp];
        const unsigned char* tile_packed = p.packed + (size_t)rs * RS;
        const int nb = (n + FAT_SB - 1) / FAT_SB;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < FAT_NST; s++) mbar_init(&full[s], 1);
            fence_mbar_init();
#pragma unroll
            for (int s = 0; s < FAT_NST; s++)
                if (s < nb) fat_fetch(ring + s * FAT_SB * RS, &full[s], tile_packed, s, n);
        }
        __syncwarp();
        int b = 0;
#pragma unroll 1
        for (; b < nb; b++) {
            const int s = b % FAT_NST;
            const unsigned char* buf = ring + s * FAT_SB * RS;
            mbar_wait(&full[s], (uint32_t)(b / FAT_NST) & 1u);
            const unsigned live = __ballot_sync(FULL, !done);
            if (live == 0) break;  // every pixel of the block is finished: this warp is done with the tile
            const int cnt = min(FAT_SB, n - b * FAT_SB);
            // per-pixel hit queue of the batch: bit k of word r = record 32 r + k may touch my pixel
            unsigned w[FAT_NW];
            int left = 0;
#pragma unroll
            for (int r = 0; r < FAT_NW; r++) {
                w[r] = 0u;
                if (r * 32 < cnt) {
                    const unsigned pm = (r * 32 + lane < cnt) ? (block_pixel_mask(buf + (r * 32 + lane) * RS, g) & live) : 0u;
                    if (__any_sync(FULL, pm != 0u)) w[r] = transpose32(pm, lane);
                }
                left += __popc(w[r]);
            }
            unsigned cw = w[0];  // word being walked, and its index
            int cr = 0;
#pragma unroll 1
            while (__any_sync(FULL, left > 0)) {
                int sl[4];
                bool ok[4];
                float al[4], cr_[4], cg[4], cbv[4];
                float e0[4], e1[4], e2[4], e3[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool have = left > 0;
                    if (have) {
                        while (cw == 0u) {  // next non-empty word (left > 0 guarantees there is one)
                            cr++;
                            unsigned nx = 0u;
#pragma unroll
                            for (int r = 1; r < FAT_NW; r++) nx = (cr == r) ? w[r] : nx;
                            cw = nx;
                        }
                        sl[u] = cr * 32 + __ffs(cw) - 1;
                        cw &= cw - 1u;
                        left--;
                    } else {
                        sl[u] = 0;
                    }
                    const unsigned char* rp = buf + sl[u] * RS;
                    const float4 q0 = *reinterpret_cast<const float4*>(rp);       // x y A B
                    const float4 q1 = *reinterpret_cast<const float4*>(rp + 16);  // C o r g
                    cbv[u] = *reinterpret_cast<const float*>(rp + 40);             // b
                    cr_[u] = q1.z; cg[u] = q1.w;
                    if constexpr (NP == 2) {
                        const float4 c2 = __ldg(reinterpret_cast<const float4*>(p.colors2) + *reinterpret_cast<const uint32_t*>(rp + 36));  // by gid
                        e0[u] = c2.x; e1[u] = c2.y; e2[u] = c2.z; e3[u] = c2.w;
                    }
                    float dx, dy;
                    const float power = eval_power(q0.x, q0.y, q0.z, q0.w, q1.x, pxf, pyf, dx, dy);
                    al[u] = fminf(0.99f, __fmul_rn(q1.y, expf(power)));
                    ok[u] = have && !(power > 0.0f) && !(al[u] < 1.0f / 255.0f);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (ok[u] && !done) {
                        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al[u]));
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            C0 = __fmaf_rn(T, __fmul_rn(al[u], cr_[u]), C0);
                            C1 = __fmaf_rn(T, __fmul_rn(al[u], cg[u]), C1);
                            C2 = __fmaf_rn(T, __fmul_rn(al[u], cbv[u]), C2);
                            if constexpr (NP == 2) {
                                D0 = __fmaf_rn(T, __fmul_rn(al[u], e0[u]), D0);
                                D1 = __fmaf_rn(T, __fmul_rn(al[u], e1[u]), D1);
                                D2 = __fmaf_rn(T, __fmul_rn(al[u], e2[u]), D2);
                                D3 = __fmaf_rn(T, __fmul_rn(al[u], e3[u]), D3);
                            }
                            if (log_on) {
                                const uint4 tail = *reinterpret_cast<const uint4*>(buf + sl[u] * RS + 32);  // foot gid b slot
                                const Foot f = unpack_foot(tail.x);
                                if constexpr (NP == 1) {
                                    GHit h;
                                    h.T = T; h.c0 = C0; h.c1 = C1; h.c2 = C2;
                                    reinterpret_cast<GHit*>(hitlog)[(size_t)tail.w + (uint32_t)((ly - f.y0) * f.w + (lx - f.x0))] = h;
                                } else {
                                    float4* row = reinterpret_cast<float4*>(hitlog + ((size_t)tail.w + (uint32_t)((ly - f.y0) * f.w + (lx - f.x0))) * (NP * sizeof(GHit)));
                                    row[0] = make_float4(C0, C1, C2, T);
                                    row[1] = make_float4(D0, D1, D2, D3);
                                }
                            }
                            T = test_T;
                            last = (uint32_t)(b * FAT_SB + sl[u] + 1);
                        }
                    }
                }
                if (done) left = 0;  // a finished pixel drops the rest of its queue
            }
            __syncwarp();
            if (lane == 0 && b + FAT_NST < nb) {
                fence_proxy_async();  // the warp's reads of this stage are ordered before the copy that overwrites it
                fat_fetch(ring + s * FAT_SB * RS, &full[s], tile_packed, b + FAT_NST, n);
            }
        }
        // leaving early: the copies already issued for the next stages must have landed before the CTA can retire
        for (int pb = b + 1; pb < min(nb, b + FAT_NST); pb++) mbar_wait(&full[pb % FAT_NST], (uint32_t)(pb / FAT_NST) & 1u);
        __syncwarp();
        if (lane == 0) {  // the persistent caller runs the next marked tile through the same barriers
#pragma unroll
            for (int s = 0; s < FAT_NST; s++) mbar_inval(&full[s]);
        }
        __syncwarp();
    }
    if (g.inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pid = (size_t)g.py * p.W + g.px;
        p.final_T[pid] = T;
        p.n_contrib[pid] = last;
        if (n > 0 && log_on) p.pixstate[pid] = make_float4(C0, C1, C2, T);  // read back by the hit-log backward only
        p.out_color[pid] = __fmaf_rn(__ldg(p.bg + 0), T, C0);  // forward.cu:372
        p.out_color[HW + pid] = __fmaf_rn(__ldg(p.bg + 1), T, C1);
        p.out_color[2 * HW + pid] = __fmaf_rn(__ldg(p.bg + 2), T, C2);
        if constexpr (NP == 2) {
            if (n > 0 && log_on) pixstate2_of(p)[pid] = make_float4(D0, D1, D2, D3);
            p.out_color2[pid] = __fmaf_rn(__ldg(p.bg2 + 0), T, D0);
            if (p.ch2 > 1) p.out_color2[HW + pid] = __fmaf_rn(__ldg(p.bg2 + 1), T, D1);
            if (p.ch2 > 2) p.out_color2[2 * HW + pid] = __fmaf_rn(__ldg(p.bg2 + 2), T, D2);
            if (p.ch2 > 3) p.out_color2[3 * HW + pid] = __fmaf_rn(__ldg(p.bg2 + 3), T, D3);
        }
    }
}


// Persistent: a few CTAs per SM stride over the non-empty tiles and take the marked ones (none on a surface scene: the
// kernel then costs a launch).
constexpr int FAT_CTAS_PER_SM = 4;
template <int NP>
__device__ __forceinline__ void blend_fwd_fat_body(const BlendParams& p, unsigned char* s_dyn)
{
    if (p.hdr->overflow || !p.tile_lanes) return;
    const uint32_t ntiles = p.hdr->cls_end[3];  // the non-empty tiles lead tile_order, longest list first
    __shared__ uint32_t s_item;
    uint32_t* const cursor = const_cast<uint32_t*>(&p.hdr->pad0[2]);  // zeroed by tile_scan (and by k_recolor for a re-blend)
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t i = s_item;
        __syncthreads();  // (everyone has read the ticket before the next one overwrites it; also fences the previous tile's rings)
        if (i >= ntiles) break;
        const int tile = (int)p.tile_order[i];
        if (p.tile_lanes[tile] & 0x80u) blend_fwd_fat_tile<NP>(p, tile, s_dyn);
    }
}
__global__ void __launch_bounds__(NCONS * 32, FAT_CTAS_PER_SM) k_blend_fwd_fat(BlendParams p)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    blend_fwd_fat_body<1>(p, s_dyn);
}
__global__ void __launch_bounds__(NCONS * 32, 3) k_blend_fwd_fat2(BlendParams p)  // two feature passes in one (BlendParams::colors2)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    blend_fwd_fat_body<2>(p, s_dyn);
}

// ---- K6: forward blend ---------------------------------------------------------------------------------------------
// One CTA (256 threads) per 16x16 tile; the tile's packed records arrive in batches of FWD_NB through a ring of TMA bulk
// copies (thread 0 issues one cp.async.bulk per batch; completion on the stage's mbarrier).  A batch goes through two
// phases (what a phase leaves for the next is double-buffered, so fast warps run ahead into the next batch's P1):
//
//  P1  alpha for every (record, pixel-of-its-footprint) PAIR, lanes = pairs.  tile_sort handed every instance one
//      hit-log slot per pixel of its clipped alpha-bounds, in list order, so the slot numbers of a 32-record group ARE its
//      pair index space.  The group's first warp marks the record starts in a bit vector; lane j of chunk c finds the
//      record of pair 32c+j by a popcount -- no search, no imbalance: a 3x3 splat costs 9 lane-evaluations, not 256
//      (reference: every pixel of the tile evaluates every record, forward.cu:329-345) and not "as many passes as the
//      busiest pixel of a warp" (round 1).  Survivors of the reference's two tests (power > 0, alpha < 1/255) store
//      alpha at their pair index and set the record's bit in their pixel's hit word (shared-memory atomicOr).
//  P2  the serial part: per pixel, its hits in list order -- load alpha, the T recurrence in the reference's order and
//      rounding (forward.cu:346-358), the hit-log row at the pair's own slot.  Consecutive records of a depth-sorted
//      list lie on an iso-depth band of the surface, so only a fraction of the tile's pixels has hits in a given batch:
//      the pixels that do are COMPACTED into a work list first (pixel state lives in shared memory), and only as many
//      warps as the list needs run P2 while the others go on to the next batch's P1.
//
// A group whose pairs do not fit the alpha buffer (fat splats: a generic 3DGS cloud) gets its hit words straight from
// the footprint rectangles (bit-matrix transposes, as round 1 did for everything) and P2 evaluates those pairs itself
// -- with footprints that cover most of a tile, lanes = pixels is dense anyway.  The arithmetic per pair is the
// reference's in both paths, so out_color / final_T / n_contrib are bit-identical.
#ifndef GSTAR_FWD_ACAP
#define GSTAR_FWD_ACAP 3584
#endif
#ifndef GSTAR_FWD_CTAS
#define GSTAR_FWD_CTAS 3
#endif
constexpr int FWD_NB = 128;                 // records per batch
constexpr int FWD_NG = FWD_NB / 32;         // 32-record groups per batch == hit words per pixel
constexpr int FWD_WARPS = 8;
#ifndef GSTAR_FWD_NST
#define GSTAR_FWD_NST 3
#endif
constexpr int FWD_NST = GSTAR_FWD_NST;      // record stages in flight
constexpr int FWD_ACAP = GSTAR_FWD_ACAP;    // alpha slots per batch
constexpr int FWD_GCAP = 2048;              // pairs of a group that goes through the alpha buffer
constexpr int FWD_GW = FWD_GCAP / 32;       // words of its record-start vector
constexpr int FWD_CTAS_PER_SM = GSTAR_FWD_CTAS;
constexpr int FWD_THREADS = FWD_WARPS * 32;
static_assert(FWD_NG == 4 && FWD_GW == 64, "P1 assumes four groups per batch and two start words per lane");

struct __align__(16) FwdHdr {   // what P2 needs of a record
    float r, g, b;
    uint32_t pk;  // footprint width | direct << 5 | (256 + pair index of the record's (virtual) pair for tile pixel (0,0), batch-relative) << 6
};
struct FwdGroup {               // P1 scratch of a 32-record group, written by the warp that drew the group's header ticket
    uint32_t words[FWD_GW];     // record starts in the group's pair space (bit q: a record's first pair is q)
    uint2 wp[FWD_GW];           // (words[c], records that start in the words before word c)
    uint2 prec[32];             // rank -> (first pair | w << 16 | record's lane << 24,  magic | x0 << 16 | y0 << 24)
    uint32_t total, light, rel0, pad;
};
struct FwdSmem {
    float alpha[2][FWD_ACAP];
    FwdHdr hdr[2][FWD_NB];
    uint32_t mask[2][FWD_NG][256];   // hit words, [word][pixel]
    uint32_t item[256];              // P2 work list: pixel ids, longest chains first
    float4 state[256];               // per pixel (C0, C1, C2, T)
    uint32_t last[256];              // per pixel last contributor
    FwdGroup grp[FWD_NG];
    uint32_t magic[17];              // floor(t / w) == t * magic[w] >> 12 for t < 256, w <= 16
    uint32_t done[8];                // pixels that reached T < 1e-4 (or lie outside the image)
    uint32_t hist[2][32];            // P2 work list, counting sort by chain length: pixels per length class
    // P1 work of a batch is drawn by whichever warp is free (the warps still busy with the previous batch's P2 draw less)
    uint32_t tick_c[2], anyfat[2];
    uint64_t full[FWD_NST];
    uint64_t hbar[2];                // the batch's group headers are complete (4 arrivals: warps 4..7)
};
#ifndef GSTAR_FWD_SMEM_PAD
#define GSTAR_FWD_SMEM_PAD 0
#endif
constexpr int FWD_DYN_SMEM = FWD_NST * FWD_NB * RS + (int)sizeof(FwdSmem) + GSTAR_FWD_SMEM_PAD;  // (the pad: occupancy experiments)
struct FwdSmem2 {                    // NP == 2: what the second feature pass adds, behind FwdSmem
    float4 hdr2[2][FWD_NB];          // the records' second colour (gathered by id in the group header)
    float4 state2[256];              // per pixel (C3, C4, C5, C6)
};
#ifndef GSTAR_FWD2_NST
#define GSTAR_FWD2_NST 2
#define GSTAR_FWD2_CTAS 3
#endif
constexpr int FWD2_NST = GSTAR_FWD2_NST;  // two stages leave room for three CTAs per SM (233 us at the headline; three stages and two CTAs: 286)
constexpr int FWD_DYN_SMEM2 = FWD2_NST * FWD_NB * RS + (int)sizeof(FwdSmem) + GSTAR_FWD_SMEM_PAD + (int)sizeof(FwdSmem2);
constexpr int FWD_UNITS = 4;  // P1 work units per group (strided chunks)
#ifndef GSTAR_FWD_U
#define GSTAR_FWD_U 3
#endif
constexpr int FWD_U = GSTAR_FWD_U;  // hits per P2 pass

__device__ __forceinline__ void fwd_fetch(unsigned char* stage, uint64_t* bar, const unsigned char* tile_packed, int b, int n)
{
    const uint32_t bytes = (uint32_t)min(FWD_NB, n - b * FWD_NB) * RS;
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(stage, tile_packed + (size_t)b * FWD_NB * RS, bytes, bar);
}

// 32-bit mask over the pixels [32 wr, 32 wr + 32) of the tile (rows 2 wr and 2 wr + 1) inside a footprint
__device__ __forceinline__ unsigned rows_pixel_mask(const Foot& f, int wr)
{
    if (f.w <= 0) return 0u;
    const unsigned cols = ((1u << f.w) - 1u) << f.x0;  // 16 bits
    const int y = 2 * wr;
    unsigned m = 0u;
    if (y >= f.y0 && y < f.y0 + f.h) m |= cols;
    if (y + 1 >= f.y0 && y + 1 < f.y0 + f.h) m |= cols << 16;
    return m;
}
// exact (float) of an integer below 2^23 without the conversion pipe: kbase = 0x4B000000 + integer
__device__ __forceinline__ float small_int_to_float(uint32_t kbase_plus_i) { return __uint_as_float(kbase_plus_i) - 8388608.0f; }
__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ uint32_t warp_ticket(uint32_t* ctr, int lane)
{
    uint32_t t = 0u;
    if (lane == 0) t = atomicAdd(ctr, 1u);
    return __shfl_sync(FULL, t, 0);
}

// P1, group header: record starts, rank table, and what P2 needs of every record of the group
template <int NP>
__device__ __forceinline__ void fwd_group_header(FwdSmem& sm, const unsigned char* stage, int buf, int grp, int cnt, uint32_t slot_b0, int lane,
                                                 FwdSmem2* sm2, const float* __restrict__ colors2)
{
    FwdGroup& G = sm.grp[grp];
    const unsigned lt_mask = (1u << lane) - 1u;
    const int ri = grp * 32 + lane;
    const bool valid = ri < cnt;
    float4 q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) q2 = *reinterpret_cast<const float4*>(stage + ri * RS + 32);  // foot gid b slot
    const Foot f = unpack_foot(__float_as_uint(q2.x));
    const uint32_t area = valid ? (uint32_t)(f.w * f.h) : 0u;
    const uint32_t slot = __float_as_uint(q2.w);
    const uint32_t slot_g0 = __shfl_sync(FULL, slot, 0);
    const uint32_t total = __reduce_add_sync(FULL, area);
    const uint32_t rel0 = slot_g0 - slot_b0;
    const bool light = total <= (uint32_t)FWD_GCAP && rel0 + total <= (uint32_t)FWD_ACAP;
    const uint32_t s0 = slot - slot_g0;  // my first pair in the group's pair space (list-order slots: binning.cu)
    if (valid) {
        const float4 q1 = *reinterpret_cast<const float4*>(stage + ri * RS + 16);  // C o r g
        FwdHdr h;
        h.r = q1.z; h.g = q1.w; h.b = q2.z;
        h.pk = (uint32_t)f.w | (light ? 0u : 32u) | ((256u + (slot - slot_b0) - (uint32_t)(f.y0 * f.w + f.x0)) << 6);
        sm.hdr[buf][ri] = h;
        if constexpr (NP == 2) {
            sm2->hdr2[buf][ri] = __ldg(reinterpret_cast<const float4*>(colors2) + __float_as_uint(q2.y));  // by Gaussian id
        }
    }
    if (lane == 0) {
        G.total = total; G.light = light ? 1u : 0u; G.rel0 = rel0;
        if (!light) sm.anyfat[buf] = 1u;
    }
    if (light && total > 0u) {
        const unsigned nonempty = __ballot_sync(FULL, area > 0u);
        G.words[lane] = 0u; G.words[lane + 32] = 0u;
        __syncwarp();
        if (area > 0u) {
            atomicOr(&G.words[s0 >> 5], 1u << (s0 & 31u));
            G.prec[__popc(nonempty & lt_mask)] =
                make_uint2(s0 | ((uint32_t)f.w << 16) | ((uint32_t)lane << 24), sm.magic[f.w] | ((uint32_t)f.x0 << 16) | ((uint32_t)f.y0 << 24));
        }
        __syncwarp();
        const uint2 ww = *reinterpret_cast<const uint2*>(&G.words[2 * lane]);  // lane owns words 2*lane, 2*lane+1
        const uint32_t c0 = (uint32_t)__popc(ww.x), c1 = (uint32_t)__popc(ww.y);
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t ex = incl - c0 - c1;
        *reinterpret_cast<uint4*>(&G.wp[2 * lane]) = make_uint4(ww.x, ex, ww.y, ex + c0);
    }
}

template <int NP>
__device__ __forceinline__ void blend_fwd_body(const BlendParams& p, unsigned char* const s_dyn)
{
    if (p.hdr->overflow) return;
    constexpr int NST = NP == 2 ? FWD2_NST : FWD_NST;
    unsigned char* const s_stage = s_dyn;  // [NST][FWD_NB * RS]
    FwdSmem& sm = *reinterpret_cast<FwdSmem*>(s_dyn + NST * FWD_NB * RS);
    FwdSmem2* const sm2 = reinterpret_cast<FwdSmem2*>(s_dyn + NST * FWD_NB * RS + sizeof(FwdSmem) + GSTAR_FWD_SMEM_PAD);  // NP == 2 only
    constexpr size_t ROW = NP * sizeof(GHit);  // bytes of a hit-log row
    const int tile = (int)p.tile_order[blockIdx.x];  // longest lists first
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
#ifdef GSTAR_FWD_SKIP_TOP
    const int n = blockIdx.x < GSTAR_FWD_SKIP_TOP ? 0 : (int)(re - rs);  // experiment: how long does the kernel take without its longest lists?
#else
    const int n = (int)(re - rs);
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_ty = tile / p.gx;
    const int tile_x0 = (tile - tile_ty * p.gx) * GSTAR_TILE, tile_y0 = tile_ty * GSTAR_TILE;
    const bool inside = tile_x0 + (tid & 15) < p.W && tile_y0 + (tid >> 4) < p.H;  // thread <-> pixel for setup and output
    // hit log (see k_blend_bwd_gather): every blended (instance, pixel) pair records the transmittance in front of it
    // and the colour accumulated up to and including it in the instance's slot for this pixel
    const bool log_on = p.hdr->log_overflow == 0u;
    char* const hitlog = reinterpret_cast<char*>(const_cast<unsigned char*>(p.packed) + p.hdr->off_log);
    if (blockIdx.x == 0 && tid == 0 && p.host_counts) {  // size the next call's log: slots this view needed
        const unsigned long long need = p.hdr->log_cursor;
        p.host_counts[4] = (uint32_t)need;
        p.host_counts[5] = (uint32_t)(need >> 32);
    }
    // tile_sort marks the tiles whose footprints average more than 24 pixels: their pairs would not fit the alpha buffer
    // anyway, and with footprints that large the pixel-parallel formulation is the dense one
    if (n > 0 && p.tile_lanes && (p.tile_lanes[tile] & 0x80u)) return;  // k_blend_fwd_fat's
    float4 fin = make_float4(0.f, 0.f, 0.f, 1.0f);  // (C, T) of my pixel
    float4 fin2 = make_float4(0.f, 0.f, 0.f, 0.f);  // NP == 2: the second pass's colour
    uint32_t fin_last = 0;
#ifdef GSTAR_FWD_DEBUG_TIME
    unsigned long long t_start = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    long long dbg_t[6] = {0, 0, 0, 0, 0, 0}, dbg_c = 0;
    int dbg_n = 0, dbg_rounds = 0, dbg_hits = 0, dbg_items = 0;
#endif

    if (n > 0) {
        const unsigned char* tile_packed = p.packed + (size_t)rs * RS;
        const int nb = (n + FWD_NB - 1) / FWD_NB;
        for (int i = tid; i < 2 * FWD_NG * 256; i += FWD_THREADS) (&sm.mask[0][0][0])[i] = 0u;
        sm.state[tid] = fin;
        if constexpr (NP == 2) sm2->state2[tid] = fin2;
        sm.last[tid] = 0u;
        if (tid < 17) sm.magic[tid] = tid ? (4096u + (uint32_t)tid - 1u) / (uint32_t)tid : 0u;
        if (tid < 64) (&sm.hist[0][0])[tid] = 0u;
        if (tid < 2) { sm.tick_c[tid] = 0u; sm.anyfat[tid] = 0u; }
        {
            const unsigned out = __ballot_sync(FULL, !inside);
            if (lane == 0) sm.done[warp] = out;
        }
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NST; s++) mbar_init(&sm.full[s], 1);
            mbar_init(&sm.hbar[0], FWD_NG);
            mbar_init(&sm.hbar[1], FWD_NG);
            fence_mbar_init();
#pragma unroll
            for (int s = 0; s < NST; s++)
                if (s < nb) fwd_fetch(s_stage + s * FWD_NB * RS, &sm.full[s], tile_packed, s, n);
        }
        int issued = min(nb, NST);  // batches whose copy has been issued (meaningful in thread 0)
        __syncthreads();
        // Warp roles: warp w takes entries [32 w, 32 w + 32) of a batch's P2 work list (longest chains first, so warp 0 mostly
        // runs the serial recurrences and draws little pair work); warps 4..7 build the group headers.  (Keeping the pair
        // work off the schedulers of the long-chain warps was tried: no gain.)
        const int p2_rank = warp;
        const int hdr_grp = warp >= FWD_NG ? warp - FWD_NG : -1;
        const uint32_t kx = 0x4B000000u + (uint32_t)tile_x0, ky = 0x4B000000u + (uint32_t)tile_y0;
        const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
        int b = 0;
#pragma unroll 1
        for (; b < nb; b++) {
            const int st = b % NST, buf = b & 1;
            const unsigned char* stage = s_stage + st * FWD_NB * RS;
#ifdef GSTAR_FWD_DEBUG_TIME
            dbg_c = clock64();
#endif
            mbar_wait(&sm.full[st], (uint32_t)(b / NST) & 1u);
#ifdef GSTAR_FWD_DEBUG_TIME
            dbg_t[5] += clock64() - dbg_c; dbg_c = clock64();
#endif
            const int cnt = min(FWD_NB, n - b * FWD_NB);
            const uint32_t slot_b0 = *reinterpret_cast<const uint32_t*>(stage + 44);  // first pair of the batch
            // ================= P1: group headers (warps 4..7), then the batch's pairs in 16 work units drawn by ticket =================
            {
                const uint32_t ng = (uint32_t)(cnt + 31) >> 5;
                if (hdr_grp >= 0) {
                    if ((uint32_t)hdr_grp < ng) fwd_group_header<NP>(sm, stage, buf, hdr_grp, cnt, slot_b0, lane, sm2, p.colors2);
                    __syncwarp();
#ifndef GSTAR_RACECHECK_BUILD
                    if (lane == 0) mbar_arrive(&sm.hbar[buf]);
#endif
                }
#ifdef GSTAR_RACECHECK_BUILD
                __syncthreads();  // (compute-sanitizer's racecheck does not model the mbarrier hand-over below: this build lets it check the rest)
#else
                mbar_wait(&sm.hbar[buf], (uint32_t)(b >> 1) & 1u);
#endif
#ifdef GSTAR_FWD_DEBUG_TIME
                dbg_t[0] += clock64() - dbg_c; dbg_c = clock64();
#endif
                float* const abase = sm.alpha[buf];
#pragma unroll 1
                for (;;) {
                    const uint32_t t = warp_ticket(&sm.tick_c[buf], lane);
                    if (t >= ng * FWD_UNITS) break;
                    const int g = (int)(t / FWD_UNITS);
                    const uint32_t j = t % FWD_UNITS;
                    const FwdGroup& G = sm.grp[g];
                    if (G.light) {
                        const uint32_t total = G.total;
                        float* const ab = abase + G.rel0;
                        uint32_t* const mk = sm.mask[buf][g];
                        const unsigned char* const gstage = stage + g * 32 * RS;
                        // chunks j, j+4, j+8, ... of the group's pair space, two per iteration while two are left (their loads and
                        // exponentials overlap), then the odd one
                        const uint32_t nch = (total + 31u) >> 5;
                        uint32_t c = j;
#pragma unroll 1
                        for (; c + FWD_UNITS < nch; c += 2u * FWD_UNITS) {
                            uint2 pr[2];
                            uint32_t qq[2];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                qq[u] = (c + (uint32_t)u * FWD_UNITS) * 32u + (uint32_t)lane;
                                const uint2 wp = G.wp[c + (uint32_t)u * FWD_UNITS];
                                pr[u] = G.prec[wp.y + (uint32_t)__popc(wp.x & le_mask) - 1u];
                            }
                            float4 q0[2];
                            float2 co[2];
                            uint32_t plx[2], ply[2];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                const unsigned char* rp = gstage + (pr[u].x >> 24) * RS;
                                q0[u] = *reinterpret_cast<const float4*>(rp);       // x y A B
                                co[u] = *reinterpret_cast<const float2*>(rp + 16);  // C o
                                const uint32_t off = qq[u] - (pr[u].x & 0xffffu), fw = __byte_perm(pr[u].x, 0u, 0x4442);
                                const uint32_t yy = (off * (pr[u].y & 0xffffu)) >> 12, xx = off - yy * fw;
                                plx[u] = __byte_perm(pr[u].y, 0u, 0x4442) + xx;
                                ply[u] = (pr[u].y >> 24) + yy;
                            }
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                float dx, dy;
                                const float power = eval_power(q0[u].x, q0[u].y, q0[u].z, q0[u].w, co[u].x, small_int_to_float(kx + plx[u]),
                                                               small_int_to_float(ky + ply[u]), dx, dy);
                                const float al = fminf(0.99f, __fmul_rn(co[u].y, expf(power)));
                                // (the last chunk of a group may be partial: only the second of the two can be it)
                                if ((u == 0 || qq[u] < total) && !(power > 0.0f) && !(al < 1.0f / 255.0f)) {  // forward.cu:336-345
                                    ab[qq[u]] = al;
                                    atomicOr(&mk[ply[u] * 16u + plx[u]], 1u << (pr[u].x >> 24));
                                }
                            }
                        }
                        if (c < nch) {
                            const uint32_t q = c * 32u + (uint32_t)lane;
                            if (q < total) {
                                const uint2 wp = G.wp[c];
                                const uint2 pr = G.prec[wp.y + (uint32_t)__popc(wp.x & le_mask) - 1u];
                                const uint32_t rl = pr.x >> 24;
                                const unsigned char* rp = gstage + rl * RS;
                                const float4 q0 = *reinterpret_cast<const float4*>(rp);       // x y A B
                                const float2 co = *reinterpret_cast<const float2*>(rp + 16);  // C o
                                const uint32_t off = q - (pr.x & 0xffffu), fw = __byte_perm(pr.x, 0u, 0x4442);
                                const uint32_t yy = (off * (pr.y & 0xffffu)) >> 12, xx = off - yy * fw;
                                const uint32_t plx = __byte_perm(pr.y, 0u, 0x4442) + xx, ply = (pr.y >> 24) + yy;
                                float dx, dy;
                                const float power = eval_power(q0.x, q0.y, q0.z, q0.w, co.x, small_int_to_float(kx + plx), small_int_to_float(ky + ply), dx, dy);
                                const float al = fminf(0.99f, __fmul_rn(co.y, expf(power)));
                                if (!(power > 0.0f) && !(al < 1.0f / 255.0f)) {  // forward.cu:336-345
                                    ab[q] = al;
                                    atomicOr(&mk[ply * 16u + plx], 1u << rl);
                                }
                            }
                        }
                    } else {
                        // fat group: hit words straight from the footprint rectangles, one transpose per 32 pixels
                        const int ri = g * 32 + lane;
                        uint32_t fw = 0u;
                        if (ri < cnt) fw = *reinterpret_cast<const uint32_t*>(stage + ri * RS + 32);
                        const Foot f = unpack_foot(fw);
#pragma unroll 1
                        for (uint32_t wr = j; wr < (uint32_t)FWD_WARPS; wr += FWD_UNITS) {
                            const unsigned pm = rows_pixel_mask(f, (int)wr);
                            if (__any_sync(FULL, pm != 0u)) sm.mask[buf][g][wr * 32u + (uint32_t)lane] = transpose32(pm, lane);
                        }
                    }
                }
            }
#ifdef GSTAR_FWD_DEBUG_TIME
            dbg_t[1] += clock64() - dbg_c; dbg_c = clock64();
#endif
            const int alive = __syncthreads_or(sm.done[lane & 7] != 0xffffffffu);  // barrier A: the batch's hit words are complete
#ifdef GSTAR_FWD_DEBUG_TIME
            dbg_t[2] += clock64() - dbg_c; dbg_c = clock64();
#endif
            if (alive == 0) break;  // every pixel of the tile is finished (forward.cu:309-311)
            if (tid == 0 && b >= 1 && b - 1 + NST < nb) {
                // every warp is past P2 of batch b-1: its stage can be refilled
                fence_proxy_async();
                fwd_fetch(s_stage + ((b - 1) % NST) * FWD_NB * RS, &sm.full[(b - 1) % NST], tile_packed, b - 1 + NST, n);
                issued = b + NST;
            }
            // ---- work list of the pixels that have hits in this batch, longest chains first (counting sort): a P2 warp then
            // holds 32 chains of about the same length ----
            const bool fat = sm.anyfat[buf] != 0u;
            {
                unsigned w[FWD_NG], any = 0u;
#pragma unroll
                for (int k = 0; k < FWD_NG; k++) {
                    w[k] = (k * 32 < cnt) ? sm.mask[buf][k][tid] : 0u;
                    any |= w[k];
                }
                const bool act = any != 0u && !((sm.done[warp] >> lane) & 1u);
                if (any != 0u && !act) {  // hits of a finished pixel: dropped here (a listed pixel's words are cleared by its P2 lane)
#pragma unroll
                    for (int k = 0; k < FWD_NG; k++) sm.mask[buf][k][tid] = 0u;
                }
                int hits = 0;
#pragma unroll
                for (int k = 0; k < FWD_NG; k++) hits += __popc(w[k]);
                const int cls = 31 - min(hits - 1, 31);  // class 0 = the longest chains
                uint32_t rank = 0;
                if (act) rank = atomicAdd(&sm.hist[buf][cls], 1u);
                if (tid < 32) sm.hist[buf ^ 1][tid] = 0u;  // every warp is past P2 of batch b-1; nobody is at batch b+1's list yet
                if (tid == 32) { sm.tick_c[buf ^ 1] = 0u; sm.anyfat[buf ^ 1] = 0u; }  // batch b+1's P1 starts behind barrier B
                __syncthreads();
                uint32_t incl = sm.hist[buf][lane];
                const uint32_t mine = incl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += v;
                }
                const uint32_t base = __shfl_sync(FULL, incl - mine, cls);
                if (act) {
                    sm.item[base + rank] = (uint32_t)tid;
                }
                __syncthreads();  // barrier B: the work list is complete
            }
            // ================= P2: one listed pixel per lane, its hits in list order =================
            int nact;
            {
                uint32_t tot = sm.hist[buf][lane];
                tot = __reduce_add_sync(FULL, tot);
                nact = (int)tot;
            }
#ifdef GSTAR_FWD_DEBUG_TIME
            dbg_t[3] += clock64() - dbg_c; dbg_c = clock64(); dbg_n += (warp * 32 < nact);
#endif
            if (p2_rank * 32 < nact) {
                const int li = p2_rank * 32 + lane;  // my entry of the work list
                const bool have = li < nact;
                const int pid = have ? (int)sm.item[li] : 0;
                unsigned w[FWD_NG];
                int left = 0;
#pragma unroll
                for (int k = 0; k < FWD_NG; k++) {
                    w[k] = (have && k * 32 < cnt) ? sm.mask[buf][k][pid] : 0u;
                    if (w[k]) sm.mask[buf][k][pid] = 0u;  // the buffer is clean again for batch b+2
                    left += __popc(w[k]);
                }
                const uint32_t lx = (uint32_t)pid & 15u, ly = (uint32_t)pid >> 4;
                float4 S = make_float4(0.f, 0.f, 0.f, 1.0f);  // C0 C1 C2 T
                float4 S2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) {  // (a lane without a list entry must not even read pixel 0's state: its owner may be writing it)
                    S = sm.state[pid];
                    if constexpr (NP == 2) S2 = sm2->state2[pid];
                }
                uint32_t lastr = 0xffffffffu, fin_flag = 0;
                unsigned cw = w[0];
                int ck = 0;
                const float* const abuf = sm.alpha[buf];  // (indices carry a +256 bias: see FwdHdr::pk)
                const FwdHdr* const hbuf = sm.hdr[buf];
                char* const hlb = NP == 1 ? reinterpret_cast<char*>(reinterpret_cast<GHit*>(hitlog) + slot_b0 - 256) : hitlog + (size_t)slot_b0 * ROW - 256 * ROW;
                int rounds = __reduce_max_sync(FULL, left);
#ifdef GSTAR_FWD_DEBUG_TIME
                dbg_rounds += rounds; dbg_hits += __reduce_add_sync(FULL, left); dbg_items += min(32, nact - p2_rank * 32);
#endif
                if (!fat) {
                    // Word by word (32 records each) with a warp-uniform trip count per word, FWD_U hits per pass, and no
                    // data-dependent branches in the body: a lane without a hit in this pass runs through predicated off.  The
                    // record headers and alphas of a pass are fetched together; only the T recurrence is a dependent chain.
                    bool live = have;
#pragma unroll
                    for (int k = 0; k < FWD_NG; k++) {
                        unsigned cw = live ? w[k] : 0u;
                        int rk = __reduce_max_sync(FULL, __popc(cw));
#pragma unroll 1
                        for (; rk > 0; rk -= FWD_U) {
                            uint32_t r[FWD_U], idx[FWD_U];
                            float4 h[FWD_U], h2[FWD_U];
                            float al[FWD_U];
                            bool v[FWD_U];
#pragma unroll
                            for (int u = 0; u < FWD_U; u++) {
                                v[u] = cw != 0u;
                                r[u] = (uint32_t)(k * 32) + (v[u] ? (uint32_t)(__ffs(cw) - 1) : 0u);
                                cw &= cw - 1u;
                                h[u] = *reinterpret_cast<const float4*>(&hbuf[r[u]]);
                                if constexpr (NP == 2) h2[u] = sm2->hdr2[buf][r[u]];
                            }
#pragma unroll
                            for (int u = 0; u < FWD_U; u++) {
                                const uint32_t pk = __float_as_uint(h[u].w);
                                idx[u] = (pk >> 6) + ly * (pk & 31u) + lx;  // 256 + the pair's index in the batch
                                al[u] = abuf[v[u] ? idx[u] - 256u : 0u];
                            }
#pragma unroll
                            for (int u = 0; u < FWD_U; u++) {
                                const float test_T = __fmul_rn(S.w, __fsub_rn(1.0f, al[u]));
                                const bool hit = v[u] && live;
                                const bool fin = hit && test_T < 0.0001f;  // forward.cu:351-355: the pair is not blended, the pixel is finished
                                const bool upd = hit && !fin;
                                live = live && !fin;
                                const float c0 = __fmaf_rn(S.w, __fmul_rn(al[u], h[u].x), S.x);
                                const float c1 = __fmaf_rn(S.w, __fmul_rn(al[u], h[u].y), S.y);
                                const float c2 = __fmaf_rn(S.w, __fmul_rn(al[u], h[u].z), S.z);
                                S.x = upd ? c0 : S.x; S.y = upd ? c1 : S.y; S.z = upd ? c2 : S.z;
                                if constexpr (NP == 2) {
                                    const float c3 = __fmaf_rn(S.w, __fmul_rn(al[u], h2[u].x), S2.x);
                                    const float c4 = __fmaf_rn(S.w, __fmul_rn(al[u], h2[u].y), S2.y);
                                    const float c5 = __fmaf_rn(S.w, __fmul_rn(al[u], h2[u].z), S2.z);
                                    const float c6 = __fmaf_rn(S.w, __fmul_rn(al[u], h2[u].w), S2.w);
                                    S2.x = upd ? c3 : S2.x; S2.y = upd ? c4 : S2.y; S2.z = upd ? c5 : S2.z; S2.w = upd ? c6 : S2.w;
                                }
                                // (C_i, T_i): colour including the pair, transmittance in front of it
                                if constexpr (NP == 1) {
                                    if (upd && log_on) *reinterpret_cast<float4*>(hlb + (size_t)idx[u] * sizeof(GHit)) = S;
                                } else {
                                    if (upd && log_on) {
                                        float4* row = reinterpret_cast<float4*>(hlb + (size_t)idx[u] * ROW);
                                        row[0] = S;
                                        row[1] = S2;
                                    }
                                }
                                S.w = upd ? test_T : S.w;
                                lastr = upd ? r[u] : lastr;
                            }
                            if (!live) cw = 0u;
                        }
                    }
                    fin_flag = (have && !live) ? 1u : 0u;
                } else {
                    // a batch with fat groups: their pairs are evaluated here, by the pixel (the reference's own loop body)
#pragma unroll 1
                    for (; rounds > 0; rounds--) {
                        if (left > 0) {
                            while (cw == 0u) {
                                ck++;
                                unsigned nx = 0u;
#pragma unroll
                                for (int k = 1; k < FWD_NG; k++) nx = (ck == k) ? w[k] : nx;
                                cw = nx;
                            }
                            const uint32_t r = (uint32_t)(ck * 32 + __ffs(cw) - 1);
                            cw &= cw - 1u;
                            left--;
                            const float4 h = *reinterpret_cast<const float4*>(&hbuf[r]);
                            const uint32_t pk = __float_as_uint(h.w);
                            const uint32_t idx = (pk >> 6) + ly * (pk & 31u) + lx;
                            float al;
                            bool ok = true;
                            if (!(pk & 32u)) {
                                al = abuf[idx - 256u];
                            } else {
                                const unsigned char* rp = stage + r * RS;
                                const float4 q0 = *reinterpret_cast<const float4*>(rp);
                                const float2 co = *reinterpret_cast<const float2*>(rp + 16);
                                float dx, dy;
                                const float power = eval_power(q0.x, q0.y, q0.z, q0.w, co.x, small_int_to_float(kx + lx), small_int_to_float(ky + ly), dx, dy);
                                al = fminf(0.99f, __fmul_rn(co.y, expf(power)));
                                ok = !(power > 0.0f) && !(al < 1.0f / 255.0f);
                            }
                            if (ok) {
                                const float test_T = __fmul_rn(S.w, __fsub_rn(1.0f, al));
                                if (test_T < 0.0001f) {
                                    fin_flag = 1u;
                                    left = 0;
                                } else {
                                    S.x = __fmaf_rn(S.w, __fmul_rn(al, h.x), S.x);
                                    S.y = __fmaf_rn(S.w, __fmul_rn(al, h.y), S.y);
                                    S.z = __fmaf_rn(S.w, __fmul_rn(al, h.z), S.z);
                                    if constexpr (NP == 2) {
                                        const float4 h2 = sm2->hdr2[buf][r];
                                        S2.x = __fmaf_rn(S.w, __fmul_rn(al, h2.x), S2.x);
                                        S2.y = __fmaf_rn(S.w, __fmul_rn(al, h2.y), S2.y);
                                        S2.z = __fmaf_rn(S.w, __fmul_rn(al, h2.z), S2.z);
                                        S2.w = __fmaf_rn(S.w, __fmul_rn(al, h2.w), S2.w);
                                    }
                                    if constexpr (NP == 1) {
                                        if (log_on) *reinterpret_cast<float4*>(reinterpret_cast<GHit*>(hlb) + idx) = S;
                                    } else {
                                        if (log_on) {
                                            float4* row = reinterpret_cast<float4*>(hlb + (size_t)idx * ROW);
                                            row[0] = S;
                                            row[1] = S2;
                                        }
                                    }
                                    S.w = test_T;
                                    lastr = r;
                                }
                            }
                        }
                    }
                }
#ifdef GSTAR_FWD_DEBUG_TIME
                dbg_t[4] += clock64() - dbg_c;
#endif
                if (have) {
                    sm.state[pid] = S;
                    if constexpr (NP == 2) sm2->state2[pid] = S2;
                    if (lastr != 0xffffffffu) sm.last[pid] = (uint32_t)(b * FWD_NB + 1) + lastr;
                    if (fin_flag) atomicOr(&sm.done[pid >> 5], 1u << (pid & 31));
                }
            }
        }
        // leaving early: the copies already issued must have landed before the CTA can retire
        if (b < nb && tid == 0)
            for (int pb = b + 1; pb < issued; pb++) mbar_wait(&sm.full[pb % NST], (uint32_t)(pb / NST) & 1u);
        __syncthreads();
        fin = sm.state[tid];
        if constexpr (NP == 2) fin2 = sm2->state2[tid];
        fin_last = sm.last[tid];
#ifdef GSTAR_FWD_DEBUG_TIME
        if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 20 || blockIdx.x == 200 || blockIdx.x == 600 || blockIdx.x == 1000)) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            printf("blk %d n=%d batches=%d start=%llu dur=%llu ns\n", blockIdx.x, n, nb, t_start % 100000000ull, t_end - t_start);
        }
        if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 600))
            printf("  blk %d warp %d cycles: stagewait %lld hdr+wait %lld units %lld barA %lld build..B %lld P2 %lld (P2 batches %d)\n", blockIdx.x, warp, dbg_t[5], dbg_t[0], dbg_t[1], dbg_t[2],
                   dbg_t[3], dbg_t[4], dbg_n);
        if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 600 || blockIdx.x == 300))
            printf("  blk %d warp %d P2: batches %d rounds %d hits %d items %d\n", blockIdx.x, warp, dbg_n, dbg_rounds, dbg_hits, dbg_items);
#endif
    }
    if (inside) {
        const int px = tile_x0 + (tid & 15), py = tile_y0 + (tid >> 4);
        const size_t HW = (size_t)p.H * p.W;
        const size_t pid = (size_t)py * p.W + px;
        p.final_T[pid] = fin.w;
        p.n_contrib[pid] = fin_last;
        if (n > 0 && log_on) p.pixstate[pid] = fin;  // (C, T): read back by the hit-log backward only
        p.out_color[pid] = __fmaf_rn(__ldg(p.bg + 0), fin.w, fin.x);  // forward.cu:372
        p.out_color[HW + pid] = __fmaf_rn(__ldg(p.bg + 1), fin.w, fin.y);
        p.out_color[2 * HW + pid] = __fmaf_rn(__ldg(p.bg + 2), fin.w, fin.z);
        if constexpr (NP == 2) {
            if (n > 0 && log_on) pixstate2_of(p)[pid] = fin2;
            p.out_color2[pid] = __fmaf_rn(__ldg(p.bg2 + 0), fin.w, fin2.x);
            if (p.ch2 > 1) p.out_color2[HW + pid] = __fmaf_rn(__ldg(p.bg2 + 1), fin.w, fin2.y);
            if (p.ch2 > 2) p.out_color2[2 * HW + pid] = __fmaf_rn(__ldg(p.bg2 + 2), fin.w, fin2.z);
            if (p.ch2 > 3) p.out_color2[3 * HW + pid] = __fmaf_rn(__ldg(p.bg2 + 3), fin.w, fin2.w);
        }
    }
}
__global__ void __launch_bounds__(FWD_THREADS, FWD_CTAS_PER_SM) k_blend_fwd(BlendParams p)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];  // (no pointer arithmetic through integers: it would demote every access to generic LD/ST/ATOM)
    blend_fwd_body<1>(p, s_dyn);
}
__global__ void __launch_bounds__(FWD_THREADS, GSTAR_FWD2_CTAS) k_blend_fwd2(BlendParams p)  // two feature passes in one (BlendParams::colors2)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    blend_fwd_body<2>(p, s_dyn);
}

// keep = own half, send = other half; after the exchange every lane holds the pair-sum of its half
__device__ __forceinline__ float bfly_pair(float a, float b, unsigned m, unsigned lane)
{
    const bool up = (lane & m) != 0;
    const float keep = up ? b : a, send = up ? a : b;
    return keep + __shfl_xor_sync(FULL, send, m);
}

__global__ void __launch_bounds__(BLEND_THREADS, 4) k_blend_bwd(BlendParams p)
{
    __shared__ __align__(128) unsigned char s_rec[NSTAGE * GSTAR_BATCH * RS];
    __shared__ __align__(8) uint64_t s_full[NSTAGE], s_empty[NSTAGE];
    __shared__ int s_kmax;
    if (p.hdr->overflow || p.hdr->log_overflow == 0u) return;  // with a hit log, k_blend_bwd_gather does the work
    const int tile = (int)p.tile_order[blockIdx.x];  // longest lists first
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
    if (re == rs) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool consumer = warp < NCONS;
    const WarpGeom g = warp_geom(tile, p.gx, p.W, p.H);
    const float pxf = (float)g.px, pyf = (float)g.py;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pid = (size_t)g.py * p.W + g.px;
    const bool inside = g.inside && consumer;

    const float T_final = inside ? p.final_T[pid] : 0.f;
    const int last_contributor = inside ? (int)p.n_contrib[pid] : 0;
    float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f;
    if (inside) { dpx0 = p.dL_dpix[pid]; dpx1 = p.dL_dpix[HW + pid]; dpx2 = p.dL_dpix[2 * HW + pid]; }
    float bg_dot = 0.f;  // backward.cu:531-533
    bg_dot += __ldg(p.bg + 0) * dpx0;
    bg_dot += __ldg(p.bg + 1) * dpx1;
    bg_dot += __ldg(p.bg + 2) * dpx2;

    if (tid == 0) {
        s_kmax = 0;
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NCONS); }
        fence_mbar_init();
    }
    __syncthreads();
    const int warp_kmax = __reduce_max_sync(FULL, last_contributor);
    if (lane == 0 && warp_kmax > 0) atomicMax(&s_kmax, warp_kmax);
    __syncthreads();
    const int total = s_kmax;  // entries [0,total) of the tile list can matter; the rest is behind every pixel's last contributor
    if (total == 0) return;
    const int nb = (total + GSTAR_BATCH - 1) / GSTAR_BATCH;
    Ring ring{s_rec, s_full, s_empty};
    if (!consumer) {
        // ===== producer warp: batch b holds list positions [total-(b*256+cnt), total-b*256) in ascending order; consumers
        // walk a batch from its last entry to its first (back to front, backward.cu:472)
        const unsigned char* tile_packed = p.packed + (size_t)rs * RS;
#pragma unroll 1
        for (int b = 0; b < nb; b++) {
            const int cnt = min(GSTAR_BATCH, total - b * GSTAR_BATCH);
            produce_batch(ring, b, total - b * GSTAR_BATCH - cnt, cnt, tile_packed, lane);
        }
        return;
    }
    // ===== consumer warps =====
    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
#pragma unroll 1
    for (int b = 0; b < nb; b++) {
        const int s = b % NSTAGE;
        const int cnt = min(GSTAR_BATCH, total - b * GSTAR_BATCH);
        const int kbase = total - 1 - b * GSTAR_BATCH;
        // Every consumer waits for every batch -- even one it will skip -- so that it can never arrive twice on the
        // same phase of a stage's `empty` barrier (the ring keeps all warps within NSTAGE batches of each other).
        mbar_wait(&s_full[s], (uint32_t)(b / NSTAGE) & 1u);
        if (kbase - (cnt - 1) < warp_kmax) {  // some entry of this batch is in front of this warp's last contributor
            // slot j of the batch (j-th entry from the back) sits at buffer position cnt-1-j
            const unsigned char* buf = s_rec + (size_t)s * GSTAR_BATCH * RS + (size_t)(cnt - 1) * RS;
#pragma unroll 1
            for (int r0 = 0; r0 < cnt; r0 += 32) {
                // pixels that can receive gradient from some entry of this round: last_contributor > smallest k of the round
                const unsigned live = __ballot_sync(FULL, last_contributor > kbase - r0 - 31);
                if (live == 0) continue;
                const LiveBox lb = live_box(live, g);
                const int j = r0 + lane;
                const bool ov = (j < cnt) && (kbase - j < warp_kmax) && bbox_overlaps_box(buf - j * RS, lb);
                unsigned m = __ballot_sync(FULL, ov);
                // Survivors are taken two at a time (A = further back, then B): the two alpha evaluations overlap, the
                // per-pixel state is advanced A then B exactly as the reference does, and both records' nine partial
                // sums go through ONE 16-value butterfly (lanes 0-15 end up with A's sums, 16-31 with B's).
#pragma unroll 1
                while (m) {
                    int slot[2];
                    bool pass[2];
                    float G[2], alpha[2], ddx[2], ddy[2];
                    float4 q0[2], q1[2];
                    float cbv[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const bool have = m != 0;
                        slot[u] = r0 + (have ? __ffs(m) - 1 : 0);
                        m &= m - 1;
                        const unsigned char* rp = buf - slot[u] * RS;
                        q0[u] = *reinterpret_cast<const float4*>(rp);
                        q1[u] = *reinterpret_cast<const float4*>(rp + 16);
                        cbv[u] = *reinterpret_cast<const float*>(rp + 40);
                        const float power = eval_power(q0[u].x, q0[u].y, q0[u].z, q0[u].w, q1[u].x, pxf, pyf, ddx[u], ddy[u]);
                        G[u] = expf(power);
                        alpha[u] = fminf(0.99f, __fmul_rn(q1[u].y, G[u]));
                        // kbase - slot == the reference's `contributor` after its decrement (backward.cu:486-488)
                        pass[u] = have && (kbase - slot[u] < last_contributor) && !(power > 0.0f) && !(alpha[u] < 1.0f / 255.0f);
                    }
                    // Per-lane partials are the raw moments of s = G*dL/dG over the pixel offsets and the colour weights
                    //   [S, S.dx, S.dy, S.dx^2, S.dx.dy, S.dy^2, w.dpx_r, w.dpx_g, w.dpx_b]   (w = alpha*T)
                    // The per-Gaussian factors (conic, opacity, 0.5*W, 0.5*H, -0.5) are linear and are applied once per
                    // Gaussian in preprocess_bwd instead of once per (Gaussian, pixel) here.
                    float v[2][9];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
#pragma unroll
                        for (int c = 0; c < 9; c++) v[u][c] = 0.f;
                        if (pass[u]) {
                            const float al = alpha[u], dx = ddx[u], dy = ddy[u];
                            const float inv1ma = __frcp_rn(__fsub_rn(1.f, al));
                            T = T * inv1ma;  // backward.cu:503 (T / (1 - alpha))
                            const float w = al * T;  // dchannel_dcolor
                            float dL_dalpha = 0.0f;
                            const float oml = 1.f - last_alpha;
                            acc0 = last_alpha * lc0 + oml * acc0; lc0 = q1[u].z; dL_dalpha += (q1[u].z - acc0) * dpx0;
                            acc1 = last_alpha * lc1 + oml * acc1; lc1 = q1[u].w; dL_dalpha += (q1[u].w - acc1) * dpx1;
                            acc2 = last_alpha * lc2 + oml * acc2; lc2 = cbv[u];  dL_dalpha += (cbv[u] - acc2) * dpx2;
                            v[u][6] = w * dpx0;
                            v[u][7] = w * dpx1;
                            v[u][8] = w * dpx2;
                            dL_dalpha *= T;
                            last_alpha = al;
                            dL_dalpha += (-T_final * inv1ma) * bg_dot;
                            const float sG = (q1[u].y * dL_dalpha) * G[u];  // dL_dG * G
                            const float sx = sG * dx, sy = sG * dy;
                            v[u][0] = sG;
                            v[u][1] = sx;
                            v[u][2] = sy;
                            v[u][3] = sx * dx;
                            v[u][4] = sx * dy;
                            v[u][5] = sy * dy;
                        }
                    }
                    const unsigned anyA = __ballot_sync(FULL, pass[0]), anyB = __ballot_sync(FULL, pass[1]);
                    if (anyA | anyB) {
                        float x[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) x[c] = bfly_pair(v[0][c], v[1][c], 16, lane);
                        const float y0 = bfly_pair(x[0], x[4], 8, lane), y1 = bfly_pair(x[1], x[5], 8, lane);
                        const float y2 = bfly_pair(x[2], x[6], 8, lane), y3 = bfly_pair(x[3], x[7], 8, lane);
                        const float z0 = bfly_pair(y0, y2, 4, lane), z1 = bfly_pair(y1, y3, 4, lane);
                        float t = bfly_pair(z0, z1, 2, lane);
                        t += __shfl_xor_sync(FULL, t, 1);
                        float o = bfly_pair(v[0][8], v[1][8], 16, lane);
                        o += __shfl_xor_sync(FULL, o, 8);
                        o += __shfl_xor_sync(FULL, o, 4);
                        o += __shfl_xor_sync(FULL, o, 2);
                        o += __shfl_xor_sync(FULL, o, 1);
                        // lane L: record = L>>4, value index = 4*bit3 + 2*bit2 + bit1
                        const int rsel = lane >> 4;
                        const bool any_mine = rsel ? (anyB != 0) : (anyA != 0);
                        if (any_mine) {
                            const uint32_t gid = *reinterpret_cast<const uint32_t*>(buf - slot[rsel] * RS + 36);
                            float* dst = p.gacc + (size_t)gid * GSTAR_GACC;
                            if ((lane & 1) == 0) atomicAdd(dst + (((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)), t);
                            else if ((lane & 15) == 1) atomicAdd(dst + 8, o);  // gacc row = the nine raw moments
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
    }
}


// ---- K7 (hit-log path): instance-parallel backward ------------------------------------------------------------------
// The reference walks every pixel's list back to front because dL/dalpha of a pair needs the transmittance in front of
// it and the colour behind it (backward.cu:472-556).  The forward blend already had both when it blended the pair and
// left them in the hit log: T_i and C_i = sum_{j<=i} c_j alpha_j T_j.  With C_fin = sum_j c_j alpha_j T_j,
//   dC/dalpha_i = T_i c_i - (C_fin - C_i + T_final bg) / (1 - alpha_i)
// which is backward.cu:505-534 in closed form, so no pixel-serial recurrence is left: one thread per INSTANCE walks the
// few pixels of its footprint in this tile, keeps the nine raw moments [S, S dx, S dy, S dx^2, S dx dy, S dy^2, w dpx_rgb]
// (S = o G dL/dalpha, w = alpha T_i; see k_blend_bwd) in registers and issues three reductions per instance -- no
// cross-lane reduction and no per-pixel atomics at all.  A pair was blended iff it lies in front of the pixel's last
// contributor and passes the reference's power / alpha tests, which are re-evaluated here with the forward's arithmetic.
constexpr int GATHER_THREADS = 256;
// NP == 2: two feature passes blended in one (k_blend_fwd2).  The pair's alpha and T are shared, so dL/dalpha is the SUM of the
// two passes' terms -- cdot and `behind` simply run over six channels -- and the three colour moments of the second pass go to
// slots 9..11 of the Gaussian's gacc row (= that pass's dL_dcolors).
template <int NP>
__device__ __forceinline__ void blend_bwd_gather_body(const BlendParams& p)
{
    __shared__ float4 s_pix[GSTAR_TILE * GSTAR_TILE];    // dL_dpix rgb, (C_fin . dpx + T_final bg . dpx) summed over the passes
    __shared__ float4 s_pix2[NP == 2 ? GSTAR_TILE * GSTAR_TILE : 1];  // dL_dpix of the second pass
    __shared__ uint32_t s_nc[GSTAR_TILE * GSTAR_TILE];  // n_contrib
    __shared__ uint32_t s_total;
    if (p.hdr->overflow || p.hdr->log_overflow) return;
    const int tile = (int)p.tile_order[blockIdx.x];
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
    if (re == rs) return;
    const int tid = threadIdx.x;
    const int tile_x0 = (tile % p.gx) * GSTAR_TILE, tile_y0 = (tile / p.gx) * GSTAR_TILE;
    {
        const int px = tile_x0 + (tid & 15), py = tile_y0 + (tid >> 4);
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), pv2 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t nc = 0;
        if (px < p.W && py < p.H) {
            const size_t HW = (size_t)p.H * p.W, pid = (size_t)py * p.W + px;
            nc = p.n_contrib[pid];
            if (nc) {
                const float d0 = p.dL_dpix[pid], d1 = p.dL_dpix[HW + pid], d2 = p.dL_dpix[2 * HW + pid];
                const float4 st = p.pixstate[pid];
                const float bg_dot = __ldg(p.bg + 0) * d0 + __ldg(p.bg + 1) * d1 + __ldg(p.bg + 2) * d2;
                pv = make_float4(d0, d1, d2, st.x * d0 + st.y * d1 + st.z * d2 + st.w * bg_dot);
                if constexpr (NP == 2) {
                    const int ch2 = p.ch2;
                    const float d3 = p.dL_dpix2[pid], d4 = ch2 > 1 ? p.dL_dpix2[HW + pid] : 0.f, d5 = ch2 > 2 ? p.dL_dpix2[2 * HW + pid] : 0.f,
                                d6 = ch2 > 3 ? p.dL_dpix2[3 * HW + pid] : 0.f;
                    const float4 st2 = pixstate2_of(p)[pid];
                    const float bg2_dot = __ldg(p.bg2 + 0) * d3 + (ch2 > 1 ? __ldg(p.bg2 + 1) * d4 : 0.f) + (ch2 > 2 ? __ldg(p.bg2 + 2) * d5 : 0.f) +
                                          (ch2 > 3 ? __ldg(p.bg2 + 3) * d6 : 0.f);
                    pv2 = make_float4(d3, d4, d5, d6);
                    pv.w += st2.x * d3 + st2.y * d4 + st2.z * d5 + st2.w * d6 + st.w * bg2_dot;
                }
            }
        }
        s_pix[tid] = pv;
        if constexpr (NP == 2) s_pix2[tid] = pv2;
        s_nc[tid] = nc;
        if (tid == 0) s_total = 0;
        __syncthreads();
        const uint32_t wmax = __reduce_max_sync(FULL, nc);
        if ((tid & 31) == 0 && wmax) atomicMax(&s_total, wmax);
        __syncthreads();
    }
    const uint32_t total = s_total;  // instances behind every pixel's last contributor were never blended
    const GHit* const hitlog = reinterpret_cast<const GHit*>(p.packed + p.hdr->off_log);  // rows of NP GHit: (C_rgb, T)[, (C2_rgb, -)]
    const float4* const tile_packed = reinterpret_cast<const float4*>(p.packed + (size_t)rs * RS);
    const float tx0f = (float)tile_x0, ty0f = (float)tile_y0;
    // Lanes per instance, chosen per tile by the sort kernel from the mean footprint area and the list length: one for
    // the tiny splats of a dense surface, 2/4/8 when footprints are large or the tile has fewer instances than threads
    // (a single lane would walk a 50-pixel footprint through 50 dependent hit-log loads).
    const int L = p.tile_lanes ? (int)(p.tile_lanes[tile] & 0x7fu) : 1;  // (bit 7: the forward's fat-tile mark)
    if (L <= 1) {
#pragma unroll 1
    for (uint32_t i = tid; i < total; i += GATHER_THREADS) {
        const float4 q2 = ldg_nc_f4(tile_packed + (size_t)i * 3 + 2);  // foot gid b slot
        const Foot f = unpack_foot(__float_as_uint(q2.x));
        if (f.w <= 0 || f.h <= 0) continue;
        const float4 q0 = ldg_nc_f4(tile_packed + (size_t)i * 3);      // x y A B
        const float4 q1 = ldg_nc_f4(tile_packed + (size_t)i * 3 + 1);  // C o r g
        const GHit* hrow = hitlog + (size_t)__float_as_uint(q2.w) * NP;
        const int area = f.w * f.h;
        // The footprint's log slots are contiguous: pull all of its lines towards L2 now, for every lane at once, so
        // that the dependent loads in the pixel loop below find them there (the loop itself exposes one miss at a time).
        for (int off = 0; off < area * NP * (int)sizeof(GHit); off += 128) prefetch_l2(reinterpret_cast<const char*>(hrow) + off);
        prefetch_l2(hrow + (area * NP - 1));  // the row need not start on a line boundary
        if (i + GATHER_THREADS < total) prefetch_l2(tile_packed + (size_t)(i + GATHER_THREADS) * 3);
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f, m5 = 0.f, m6 = 0.f, m7 = 0.f, m8 = 0.f;
        float m9 = 0.f, m10 = 0.f, m11 = 0.f, m12 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
        if constexpr (NP == 2) {
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(p.colors2) + __float_as_uint(q2.y));  // the record's second colour, by Gaussian id
            e0 = c2.x; e1 = c2.y; e2 = c2.z; e3 = c2.w;
        }
        bool any = false;
        int xx = 0, pl = f.y0 * GSTAR_TILE + f.x0;
        GHit hn = hrow[0], hn2 = hn;  // the row of the next pixel is requested one step ahead of its use (the rows of a footprint are consecutive)
        if constexpr (NP == 2) hn2 = hrow[1];
#pragma unroll 1
        for (int s = 0; s < area; s++) {
            const int cur_pl = pl;
            const GHit hcur = hn, hcur2 = hn2;
            if (s + 1 < area) {
                hn = hrow[(s + 1) * NP];
                if constexpr (NP == 2) hn2 = hrow[(s + 1) * NP + 1];
            }
            if (++xx == f.w) { xx = 0; pl += GSTAR_TILE - f.w + 1; } else pl++;
            if (i >= s_nc[cur_pl]) continue;  // behind this pixel's last contributor
            float dx, dy;
            const float power = eval_power(q0.x, q0.y, q0.z, q0.w, q1.x, tx0f + (float)(cur_pl & 15), ty0f + (float)(cur_pl >> 4), dx, dy);
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
            if (alpha < 1.0f / 255.0f) continue;
            const GHit h = hcur;
            const float4 pv = s_pix[cur_pl];
            const float w = alpha * h.T;
            float cdot = q1.z * pv.x + q1.w * pv.y + q2.z * pv.z;
            float behind = pv.w - (h.c0 * pv.x + h.c1 * pv.y + h.c2 * pv.z);
            if constexpr (NP == 2) {
                const float4 pv2 = s_pix2[cur_pl];
                cdot += e0 * pv2.x + e1 * pv2.y + e2 * pv2.z + e3 * pv2.w;
                behind -= hcur2.c0 * pv2.x + hcur2.c1 * pv2.y + hcur2.c2 * pv2.z + hcur2.T * pv2.w;  // (the row's fourth float: C6)
                m9 = fmaf(w, pv2.x, m9); m10 = fmaf(w, pv2.y, m10); m11 = fmaf(w, pv2.z, m11); m12 = fmaf(w, pv2.w, m12);
            }
            const float dL_dalpha = h.T * cdot - behind * __frcp_rn(1.0f - alpha);
            const float sG = (q1.y * dL_dalpha) * G;
            const float sx = sG * dx, sy = sG * dy;
            m0 += sG; m1 += sx; m2 += sy;
            m3 = fmaf(sx, dx, m3); m4 = fmaf(sx, dy, m4); m5 = fmaf(sy, dy, m5);
            m6 = fmaf(w, pv.x, m6); m7 = fmaf(w, pv.y, m7); m8 = fmaf(w, pv.z, m8);
            any = true;
        }
        if (any) {
            if (p.det_partial) {  // deterministic test mode: this record's own row, summed per Gaussian by k_det_reduce
                float4* row = reinterpret_cast<float4*>(p.det_partial + (size_t)(rs + i) * GSTAR_GACC);
                row[0] = make_float4(m0, m1, m2, m3); row[1] = make_float4(m4, m5, m6, m7); row[2] = make_float4(m8, m9, m10, m11);
            } else {
                float* dst = p.gacc + (size_t)__float_as_uint(q2.y) * GSTAR_GACC;
                red_add_v4(dst, m0, m1, m2, m3);
                red_add_v4(dst + 4, m4, m5, m6, m7);
                if constexpr (NP == 2) {
                    red_add_v4(dst + 8, m8, m9, m10, m11);
                    if (p.gacc2) atomicAdd(p.gacc2 + __float_as_uint(q2.y), m12);  // the fourth channel's colour moment
                } else atomicAdd(dst + 8, m8);
            }
        }
    }
    return;
    }
    // ---- L = 2, 4 or 8 lanes per instance: lane q of the group takes footprint slots q, q+L, ...; the group's partial
    // moments meet in log2(L) shuffle steps.  Uniform trip counts (the shuffles need whole warps).
    const int lshift = L == 2 ? 1 : L == 4 ? 2 : 3;
    const int q = tid & (L - 1);
    const uint32_t stride = GATHER_THREADS >> lshift;
    const uint32_t rounds = (total + stride - 1) / stride;
#pragma unroll 1
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t i = r * stride + ((uint32_t)tid >> lshift);
        constexpr int NM = NP == 2 ? 13 : 9;  // moments per record
        float m[NM];
#pragma unroll
        for (int c = 0; c < NM; c++) m[c] = 0.f;
        bool any = false;
        uint32_t gid = 0;
        if (i < total) {
            const float4 q2 = ldg_nc_f4(tile_packed + (size_t)i * 3 + 2);  // foot gid b slot
            const Foot f = unpack_foot(__float_as_uint(q2.x));
            gid = __float_as_uint(q2.y);
            if (f.w > 0 && f.h > 0) {
                const float4 q0 = ldg_nc_f4(tile_packed + (size_t)i * 3);      // x y A B
                const float4 q1 = ldg_nc_f4(tile_packed + (size_t)i * 3 + 1);  // C o r g
                const GHit* hrow = hitlog + (size_t)__float_as_uint(q2.w) * NP;
                const int area = f.w * f.h;
                for (int off = q * 128; off < area * NP * (int)sizeof(GHit); off += L * 128) prefetch_l2(reinterpret_cast<const char*>(hrow) + off);
                float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
                if constexpr (NP == 2) {
                    const float4 c2 = __ldg(reinterpret_cast<const float4*>(p.colors2) + gid);
                    e0 = c2.x; e1 = c2.y; e2 = c2.z; e3 = c2.w;
                }
                int xx = q, yy = 0;
                while (xx >= f.w) { xx -= f.w; yy++; }
                GHit hn = {0.f, 0.f, 0.f, 0.f}, hn2 = {0.f, 0.f, 0.f, 0.f};
                if (q < area) {  // the row of the lane's next pixel is requested one step ahead of its use
                    hn = hrow[q * NP];
                    if constexpr (NP == 2) hn2 = hrow[q * NP + 1];
                }
#pragma unroll 1
                for (int s = q; s < area; s += L) {
                    const GHit h = hn, h2 = hn2;
                    if (s + L < area) {
                        hn = hrow[(s + L) * NP];
                        if constexpr (NP == 2) hn2 = hrow[(s + L) * NP + 1];
                    }
                    const int pl = (f.y0 + yy) * GSTAR_TILE + f.x0 + xx;
                    xx += L;
                    while (xx >= f.w) { xx -= f.w; yy++; }
                    if (i >= s_nc[pl]) continue;  // behind this pixel's last contributor
                    float dx, dy;
                    const float power = eval_power(q0.x, q0.y, q0.z, q0.w, q1.x, tx0f + (float)(pl & 15), ty0f + (float)(pl >> 4), dx, dy);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float4 pv = s_pix[pl];
                    const float w = alpha * h.T;
                    float cdot = q1.z * pv.x + q1.w * pv.y + q2.z * pv.z;
                    float behind = pv.w - (h.c0 * pv.x + h.c1 * pv.y + h.c2 * pv.z);
                    if constexpr (NP == 2) {
                        const float4 pv2 = s_pix2[pl];
                        cdot += e0 * pv2.x + e1 * pv2.y + e2 * pv2.z + e3 * pv2.w;
                        behind -= h2.c0 * pv2.x + h2.c1 * pv2.y + h2.c2 * pv2.z + h2.T * pv2.w;
                        m[9] = fmaf(w, pv2.x, m[9]); m[10] = fmaf(w, pv2.y, m[10]); m[11] = fmaf(w, pv2.z, m[11]); m[12] = fmaf(w, pv2.w, m[12]);
                    }
                    const float dL_dalpha = h.T * cdot - behind * __frcp_rn(1.0f - alpha);
                    const float sG = (q1.y * dL_dalpha) * G;
                    const float sx = sG * dx, sy = sG * dy;
                    m[0] += sG; m[1] += sx; m[2] += sy;
                    m[3] = fmaf(sx, dx, m[3]); m[4] = fmaf(sx, dy, m[4]); m[5] = fmaf(sy, dy, m[5]);
                    m[6] = fmaf(w, pv.x, m[6]); m[7] = fmaf(w, pv.y, m[7]); m[8] = fmaf(w, pv.z, m[8]);
                    any = true;
                }
            }
        }
        const unsigned anyb = __ballot_sync(FULL, any);
        if (anyb == 0u) continue;
        for (int d = 1; d < L; d <<= 1) {
#pragma unroll
            for (int c = 0; c < NM; c++) m[c] += __shfl_xor_sync(FULL, m[c], d);
        }
        const unsigned grp = (anyb >> ((tid & 31) & ~(L - 1))) & ((1u << L) - 1u);
        if (grp != 0u && q == 0) {
            if (p.det_partial) {
                float4* row = reinterpret_cast<float4*>(p.det_partial + (size_t)(rs + i) * GSTAR_GACC);
                row[0] = make_float4(m[0], m[1], m[2], m[3]); row[1] = make_float4(m[4], m[5], m[6], m[7]);
                if constexpr (NP == 2) row[2] = make_float4(m[8], m[9], m[10], m[11]);
                else row[2] = make_float4(m[8], 0.f, 0.f, 0.f);
            } else {
                float* dst = p.gacc + (size_t)gid * GSTAR_GACC;
                red_add_v4(dst, m[0], m[1], m[2], m[3]);
                red_add_v4(dst + 4, m[4], m[5], m[6], m[7]);
                if constexpr (NP == 2) {
                    red_add_v4(dst + 8, m[8], m[9], m[10], m[11]);
                    if (p.gacc2) atomicAdd(p.gacc2 + gid, m[12]);
                } else atomicAdd(dst + 8, m[8]);
            }
        }
    }
}
#ifndef GSTAR_GATHER_CTAS
#define GSTAR_GATHER_CTAS 4
#endif
__global__ void __launch_bounds__(GATHER_THREADS, GSTAR_GATHER_CTAS) k_blend_bwd_gather(BlendParams p) { blend_bwd_gather_body<1>(p); }
#ifndef GSTAR_GATHER2_CTAS
#define GSTAR_GATHER2_CTAS 4
#endif
__global__ void __launch_bounds__(GATHER_THREADS, GSTAR_GATHER2_CTAS) k_blend_bwd_gather2(BlendParams p) { blend_bwd_gather_body<2>(p); }

// ---- deterministic test mode: per-Gaussian sum of the per-record rows in a FIXED order ----------------------------------------
// The only run-to-run freedom of the backward is the order in which the records of one Gaussian (one per tile of its rect) reach
// its gacc row through fp32 reductions (the reference has the same freedom per (Gaussian, pixel): backward.cu:523-554).  With
// gstar_set_deterministic(1) k_blend_bwd_gather leaves every record's moments in its own row and one thread per Gaussian adds the
// rows up tile by tile in row-major order of its rect.  The record of Gaussian g in a tile is found by binary search: the tile's
// list is sorted by (depth bits, id), both known per Gaussian.  A view without a usable hit log has no rows: its moments become
// NaN (never a silently non-deterministic gradient).  Test mode: R x 48 bytes of scratch and a search per (Gaussian, tile).
__global__ void __launch_bounds__(256) k_det_reduce(BlendParams p)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.P || p.hdr->overflow) return;
    const uint4 a = *reinterpret_cast<const uint4*>(p.aux + g);  // depth, rect_min, rect_max, radius
    if ((int)a.w <= 0) return;
    float* dst = p.gacc + (size_t)g * GSTAR_GACC;
    if (p.hdr->log_overflow) {
        const float nan = __int_as_float(0x7fc00000);
        for (int c = 0; c < GSTAR_GACC; c++) dst[c] = nan;
        return;
    }
    const uint64_t key = ((uint64_t)a.x << 32) | (uint32_t)g;
    const int x0 = (int)(a.y & 0xffffu), y0 = (int)(a.y >> 16), x1 = (int)(a.z & 0xffffu), y1 = (int)(a.z >> 16);
    float acc[GSTAR_GACC];  // (slots 9..11: the second pass's colour moments, zero rows unless NP == 2)
#pragma unroll
    for (int c = 0; c < GSTAR_GACC; c++) acc[c] = 0.f;
    for (int ty = y0; ty < y1; ty++)
        for (int tx = x0; tx < x1; tx++) {
            const int tile = ty * p.gx + tx;
            const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
            uint32_t lo = rs, hi = re;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                const uint32_t gm = __float_as_uint(reinterpret_cast<const float4*>(p.packed + (size_t)mid * RS)[2].y);
                const uint64_t km = ((uint64_t)__float_as_uint(p.aux[gm].depth) << 32) | gm;
                if (km < key) lo = mid + 1; else hi = mid;
            }
            if (lo >= re) continue;
            if (__float_as_uint(reinterpret_cast<const float4*>(p.packed + (size_t)lo * RS)[2].y) != (uint32_t)g) continue;
            const float4* row = reinterpret_cast<const float4*>(p.det_partial + (size_t)lo * GSTAR_GACC);
            const float4 r0 = row[0], r1 = row[1], r2 = row[2];
            acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
            acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
            acc[8] += r2.x; acc[9] += r2.y; acc[10] += r2.z; acc[11] += r2.w;
        }
#pragma unroll
    for (int c = 0; c < GSTAR_GACC; c++) dst[c] += acc[c];  // += : a scratch pre-loaded by an earlier pass over this geometry is added to (blend_only)
}

void launch_det_reduce(const BlendParams& p, cudaStream_t s)
{
    if (p.P > 0) k_det_reduce<<<(p.P + 255) / 256, 256, 0, s>>>(p);
}

// ---- shared-geometry re-blend: a second pass over the SAME Gaussians and camera with other per-Gaussian
// colours (GauSTAR renders RGB, then depth as three equal channels: refine.py:552-564 / :607-616).  Everything the first
// pass computed up to and including the sort is reused; only the colour fields of the blend-order records change.  One
// thread per instance streams its 48-byte record from the first pass's packed stream into the new one with r,g,b taken
// from colors[gid] (what preprocess_fwd stores for colors_precomp, forward.cu:241-248 skipped).  HBM: 100 B per instance
// + an L2-resident gather of 12 B per instance.
__global__ void __launch_bounds__(256) k_recolor(const float4* __restrict__ src, float4* __restrict__ dst, uint32_t* __restrict__ dst_point_list,
                                                 uint32_t R, const float* __restrict__ colors, GHeader* hdr, int disable_log, CameraCheck cam)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && disable_log) hdr->log_overflow = 1u;  // inference re-blend: no hit log (the new binning buffer has none)
    if (i == 0) hdr->pad0[2] = 0u;  // k_blend_fwd_fat's tile cursor (the header is a copy of the source call's)
    if (i == 0) { hdr->log_row_bytes = (uint32_t)sizeof(GHit); hdr->off_pixstate2 = 0ull; }  // the re-blend is a single pass whatever its source was
    if (i == 0 && cam.src_view) {
        // the caller's "same camera" claim, verified where it costs nothing: a re-blend through another camera must not
        // produce a plausible image of the wrong view.  overflow = 2 makes the blend kernels skip this call; k_poison
        // then fills the image with NaN.
        bool same = true;
        for (int k = 0; k < 16; k++)
            same = same && __float_as_uint(cam.src_view[k]) == __float_as_uint(cam.view[k]) &&
                   __float_as_uint(cam.src_proj[k]) == __float_as_uint(cam.proj[k]);
        if (!same) hdr->overflow = 2u;
    }
    // R is the host's count; for a source call that was captured into a CUDA graph that is only the capacity, and the
    // records behind the device's own count are uninitialised (their ids must not be used as indices)
    if (i >= R || i >= hdr->num_rendered) return;
    const float4 a = ldg_nc_f4(src + (size_t)i * 3);
    float4 b = ldg_nc_f4(src + (size_t)i * 3 + 1);
    float4 c = ldg_nc_f4(src + (size_t)i * 3 + 2);
    const uint32_t gid = __float_as_uint(c.y);  // c = (foot, gid, b, slot)
    const float* col = colors + (size_t)gid * 3;
    dst_point_list[i] = gid;  // the new binning buffer is a complete BinningState (the sorted value list, rasterizer_impl.h:58-68)
    b.z = __ldg(col); b.w = __ldg(col + 1); c.z = __ldg(col + 2);
    dst[(size_t)i * 3] = a; dst[(size_t)i * 3 + 1] = b; dst[(size_t)i * 3 + 2] = c;
}
__global__ void __launch_bounds__(256) k_poison(const GHeader* __restrict__ hdr, float* __restrict__ out, size_t n, uint32_t any_overflow)
{
    // overflow == 2: a refused re-blend; any_overflow: also 1 -- a replayed CUDA graph whose view needs more instances than
    // the captured capacity (nothing was blended: the image must not look like a result)
    const uint32_t ov = hdr->overflow;
    if (!(ov == 2u || (any_overflow && ov != 0u))) return;  // the usual case: one load per thread
    const float nan = __uint_as_float(0x7fc00000u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = nan;
}
void launch_poison(const GHeader* hdr, float* out, size_t n, cudaStream_t s, int any_overflow) { k_poison<<<296, 256, 0, s>>>(hdr, out, n, (uint32_t)any_overflow); }
__global__ void __launch_bounds__(256) k_poison_no_log(const GHeader* __restrict__ hdr, float* __restrict__ gacc, size_t n)
{
    if (!hdr->log_overflow || hdr->overflow) return;  // the usual case: one load per thread
    const float nan = __uint_as_float(0x7fc00000u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) gacc[i] = nan;
}
void launch_poison_no_log(const GHeader* hdr, float* gacc, size_t n, cudaStream_t s) { k_poison_no_log<<<296, 256, 0, s>>>(hdr, gacc, n); }

void launch_recolor(const unsigned char* src_packed, unsigned char* dst_packed, uint32_t* dst_point_list, uint32_t R, const float* colors, GHeader* hdr,
                    int disable_log, const CameraCheck& cam, cudaStream_t s)
{
    const uint32_t n = R > 0u ? R : 1u;  // at least one thread: the header patch
    k_recolor<<<(n + 255u) / 256u, 256, 0, s>>>(reinterpret_cast<const float4*>(src_packed), reinterpret_cast<float4*>(dst_packed), dst_point_list, R, colors,
                                                hdr, disable_log, cam);
}

static int g_blend_sms = 0;
int blend_setup()
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&g_blend_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_blend_fwd_fat, cudaFuncAttributeMaxDynamicSharedMemorySize, FAT_DYN_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_blend_fwd_fat2, cudaFuncAttributeMaxDynamicSharedMemorySize, FAT_DYN_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_blend_fwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_DYN_SMEM2);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFuncSetAttribute(k_blend_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_DYN_SMEM);
}
void launch_blend_fwd(const BlendParams& p, cudaStream_t s)
{
    const int sms = g_blend_sms > 0 ? g_blend_sms : 148;
    if (p.colors2) {  // two feature passes in one
        k_blend_fwd2<<<p.gx * p.gy, FWD_THREADS, FWD_DYN_SMEM2, s>>>(p);
        k_blend_fwd_fat2<<<min(p.gx * p.gy, sms * 3), NCONS * 32, FAT_DYN_SMEM, s>>>(p);
        return;
    }
    k_blend_fwd<<<p.gx * p.gy, FWD_THREADS, FWD_DYN_SMEM, s>>>(p);
    k_blend_fwd_fat<<<min(p.gx * p.gy, sms * FAT_CTAS_PER_SM), NCONS * 32, FAT_DYN_SMEM, s>>>(p);
}
void launch_blend_bwd(const BlendParams& p, cudaStream_t s) { k_blend_bwd<<<p.gx * p.gy, BLEND_THREADS, 0, s>>>(p); }
void launch_blend_bwd_gather(const BlendParams& p, cudaStream_t s)
{
    if (p.dL_dpix2) k_blend_bwd_gather2<<<p.gx * p.gy, GATHER_THREADS, 0, s>>>(p);
    else k_blend_bwd_gather<<<p.gx * p.gy, GATHER_THREADS, 0, s>>>(p);
}

}  // namespace gstar
