// binning.cu -- tile binning: K2 tile_scan, K3 emit, K4 tile_sort (one persistent kernel).
//
// Replaces, with an MSD formulation of the same 64-bit tile|depth key sort,
//   cub::DeviceScan::InclusiveSum      rasterizer_impl.cu:277
//   duplicateWithKeys                  rasterizer_impl.cu:70-111
//   cub::DeviceRadixSort::SortPairs    rasterizer_impl.cu:303-308
//   identifyTileRanges                 rasterizer_impl.cu:116-138
// The reference sorts (tile<<32 | depth_bits) with a stable LSD radix sort whose input is
// Gaussian-index ordered, i.e. its output order is (tile, depth_bits, gaussian index).  Here the
// high digit (tile id) is resolved first by a counting sort over tiles -- the histogram comes out of
// preprocess_fwd, this file scans it (ranges[] falls out of the scan for free) and scatters
// (depth, idx) pairs into per-tile segments -- and each segment is then sorted on the remaining
// (depth_bits, idx) 64-bit key inside shared memory by one CTA (or a 256-thread quarter of one).  The
// concatenation of the sorted segments is bit-identical to the reference's sorted list.
#include "gstar_common.cuh"
#include "gstar_kernels.h"

namespace gstar {

constexpr int SCAN_THREADS = 1024;
constexpr uint32_t SORT_CAP = 8192;                // keys (64 KB, a power of two) one CTA sorts in shared memory
constexpr int LPT_BUCKETS = 132;                   // quarter-octave size classes for the longest-first tile order
constexpr int LPT_SMALL_END = 1 + 11 * 4;          // first LPT bucket with >= 2048 instances
constexpr int LPT_MID_END = 1 + 12 * 4;            // first LPT bucket with >= 4096 instances
constexpr int LPT_MED_END = 1 + 13 * 4;            // first LPT bucket with >= 8192 instances

__device__ __forceinline__ int lpt_bucket(uint32_t c)
{
    if (c == 0) return 0;
    const int lg = 31 - __clz(c);
    const int sub = (int)((c >> max(lg - 2, 0)) & 3u);
    return 1 + lg * 4 + sub;
}

// One atomic per distinct tile per warp instead of one per lane: neighbouring Gaussians of a mesh land in
// the same tile, so plain per-lane atomics serialise on a handful of addresses.  Returns the lane's slot.
__device__ __forceinline__ uint32_t warp_aggregated_add(uint32_t* counters, uint32_t tile, bool active)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, active ? tile : 0xffffffffu);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (active && (int)lane == leader) base = atomicAdd(counters + tile, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// ---- K2: exclusive scan over the per-tile histogram; one CTA --------------------------------------
// Thread t owns tiles [t*per, (t+1)*per).  Warp-shuffle scans (two barriers, not a 20-barrier Hillis-Steele), and the
// size-class histogram / ranking for the longest-first order use ONE shared-memory atomic per distinct class per warp
// (__match_any_sync): the non-empty tiles of a view fall into a handful of classes, plain atomics serialise on them.
__device__ __forceinline__ uint32_t class_rank_add(uint32_t* hist, int bucket, bool active)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, active ? bucket : -1);
    uint32_t base = 0;
    const int leader = __ffs(peers) - 1;
    if (active && (int)lane == leader) base = atomicAdd(hist + bucket, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(SCAN_THREADS) k_tile_scan(BinParams p)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_max, s_total, s_cls[4];
    __shared__ uint32_t s_hist[LPT_BUCKETS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.num_tiles;
    // warp w owns the contiguous tiles [w*per*32, (w+1)*per*32); in row k lane l handles tile w*per*32 + 32k + l, so
    // that every load and store of a row is one coalesced 128-byte (256-byte for ranges) access -- a thread-contiguous
    // split made every warp access touch 32 sectors and the kernel LSU-bound on its single SM
    const int per = (T + SCAN_THREADS - 1) / SCAN_THREADS;
    const int w0 = warp * per * 32;
    if (tid == 0) s_max = 0;
    for (int b = tid; b < LPT_BUCKETS; b += SCAN_THREADS) s_hist[b] = 0;
    __syncthreads();
    uint32_t sum = 0, mx = 0, n_empty = 0;
    for (int k0 = 0; k0 < per; k0 += 8) {  // uniform trip count (warp-wide matches); eight independent loads in flight
        uint32_t c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = w0 + (k0 + u) * 32 + lane;
            c[u] = (k0 + u < per && t < T) ? p.tile_count[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (k0 + u >= per) break;
            const bool in = w0 + (k0 + u) * 32 + lane < T;
            sum += c[u];
            mx = max(mx, c[u]);
            // most tiles of a view are empty: counted per thread, one atomic per warp below
            if (in && c[u] == 0u) n_empty++;
            class_rank_add(s_hist, lpt_bucket(c[u]), in && c[u] != 0u);
        }
    }
    const uint32_t warp_empty = __reduce_add_sync(0xffffffffu, n_empty);
    const uint32_t warp_sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) {
        if (warp_empty) atomicAdd(&s_hist[0], warp_empty);
        s_warp[warp] = warp_sum;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) atomicMax(&s_max, mx);
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];  // SCAN_THREADS / 32 == 32 warps
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_total = wi;
        // longest-processing-time-first order of the tiles: counting sort on the size class, largest class first; empty
        // tiles (class 0) last -- the forward still has to paint them with the background
        uint32_t acc = 0;
        if (lane == 0) {
            for (int b = LPT_BUCKETS - 1; b >= 0; b--) { const uint32_t h = s_hist[b]; s_hist[b] = acc; acc += h; }
            // s_hist[b] = tiles in classes > b = first position of class b: the list-length classes of the sort kernel
            s_cls[0] = s_hist[LPT_MED_END - 1]; s_cls[1] = s_hist[LPT_MID_END - 1]; s_cls[2] = s_hist[LPT_SMALL_END - 1]; s_cls[3] = s_hist[0];
        }
    }
    __syncthreads();
    uint32_t carry = s_warp[warp];  // exclusive prefix of this warp's tiles
    uint32_t empty_pos;  // this thread's first position among the empty tiles (class 0, the tail of tile_order)
    {
        uint32_t ei = n_empty;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, ei, d);
            if (lane >= d) ei += v;
        }
        uint32_t wbase = 0;
        if (lane == 31 && ei) wbase = atomicAdd(&s_hist[0], ei);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        empty_pos = wbase + ei - n_empty;
    }
    for (int k0 = 0; k0 < per; k0 += 8) {
        uint32_t cc[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = w0 + (k0 + u) * 32 + lane;
            cc[u] = (k0 + u < per && t < T) ? p.tile_count[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (k0 + u >= per) break;
            const int t = w0 + (k0 + u) * 32 + lane;
            const bool act = t < T;
            const uint32_t c = cc[u];
            uint32_t incl = c;  // inclusive scan of the row
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t run = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t rk = class_rank_add(s_hist, lpt_bucket(c), act && c != 0u);
            const uint32_t pos = c ? rk : empty_pos;
            if (act && c == 0u) empty_pos++;
            if (act) {
                p.tile_cursor[t] = run;
                // identifyTileRanges leaves untouched tiles at the memset value (0,0): rasterizer_impl.cu:310
                *reinterpret_cast<uint2*>(p.ranges + 2 * t) = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
                p.tile_order[pos] = (uint32_t)t;
            }
        }
    }
    if (tid == 0) {
        const uint32_t total = s_total;
        const uint32_t ovf = total > p.capacity ? 1u : 0u;
        p.hdr->num_rendered = total;
        p.hdr->capacity = p.capacity;
        p.hdr->overflow = ovf;
        p.hdr->max_tile = s_max;
        p.hdr->pad0[0] = 0u; p.hdr->pad0[1] = 0u; p.hdr->pad0[2] = 0u;
        for (int c = 0; c < 4; c++) { p.hdr->cls_end[c] = s_cls[c]; p.hdr->cls_cursor[c] = 0u; }
        p.hdr->log_overflow = p.log_capacity ? 0u : 1u;
        p.hdr->log_cursor = 0ull;
        p.hdr->log_capacity = p.log_capacity;
        p.hdr->off_point_list = p.off_point_list;
        p.hdr->off_log = p.off_log;
        p.hdr->off_pixstate2 = p.off_pixstate2;
        p.hdr->log_row_bytes = p.log_row_bytes;
        p.hdr->pad1 = 0u;
        if (p.host_counts) {  // zero-copy write of R to pinned host memory: no separate D2H memcpy
            p.host_counts[1] = ovf;
            p.host_counts[2] = s_max;
            __threadfence_system();
            p.host_counts[0] = total;
        }
    }
}

// ---- K3: scatter (idx, depth) into the tile segments ---------------------------------------------
__global__ void __launch_bounds__(256) k_emit(BinParams p)
{
    if (p.hdr->overflow) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    uint32_t minx = 0, miny = 0, w = 0, h = 0, dbits = 0;
    if (idx < p.P) {
        // depth, rect_min, rect_max, radius
        const uint4 q = *reinterpret_cast<const uint4*>(p.aux + idx);
        vis = (int)q.w > 0;
        dbits = q.x;
        minx = q.y & 0xffffu; miny = q.y >> 16;
        w = (q.z & 0xffffu) - minx; h = (q.z >> 16) - miny;
    }
    uint32_t* cur = p.tile_cursor;
    uint2* ent = p.entries;
    const int gx = p.gx;
    const unsigned lane = threadIdx.x & 31;
    // small rects (<= 4 tiles): warp-aggregated slot allocation, one atomic per distinct tile per warp
    const bool small = vis && w * h <= 4;
    const uint32_t nsmall = small ? w * h : 0u;
    const uint32_t rounds = __reduce_max_sync(0xffffffffu, nsmall);
    uint32_t tx = minx, ty = miny;  // walks the rect row by row (no integer division in the loop)
    for (uint32_t t = 0; t < rounds; t++) {
        const bool act = t < nsmall;
        const uint32_t tile = act ? ty * gx + tx : 0u;
        if (++tx == minx + w) { tx = minx; ty++; }
        const uint32_t slot = warp_aggregated_add(cur, tile, act);
        if (act) ent[slot] = make_uint2((uint32_t)idx, dbits);
    }
    // large rects: the whole warp drains one Gaussian's rect at a time (warp-cooperative emission)
    unsigned big = __ballot_sync(0xffffffffu, vis && w * h > 4);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bx = __shfl_sync(0xffffffffu, minx, src), by = __shfl_sync(0xffffffffu, miny, src);
        const uint32_t bw = __shfl_sync(0xffffffffu, w, src), ba = bw * __shfl_sync(0xffffffffu, h, src);
        const uint32_t bd = __shfl_sync(0xffffffffu, dbits, src);
        const uint32_t bi = (uint32_t)__shfl_sync(0xffffffffu, idx, src);
        for (uint32_t t = lane; t < ba; t += 32) {
            const uint32_t tile = (by + t / bw) * gx + (bx + t % bw);
            const uint32_t slot = atomicAdd(cur + tile, 1u);
            ent[slot] = make_uint2(bi, bd);
        }
    }
}

// ---- K4: per-tile sort of (depth_bits<<32 | idx) ----------------------------------------------------
// Bitonic network in its "flip/disperse" form: every compare-exchange is ascending, so the virtual +inf
// padding up to the next power of two never moves and out-of-range partners are simply skipped.
// Stages whose partners are < 8 apart run on 8 keys held in registers (levels 2,4,8 completely, and the
// j = 4,2,1 tail of every later level), which removes about a third of the shared-memory round trips and
// barriers of the textbook network.
__device__ __forceinline__ void cmpxchg(uint64_t* a, uint32_t i, uint32_t l)
{
    const uint64_t x = a[i], y = a[l];
    if (x > y) { a[i] = y; a[l] = x; }
}
__device__ __forceinline__ void cx(uint64_t& x, uint64_t& y)
{
    const uint64_t lo = x < y ? x : y, hi = x < y ? y : x;
    x = lo; y = hi;
}
__device__ __forceinline__ void load8(const uint64_t* a, uint32_t base, uint32_t n, uint64_t v[8])
{
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = (base + e < n) ? a[base + e] : ~0ull;
}
__device__ __forceinline__ void store8(uint64_t* a, uint32_t base, uint32_t n, const uint64_t v[8])
{
#pragma unroll
    for (int e = 0; e < 8; e++)
        if (base + e < n) a[base + e] = v[e];
}
__device__ __forceinline__ void tail421(uint64_t v[8])
{
    cx(v[0], v[4]); cx(v[1], v[5]); cx(v[2], v[6]); cx(v[3], v[7]);
    cx(v[0], v[2]); cx(v[1], v[3]); cx(v[4], v[6]); cx(v[5], v[7]);
    cx(v[0], v[1]); cx(v[2], v[3]); cx(v[4], v[5]); cx(v[6], v[7]);
}
__device__ __forceinline__ void sort8(uint64_t v[8])
{
    cx(v[0], v[1]); cx(v[2], v[3]); cx(v[4], v[5]); cx(v[6], v[7]);                      // k=2
    cx(v[0], v[3]); cx(v[1], v[2]); cx(v[4], v[7]); cx(v[5], v[6]);                      // k=4 flip
    cx(v[0], v[1]); cx(v[2], v[3]); cx(v[4], v[5]); cx(v[6], v[7]);                      //     j=1
    cx(v[0], v[7]); cx(v[1], v[6]); cx(v[2], v[5]); cx(v[3], v[4]);                      // k=8 flip
    cx(v[0], v[2]); cx(v[1], v[3]); cx(v[4], v[6]); cx(v[5], v[7]);                      //     j=2
    cx(v[0], v[1]); cx(v[2], v[3]); cx(v[4], v[5]); cx(v[6], v[7]);                      //     j=1
}

// A "group" is the set of threads that sorts one tile: the whole CTA, or a 256-thread quarter of it (named barrier).
struct Grp {
    uint32_t tid, nt;  // thread index inside the group, group size (a multiple of 32)
    int bar;           // 0: __syncthreads(); 1..15: bar.sync <bar>, nt
};
__device__ __forceinline__ void gsync(const Grp& g)
{
    if (g.bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.nt) : "memory");
}
// per-group scratch in shared memory
struct GrpSmem {
    uint32_t coarse[256], cbase[256];
    uint32_t w[33];            // warp partials of the group scans
    uint32_t item, dmin, dmax, flag;
    unsigned long long base;   // hit-log base of the tile
};

// disperse stages j = jstart .. 8 on the array, then the register tail (4,2,1); npad = padded length (multiple of 8)
__device__ __forceinline__ void disperse_and_tail(uint64_t* a, uint32_t n, uint32_t npad, uint32_t jstart, const Grp& g)
{
    const uint32_t pairs = npad >> 1;
    for (uint32_t j = jstart; j >= 8; j >>= 1) {
        const uint32_t jshift = 31 - __clz(j);
        for (uint32_t t = g.tid; t < pairs; t += g.nt) {
            const uint32_t i = ((t >> jshift) << (jshift + 1)) + (t & (j - 1)), l = i + j;
            if (l < n) cmpxchg(a, i, l);
        }
        gsync(g);
    }
    for (uint32_t gidx = g.tid; gidx < (npad >> 3); gidx += g.nt) {
        const uint32_t base = gidx << 3;
        if (base + 1 < n) {
            uint64_t v[8];
            load8(a, base, n, v);
            tail421(v);
            store8(a, base, n, v);
        }
    }
    gsync(g);
}

__device__ __forceinline__ void bitonic_sort(uint64_t* a, uint32_t n, const Grp& g)
{
    if (n < 2) return;
    uint32_t npad = 8;
    while (npad < n) npad <<= 1;
    for (uint32_t gidx = g.tid; gidx < (npad >> 3); gidx += g.nt) {  // levels 2,4,8 in registers
        const uint32_t base = gidx << 3;
        if (base + 1 < n) {
            uint64_t v[8];
            load8(a, base, n, v);
            sort8(v);
            store8(a, base, n, v);
        }
    }
    gsync(g);
    const uint32_t pairs = npad >> 1;
    for (uint32_t k = 16; k <= npad; k <<= 1) {
        const uint32_t half = k >> 1, hshift = 31 - __clz(half);
        for (uint32_t t = g.tid; t < pairs; t += g.nt) {  // flip stage
            const uint32_t blk = t >> hshift, off = t & (half - 1);
            const uint32_t i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < n) cmpxchg(a, i, l);
        }
        gsync(g);
        disperse_and_tail(a, n, npad, k >> 2, g);
    }
}

__device__ __forceinline__ uint32_t foot_area(uint32_t bbx, uint32_t bby, int tx0, int ty0, int limx, int limy)
{
    const Foot f = clip_foot(bbx, bby, tx0, ty0, limx, limy);
    return (f.w > 0 && f.h > 0) ? (uint32_t)(f.w * f.h) : 0u;
}

// Exclusive scan of one value per thread over the group.
__device__ __forceinline__ uint32_t group_excl_scan(uint32_t v, const Grp& g, uint32_t* s_w, uint32_t* total = nullptr)
{
    const uint32_t lane = g.tid & 31, warp = g.tid >> 5, nwarps = g.nt >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += u;
    }
    if (lane == 31) s_w[warp] = incl;
    gsync(g);
    if (warp == 0) {
        const uint32_t w = lane < nwarps ? s_w[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wi, d);
            if ((int)lane >= d) wi += u;
        }
        if (lane < nwarps) s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    gsync(g);
    const uint32_t r = s_w[warp] + incl - v;
    if (total) *total = s_w[32];
    gsync(g);  // s_w may be reused right away
    return r;
}

// Write the sorted list of one tile: point_list (the reference's sorted value list) and the tile-contiguous packed
// record stream the blend kernels read with one TMA bulk copy per batch: the first 44 bytes of the Gaussian's GRec
// followed by the instance's first HIT-LOG slot.  Every instance owns one 16-byte slot per pixel of its alpha-bounds
// inside this tile (clip_foot).  The slots of a tile are contiguous and handed out IN LIST ORDER (instance i's slots
// directly follow instance i-1's): the slot numbers double as the tile's (instance, pixel) PAIR index, which is what
// lets blend_fwd evaluate alpha with lanes = pairs (see blend.cu).  Tiles take their block from a global cursor; it
// keeps counting when the log is too small (or disabled) so that the host learns the size this view needs.  Random
// 48-byte reads hit the L2-resident GRec array; writes are contiguous.  Every gather is issued for U instances at
// once: the loops are latency-bound (one L2 round trip per step).
// s_area: shared-memory scratch of `scap` words (the areas of a chunk of the list, then their exclusive prefix).
template <int U>
__device__ __forceinline__ void write_sorted(const BinParams& p, uint32_t tile, uint32_t start, const uint64_t* keys, uint32_t n, const Grp& g,
                                             GrpSmem* sm, uint32_t* s_area, uint32_t scap)
{
    const float4* recs = reinterpret_cast<const float4*>(p.recs);
    float4* out = reinterpret_cast<float4*>(p.packed + (size_t)start * GSTAR_REC_SMEM);
    const int tx0 = (int)(tile % (uint32_t)p.gx) * GSTAR_TILE, ty0 = (int)(tile / (uint32_t)p.gx) * GSTAR_TILE;
    const int limx = min(GSTAR_TILE - 1, p.W - 1 - tx0), limy = min(GSTAR_TILE - 1, p.H - 1 - ty0);
    const unsigned char* recb = reinterpret_cast<const unsigned char*>(p.recs);
    const uint32_t tid = g.tid, nt = g.nt;
    uint32_t tile_total = 0;
    if (n > scap) {  // a list longer than the scratch (rare): its total first, so that the tile still gets ONE block of slots
        uint32_t mine = 0;
        for (uint32_t i = tid; i < n; i += nt) {
            const uint2 bb = *reinterpret_cast<const uint2*>(recb + (size_t)(uint32_t)keys[i] * GSTAR_REC_BYTES + 32);
            mine += foot_area(bb.x, bb.y, tx0, ty0, limx, limy);
        }
        group_excl_scan(mine, g, sm->w, &tile_total);
    }
    uint32_t done_slots = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += scap) {
        const uint32_t m = min(scap, n - c0);
        for (uint32_t i0 = tid; i0 < m; i0 += U * nt) {
            uint2 bb[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * nt;
                bb[u] = make_uint2(1u, 0u);  // empty box
                if (i < m) bb[u] = *reinterpret_cast<const uint2*>(recb + (size_t)(uint32_t)keys[c0 + i] * GSTAR_REC_BYTES + 32);
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (i0 + u * nt < m) s_area[i0 + u * nt] = foot_area(bb[u].x, bb[u].y, tx0, ty0, limx, limy);
        }
        gsync(g);
        // exclusive prefix of the areas in list order: thread t owns the entries [t*per, (t+1)*per)
        const uint32_t per = (m + nt - 1) / nt;
        uint32_t mine = 0;
        for (uint32_t e = 0; e < per; e++) {
            const uint32_t i = tid * per + e;
            if (i < m) mine += s_area[i];
        }
        uint32_t chunk_total;
        uint32_t run = group_excl_scan(mine, g, sm->w, &chunk_total);
        for (uint32_t e = 0; e < per; e++) {
            const uint32_t i = tid * per + e;
            if (i < m) { const uint32_t a = s_area[i]; s_area[i] = run; run += a; }
        }
        if (c0 == 0) {
            if (n <= scap) tile_total = chunk_total;
            if (tid == 0) {
                const unsigned long long base = atomicAdd(&p.hdr->log_cursor, (unsigned long long)tile_total);
                if (base + tile_total > p.log_capacity) p.hdr->log_overflow = 1u;
                sm->base = base;
                // lanes per instance for the gather backward: about a dozen footprint pixels per lane, and more lanes when the
                // tile has fewer instances than the gather CTA has threads
                const uint32_t mean = tile_total / max(n, 1u);
                uint32_t lanes = mean <= 14u ? 1u : mean <= 28u ? 2u : mean <= 64u ? 4u : 8u;
                while (lanes < 8u && n * lanes < 256u && mean > 3u * lanes) lanes <<= 1;
                // bit 7: a fat tile (mean footprint above 24 pixels) -- blend_fwd takes the pixel-parallel path for it
                p.tile_lanes[tile] = (unsigned char)(lanes | (mean > 24u ? 0x80u : 0u));
            }
        }
        gsync(g);
        const uint32_t slot0 = (uint32_t)sm->base + done_slots;
        for (uint32_t i0 = tid; i0 < m; i0 += U * nt) {
            uint32_t id[U];
            float4 a[U], b[U], c[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * nt;
                id[u] = i < m ? (uint32_t)keys[c0 + i] : 0u;
                if (i < m) {
                    a[u] = recs[(size_t)id[u] * 3]; b[u] = recs[(size_t)id[u] * 3 + 1]; c[u] = recs[(size_t)id[u] * 3 + 2];
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * nt;
                if (i < m) {
                    p.point_list[start + c0 + i] = id[u];
                    const Foot ft = clip_foot(__float_as_uint(c[u].x), __float_as_uint(c[u].y), tx0, ty0, limx, limy);
                    c[u].x = __uint_as_float(pack_foot(ft));  // the blend kernels get the clipped footprint ready-made
                    c[u].y = __uint_as_float(id[u]);
                    c[u].w = __uint_as_float(slot0 + s_area[i]);
                    out[(size_t)(c0 + i) * 3] = a[u]; out[(size_t)(c0 + i) * 3 + 1] = b[u]; out[(size_t)(c0 + i) * 3 + 2] = c[u];
                }
            }
        }
        done_slots += chunk_total;
        gsync(g);  // s_area is rewritten by the next chunk
    }
}

// Next tile of list-length class `cls` for this group (tiles were ordered longest first by tile_scan).
__device__ __forceinline__ bool next_tile(const BinParams& p, int cls, const Grp& g, GrpSmem* sm, uint32_t& tile, uint32_t& start, uint32_t& n)
{
    gsync(g);  // everyone is done with the previous tile (shared memory, sm->item)
    if (g.tid == 0) {
        const uint32_t lo = cls ? p.hdr->cls_end[cls - 1] : 0u;
        sm->item = lo + atomicAdd(&p.hdr->cls_cursor[cls], 1u);
    }
    gsync(g);
    const uint32_t item = sm->item;
    if (item >= p.hdr->cls_end[cls]) return false;
    tile = p.tile_order[item];
    start = p.ranges[2 * tile];
    n = p.ranges[2 * tile + 1] - start;
    return true;
}

// Class 0 (lists of >= 8192 instances; rare): bitonic network, chunk-wise in shared memory with the chunk-crossing stages
// on the L2-resident keys.  Also the network the bucket sort falls back to.
__device__ __forceinline__ void network_sort_tile(const BinParams& p, const Grp& g, GrpSmem* sm, uint64_t* s_keys, uint32_t tile, uint32_t start,
                                                  uint32_t n)
{
    const uint32_t tid = g.tid, nt = g.nt;
    uint64_t* gk = reinterpret_cast<uint64_t*>(p.entries) + start;
    constexpr uint32_t CH = SORT_CAP;
    if (n <= CH) {
        for (uint32_t i = tid; i < n; i += nt) s_keys[i] = gk[i];
        gsync(g);
        bitonic_sort(s_keys, n, g);
        write_sorted<2>(p, tile, start, s_keys, n, g, sm, reinterpret_cast<uint32_t*>(s_keys + CH), CH);  // the fine-bucket region behind the keys
        return;
    }
    uint32_t npad = CH;
    while (npad < n) npad <<= 1;
    const uint32_t nchunks = (n + CH - 1) / CH;
    for (uint32_t c = 0; c < nchunks; c++) {  // phase 0: every chunk fully sorted in shared memory
        const uint32_t base = c * CH, m = min(CH, n - base);
        for (uint32_t i = tid; i < m; i += nt) s_keys[i] = gk[base + i];
        gsync(g);
        bitonic_sort(s_keys, m, g);
        for (uint32_t i = tid; i < m; i += nt) gk[base + i] = s_keys[i];
        gsync(g);
    }
    for (uint32_t k = 2 * CH; k <= npad; k <<= 1) {  // merge levels
        const uint32_t half = k >> 1, hshift = 31 - __clz(half);
        for (uint32_t t = tid; t < (npad >> 1); t += nt) {  // flip stage (global)
            const uint32_t blk = t >> hshift, off = t & (half - 1);
            const uint32_t i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < n) cmpxchg(gk, i, l);
        }
        gsync(g);
        for (uint32_t j = k >> 2; j >= CH; j >>= 1) {  // disperse stages that still cross chunks (global)
            const uint32_t jshift = 31 - __clz(j);
            for (uint32_t t = tid; t < (npad >> 1); t += nt) {
                const uint32_t i = ((t >> jshift) << (jshift + 1)) + (t & (j - 1)), l = i + j;
                if (l < n) cmpxchg(gk, i, l);
            }
            gsync(g);
        }
        for (uint32_t c = 0; c < nchunks; c++) {  // stages j = CH/2 .. 1 are chunk-local (shared memory)
            const uint32_t base = c * CH, m = min(CH, n - base);
            for (uint32_t i = tid; i < m; i += nt) s_keys[i] = gk[base + i];
            gsync(g);
            disperse_and_tail(s_keys, m, CH, CH >> 1, g);
            for (uint32_t i = tid; i < m; i += nt) gk[base + i] = s_keys[i];
            gsync(g);
        }
    }
    write_sorted<2>(p, tile, start, gk, n, g, sm, reinterpret_cast<uint32_t*>(s_keys), 2 * CH);
}

// Classes 1 and 2 (lists shorter than CAP): two-level adaptive bucket sort, O(n) and ~a dozen barriers instead of the
// ~45 shared-memory passes of the network.  Keys stay in registers (CAP/NT per thread).
//   1. 256 uniform coarse bins over the tile's depth range [dmin, dmax] -> histogram.
//   2. every coarse bin is split into as many fine buckets as it holds keys (so clusters -- the front and the back
//      layer of a surface -- get resolution where the keys are): n fine buckets, ~1 key each; counting sort into them.
//   3. each fine bucket is finished by an insertion sort on the full (depth, index) key by one thread.
// Both bin functions are monotone in the depth bits, equal depths share a bucket, and step 3 orders on the whole
// 64-bit key, so the result is the reference's (tile, depth, index) order exactly.  A pathological fine bucket
// (> MAX_FINE keys, e.g. thousands of identical depths) sends the tile through the bitonic network instead.
template <uint32_t CAP, int NT>
__device__ __forceinline__ void bucket_sort_tile(const BinParams& p, const Grp& g, GrpSmem* sm, uint64_t* s_out, uint32_t* s_fine, uint32_t tile,
                                                 uint32_t start, uint32_t n)
{
    constexpr int KPT = CAP / NT;
    constexpr uint32_t MAX_FINE = 48;
    const uint32_t tid = g.tid, lane = tid & 31;
    const uint64_t* gk = reinterpret_cast<const uint64_t*>(p.entries) + start;
    uint64_t k[KPT];
    uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
    for (int e = 0; e < KPT; e++) {
        const uint32_t i = tid + e * NT;
        k[e] = i < n ? gk[i] : ~0ull;
        if (i < n) { dmin = min(dmin, (uint32_t)(k[e] >> 32)); dmax = max(dmax, (uint32_t)(k[e] >> 32)); }
    }
    if (tid == 0) { sm->dmin = 0xffffffffu; sm->dmax = 0u; sm->flag = 0u; }
    if (tid < 256) sm->coarse[tid] = 0u;
#pragma unroll
    for (int e = 0; e <= KPT; e++) {
        const uint32_t i = tid + e * NT;
        if (i < CAP + KPT) s_fine[i] = 0u;
    }
    gsync(g);
    dmin = __reduce_min_sync(0xffffffffu, dmin);
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    if (lane == 0) { atomicMin(&sm->dmin, dmin); atomicMax(&sm->dmax, dmax); }
    gsync(g);
    dmin = sm->dmin; dmax = sm->dmax;
    const float scale = 256.0f / ((float)(dmax - dmin) + 1.0f);
#pragma unroll
    for (int e = 0; e < KPT; e++) {
        const uint32_t i = tid + e * NT;
        const float x = (float)((uint32_t)(k[e] >> 32) - dmin) * scale;
        const uint32_t cb = min(255u, (uint32_t)x);
        // neighbouring keys of a surface fall into few coarse bins: one shared-memory atomic per distinct bin per warp
        const unsigned peers = __match_any_sync(0xffffffffu, i < n ? cb : 0xffffffffu);
        if (i < n && (int)lane == __ffs(peers) - 1) atomicAdd(&sm->coarse[cb], (uint32_t)__popc(peers));
    }
    gsync(g);
    {
        const uint32_t c = tid < 256 ? sm->coarse[tid] : 0u;
        const uint32_t ex = group_excl_scan(c, g, sm->w);
        if (tid < 256) sm->cbase[tid] = ex;
    }
    gsync(g);
    uint32_t fr[KPT];  // fine bucket | rank inside it << 16 (both < CAP <= 2^13)
#pragma unroll
    for (int e = 0; e < KPT; e++) {
        const uint32_t i = tid + e * NT;
        fr[e] = 0u;
        if (i < n) {
            const float x = (float)((uint32_t)(k[e] >> 32) - dmin) * scale;
            const uint32_t cb = min(255u, (uint32_t)x);
            const uint32_t cnt = sm->coarse[cb];
            const float frac = fminf(fmaxf(x - (float)cb, 0.0f), 1.0f);
            const uint32_t fb = sm->cbase[cb] + min(cnt - 1u, (uint32_t)(frac * (float)cnt));
            fr[e] = fb | (atomicAdd(&s_fine[fb], 1u) << 16);
        }
    }
    gsync(g);
    {   // counts -> exclusive offsets, KPT consecutive fine buckets per thread (the n fine buckets fit: n < CAP)
        uint32_t c[KPT], sum = 0u;
#pragma unroll
        for (int e = 0; e < KPT; e++) { c[e] = s_fine[tid * KPT + e]; sum += c[e]; }
        uint32_t run = group_excl_scan(sum, g, sm->w);
#pragma unroll
        for (int e = 0; e < KPT; e++) { s_fine[tid * KPT + e] = run; run += c[e]; }
        if (tid == NT - 1) s_fine[CAP] = run;
    }
    gsync(g);
#pragma unroll
    for (int e = 0; e < KPT; e++)
        if (tid + e * NT < n) s_out[s_fine[fr[e] & 0xffffu] + (fr[e] >> 16)] = k[e];
    gsync(g);
    for (uint32_t b = tid; b < n; b += NT) {
        const uint32_t o = s_fine[b], c = s_fine[b + 1] - o;
        if (c < 2u) continue;
        if (c > MAX_FINE) { sm->flag = 1u; atomicMax(&p.hdr->pad0[1], c); continue; }
        for (uint32_t a = 1; a < c; a++) {
            const uint64_t v = s_out[o + a];
            uint32_t j = a;
            while (j > 0 && s_out[o + j - 1] > v) { s_out[o + j] = s_out[o + j - 1]; j--; }
            s_out[o + j] = v;
        }
    }
    gsync(g);
    if (sm->flag) {
        if (tid == 0) atomicAdd(&p.hdr->pad0[0], 1u);  // statistics: tiles that took the network
#pragma unroll
        for (int e = 0; e < KPT; e++)
            if (tid + e * NT < n) s_out[tid + e * NT] = k[e];
        gsync(g);
        bitonic_sort(s_out, n, g);
    }
    write_sorted<2>(p, tile, start, s_out, n, g, sm, s_fine, CAP);
}

// ---- K4: ONE persistent kernel sorts every tile list -------------------------------------------------------------
// CTAs of 1024 threads, one per SM.  A CTA first pulls, with all its threads, the tiles of class 0 (>= 8192 instances,
// network) and class 1 (4096..8191, bucket sort), longest first; then it splits into two 512-thread groups for class 2
// (2048..4095) and into four 256-thread groups for class 3 (< 2048), each group pulling tiles independently on its own
// named barrier.  One launch, no tail between the classes, and the shorter a list the more of them an SM sorts at once
// (the per-list chain of dependent round trips is what a sort costs here, not its instructions).
constexpr int SORT_CTA = 1024;
constexpr uint32_t SMALL_CAP = 2048, MID_CAP = 4096, MED_CAP = 8192;
constexpr int SMALL_NT = 256, SMALL_GROUPS = SORT_CTA / SMALL_NT;
constexpr int MID_NT = 512, MID_GROUPS = SORT_CTA / MID_NT;
constexpr int SMALL_BYTES = SMALL_CAP * 8 + (SMALL_CAP + SMALL_CAP / SMALL_NT + 1) * 4;   // keys out + fine buckets
constexpr int SMALL_STRIDE = (SMALL_BYTES + 127) / 128 * 128;
constexpr int MID_BYTES = MID_CAP * 8 + (MID_CAP + MID_CAP / MID_NT + 1) * 4;
constexpr int MID_STRIDE = (MID_BYTES + 127) / 128 * 128;
constexpr int MED_BYTES = MED_CAP * 8 + (MED_CAP + MED_CAP / SORT_CTA + 1) * 4;
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int SORT_DYN_SMEM = cmax(MED_BYTES, cmax(MID_GROUPS * MID_STRIDE, SMALL_GROUPS * SMALL_STRIDE));

__global__ void __launch_bounds__(SORT_CTA, 1) k_tile_sort(BinParams p)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ GrpSmem s_grp[SMALL_GROUPS];
    if (p.hdr->overflow) return;
    uint32_t tile, start, n;
    {
        Grp g{threadIdx.x, SORT_CTA, 0};
        GrpSmem* sm = &s_grp[0];
        uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_raw);
        while (next_tile(p, 0, g, sm, tile, start, n)) network_sort_tile(p, g, sm, s_keys, tile, start, n);
        uint32_t* s_fine = reinterpret_cast<uint32_t*>(s_raw + MED_CAP * 8);
        while (next_tile(p, 1, g, sm, tile, start, n)) bucket_sort_tile<MED_CAP, SORT_CTA>(p, g, sm, s_keys, s_fine, tile, start, n);
        __syncthreads();
    }
    {
        const int grp = threadIdx.x / MID_NT;
        Grp g{threadIdx.x % MID_NT, MID_NT, 1 + grp};
        GrpSmem* sm = &s_grp[grp];
        uint64_t* s_out = reinterpret_cast<uint64_t*>(s_raw + (size_t)grp * MID_STRIDE);
        uint32_t* s_fine = reinterpret_cast<uint32_t*>(s_raw + (size_t)grp * MID_STRIDE + MID_CAP * 8);
        while (next_tile(p, 2, g, sm, tile, start, n)) bucket_sort_tile<MID_CAP, MID_NT>(p, g, sm, s_out, s_fine, tile, start, n);
        __syncthreads();  // both halves are done before the quarters reuse the shared memory
    }
    {
        const int grp = threadIdx.x / SMALL_NT;
        Grp g{threadIdx.x % SMALL_NT, SMALL_NT, 3 + grp};
        GrpSmem* sm = &s_grp[grp];
        uint64_t* s_out = reinterpret_cast<uint64_t*>(s_raw + (size_t)grp * SMALL_STRIDE);
        uint32_t* s_fine = reinterpret_cast<uint32_t*>(s_raw + (size_t)grp * SMALL_STRIDE + SMALL_CAP * 8);
        while (next_tile(p, 3, g, sm, tile, start, n)) bucket_sort_tile<SMALL_CAP, SMALL_NT>(p, g, sm, s_out, s_fine, tile, start, n);
    }
}

static int g_num_sms = 0;

int tile_sort_setup()
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_DYN_SMEM);
}

void launch_tile_scan(const BinParams& p, cudaStream_t s) { k_tile_scan<<<1, SCAN_THREADS, 0, s>>>(p); }
void launch_emit(const BinParams& p, cudaStream_t s) { k_emit<<<(p.P + 255) / 256, 256, 0, s>>>(p); }
__global__ void k_publish_log(const GHeader* hdr, volatile uint32_t* host_counts)
{
    // words 6..7, not the 4..5 that every blend_fwd writes when it gets there: a blend of an EARLIER call of this thread may still be
    // running on another stream, and the host reads these right behind this kernel
    const unsigned long long need = hdr->log_cursor;
    host_counts[6] = (uint32_t)need;
    host_counts[7] = (uint32_t)(need >> 32);
    __threadfence_system();
}
void launch_publish_log(const GHeader* hdr, volatile uint32_t* host_counts, cudaStream_t s) { k_publish_log<<<1, 1, 0, s>>>(hdr, host_counts); }

void launch_tile_sort(const BinParams& p, cudaStream_t s)
{
    const int sms = g_num_sms > 0 ? g_num_sms : 148;  // persistent: one CTA per SM (148 on B200)
    k_tile_sort<<<min(p.num_tiles, sms), SORT_CTA, SORT_DYN_SMEM, s>>>(p);
}

}  // namespace gstar
