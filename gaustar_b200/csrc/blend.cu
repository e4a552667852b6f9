// blend.cu -- K6 blend_fwd and K7 blend_bwd: per-tile alpha compositing.
//
// Replace renderCUDA forward (DGR/cuda_rasterizer/forward.cu:261-374) and backward
// (DGR/cuda_rasterizer/backward.cu:399-557).  Same per-pixel arithmetic and thresholds
// (power>0 skip, alpha=min(0.99,o*exp(power)), alpha<1/255 skip, T<1e-4 stop), re-mapped for B200:
//
//  * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block;
//  * the tile's depth-sorted Gaussian list is streamed through shared memory in batches of 256
//    48-byte records, each fetched by its own TMA bulk copy (cp.async.bulk -> UBLKCP) that signals
//    an mbarrier; two stages, the next batch is in flight while the current one is composited;
//  * every record carries exact-conservative pixel bounds of {alpha >= 1/255}; a warp first tests 32
//    records against its 8x4 block (one record per lane + ballot) and only walks the survivors.
//    Surface Gaussians cover ~5x5 pixels, so ~3/4 of the (record, warp) pairs of a tile are skipped
//    without evaluating a single exponential, and the result is bit-identical because a culled pair
//    can never pass the reference's own alpha test;
//  * backward: the nine per-(Gaussian,pixel) atomics of the reference become one multi-value
//    butterfly warp reduction and one vector of 9 RED ops per (Gaussian, warp) into a packed
//    [P][12] accumulator; tiles start their back-to-front walk at the tile's max n_contrib.
#include "gstar_common.cuh"
#include "gstar_kernels.h"

namespace gstar {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RS = GSTAR_REC_SMEM;

struct WarpGeom {
    int rx0, ry0, rx1, ry1, px, py;
    bool inside;
};

__device__ __forceinline__ WarpGeom warp_geom(int tile, int gx, int W, int H)
{
    WarpGeom g;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    g.rx0 = tx * GSTAR_TILE + (warp & 1) * 8;
    g.ry0 = ty * GSTAR_TILE + (warp >> 1) * 4;
    g.rx1 = g.rx0 + 7;
    g.ry1 = g.ry0 + 3;
    g.px = g.rx0 + (lane & 7);
    g.py = g.ry0 + (lane >> 3);
    g.inside = g.px < W && g.py < H;
    return g;
}

__device__ __forceinline__ bool bbox_overlaps(const unsigned char* rec, const WarpGeom& g)
{
    const uint2 bb = *reinterpret_cast<const uint2*>(rec + 32);
    const int bx0 = (int)(short)(bb.x & 0xffffu), bx1 = (int)(short)(bb.x >> 16);
    const int by0 = (int)(short)(bb.y & 0xffffu), by1 = (int)(short)(bb.y >> 16);
    return bx0 <= g.rx1 && bx1 >= g.rx0 && by0 <= g.ry1 && by1 >= g.ry0;
}

// forward.cu:332-335 in the operation order of the reference SASS
__device__ __forceinline__ float eval_power(float gx, float gy, float A, float B, float C, float pxf, float pyf, float& dx, float& dy)
{
    dx = __fsub_rn(gx, pxf);
    dy = __fsub_rn(gy, pyf);
    const float q = __fmaf_rn(dx, __fmul_rn(dx, A), __fmul_rn(dy, __fmul_rn(dy, C)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, B)));
}

__global__ void __launch_bounds__(256) k_blend_fwd(BlendParams p)
{
    __shared__ __align__(128) unsigned char s_rec[2][GSTAR_BATCH * RS];
    __shared__ __align__(8) uint64_t s_bar[2];
    if (p.hdr->overflow) return;
    const int tile = blockIdx.x;
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
    const int n = (int)(re - rs);
    const int tid = threadIdx.x, lane = tid & 31;
    const WarpGeom g = warp_geom(tile, p.gx, p.W, p.H);
    const float pxf = (float)g.px, pyf = (float)g.py;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool done = !g.inside;

    if (n > 0) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        const int nb = (n + GSTAR_BATCH - 1) / GSTAR_BATCH;
        const uint32_t* plist = p.point_list + rs;
        const unsigned char* recs = reinterpret_cast<const unsigned char*>(p.recs);
        auto issue = [&](int b) {
            const int cnt = min(GSTAR_BATCH, n - b * GSTAR_BATCH);
            uint64_t* bar = &s_bar[b & 1];
            if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * RS);
            if (tid < cnt) {
                const uint32_t id = __ldg(plist + b * GSTAR_BATCH + tid);
                bulk_g2s(&s_rec[b & 1][tid * RS], recs + (size_t)id * GSTAR_REC_BYTES, RS, bar);
            }
        };
        issue(0);
        bool warp_done = __all_sync(FULL, done);
        int all_done = __syncthreads_and(warp_done);
        for (int b = 0; b < nb; b++) {
            uint64_t* bar = &s_bar[b & 1];
            const uint32_t parity = (uint32_t)(b >> 1) & 1u;
            if (all_done) {  // batch b is still in flight: it must land before the CTA may exit
                mbar_wait(bar, parity);
                break;
            }
            if (b + 1 < nb) issue(b + 1);
            mbar_wait(bar, parity);
            if (!warp_done) {
                const unsigned char* buf = s_rec[b & 1];
                const int cnt = min(GSTAR_BATCH, n - b * GSTAR_BATCH);
                for (int r0 = 0; r0 < cnt; r0 += 32) {
                    const int j = r0 + lane;
                    const bool ov = (j < cnt) && bbox_overlaps(buf + j * RS, g);
                    unsigned m = __ballot_sync(FULL, ov);
                    while (m) {
                        const int k = __ffs(m) - 1;
                        m &= m - 1;
                        const unsigned char* rp = buf + (r0 + k) * RS;
                        const float4 q0 = *reinterpret_cast<const float4*>(rp);       // x y A B
                        const float4 q1 = *reinterpret_cast<const float4*>(rp + 16);  // C o r g
                        const float cb = *reinterpret_cast<const float*>(rp + 40);    // b
                        if (!done) {
                            float dx, dy;
                            const float power = eval_power(q0.x, q0.y, q0.z, q0.w, q1.x, pxf, pyf, dx, dy);
                            if (!(power > 0.0f)) {
                                const float alpha = fminf(0.99f, __fmul_rn(q1.y, expf(power)));
                                if (!(alpha < 1.0f / 255.0f)) {
                                    const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                                    if (test_T < 0.0001f) {
                                        done = true;
                                    } else {
                                        C0 = __fmaf_rn(T, __fmul_rn(alpha, q1.z), C0);
                                        C1 = __fmaf_rn(T, __fmul_rn(alpha, q1.w), C1);
                                        C2 = __fmaf_rn(T, __fmul_rn(alpha, cb), C2);
                                        T = test_T;
                                        last = (uint32_t)(b * GSTAR_BATCH + r0 + k + 1);
                                    }
                                }
                            }
                        }
                    }
                    if (__all_sync(FULL, done)) {
                        warp_done = true;
                        break;
                    }
                }
            }
            all_done = __syncthreads_and(warp_done);
        }
    }
    if (g.inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pid = (size_t)g.py * p.W + g.px;
        p.final_T[pid] = T;
        p.n_contrib[pid] = last;
        p.out_color[pid] = __fmaf_rn(__ldg(p.bg + 0), T, C0);  // forward.cu:372
        p.out_color[HW + pid] = __fmaf_rn(__ldg(p.bg + 1), T, C1);
        p.out_color[2 * HW + pid] = __fmaf_rn(__ldg(p.bg + 2), T, C2);
    }
}

// keep = own half, send = other half; after the exchange every lane holds the pair-sum of its half
__device__ __forceinline__ float bfly_pair(float a, float b, unsigned m, unsigned lane)
{
    const bool up = (lane & m) != 0;
    const float keep = up ? b : a, send = up ? a : b;
    return keep + __shfl_xor_sync(FULL, send, m);
}

__global__ void __launch_bounds__(256) k_blend_bwd(BlendParams p)
{
    __shared__ __align__(128) unsigned char s_rec[2][GSTAR_BATCH * RS];
    __shared__ uint32_t s_idx[2][GSTAR_BATCH];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_kmax;
    if (p.hdr->overflow) return;
    const int tile = blockIdx.x;
    const uint32_t rs = p.ranges[2 * tile], re = p.ranges[2 * tile + 1];
    if (re == rs) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const WarpGeom g = warp_geom(tile, p.gx, p.W, p.H);
    const float pxf = (float)g.px, pyf = (float)g.py;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pid = (size_t)g.py * p.W + g.px;

    const float T_final = g.inside ? p.final_T[pid] : 0.f;
    const int last_contributor = g.inside ? (int)p.n_contrib[pid] : 0;
    float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f;
    if (g.inside) { dpx0 = p.dL_dpix[pid]; dpx1 = p.dL_dpix[HW + pid]; dpx2 = p.dL_dpix[2 * HW + pid]; }
    float bg_dot = 0.f;  // backward.cu:531-533
    bg_dot += __ldg(p.bg + 0) * dpx0;
    bg_dot += __ldg(p.bg + 1) * dpx1;
    bg_dot += __ldg(p.bg + 2) * dpx2;
    const float ddelx_dx = 0.5f * p.W, ddely_dy = 0.5f * p.H;  // backward.cu:460-461

    if (tid == 0) {
        s_kmax = 0;
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int warp_kmax = __reduce_max_sync(FULL, last_contributor);
    if (lane == 0 && warp_kmax > 0) atomicMax(&s_kmax, warp_kmax);
    __syncthreads();
    const int total = s_kmax;  // entries [0,total) of the tile list can matter; the rest is behind every pixel's last contributor
    if (total == 0) return;

    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    const int nb = (total + GSTAR_BATCH - 1) / GSTAR_BATCH;
    const uint32_t* plist = p.point_list + rs;
    const unsigned char* recs = reinterpret_cast<const unsigned char*>(p.recs);
    // slot s of batch b  <->  list position k = total-1 - (b*256+s)   (back to front, backward.cu:472)
    auto issue = [&](int b) {
        const int cnt = min(GSTAR_BATCH, total - b * GSTAR_BATCH);
        uint64_t* bar = &s_bar[b & 1];
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * RS);
        if (tid < cnt) {
            const uint32_t id = __ldg(plist + (total - 1 - (b * GSTAR_BATCH + tid)));
            s_idx[b & 1][tid] = id;
            bulk_g2s(&s_rec[b & 1][tid * RS], recs + (size_t)id * GSTAR_REC_BYTES, RS, bar);
        }
    };
    issue(0);
    for (int b = 0; b < nb; b++) {
        if (b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[b & 1], (uint32_t)(b >> 1) & 1u);
        __syncthreads();  // s_idx of this batch visible to all warps
        const unsigned char* buf = s_rec[b & 1];
        const uint32_t* ids = s_idx[b & 1];
        const int cnt = min(GSTAR_BATCH, total - b * GSTAR_BATCH);
        const int kbase = total - 1 - b * GSTAR_BATCH;
        if (kbase - (cnt - 1) < warp_kmax) {  // some entry of this batch is in front of this warp's last contributor
            for (int r0 = 0; r0 < cnt; r0 += 32) {
                const int j = r0 + lane;
                const bool ov = (j < cnt) && (kbase - j < warp_kmax) && bbox_overlaps(buf + j * RS, g);
                unsigned m = __ballot_sync(FULL, ov);
                while (m) {
                    const int s = __ffs(m) - 1;
                    m &= m - 1;
                    const int slot = r0 + s;
                    const int k = kbase - slot;  // == reference `contributor` after its decrement
                    const unsigned char* rp = buf + slot * RS;
                    const float4 q0 = *reinterpret_cast<const float4*>(rp);
                    const float4 q1 = *reinterpret_cast<const float4*>(rp + 16);
                    const float cb = *reinterpret_cast<const float*>(rp + 40);
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
                    bool contrib = false;
                    if (k < last_contributor) {
                        float dx, dy;
                        const float power = eval_power(q0.x, q0.y, q0.z, q0.w, q1.x, pxf, pyf, dx, dy);
                        if (!(power > 0.0f)) {
                            const float G = expf(power);
                            const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
                            if (!(alpha < 1.0f / 255.0f)) {
                                contrib = true;
                                T = __fdiv_rn(T, __fsub_rn(1.f, alpha));
                                const float dchannel_dcolor = alpha * T;
                                float dL_dalpha = 0.0f;
                                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = q1.z; dL_dalpha += (q1.z - acc0) * dpx0;
                                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = q1.w; dL_dalpha += (q1.w - acc1) * dpx1;
                                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = cb;   dL_dalpha += (cb - acc2) * dpx2;
                                v5 = dchannel_dcolor * dpx0;
                                v6 = dchannel_dcolor * dpx1;
                                v7 = dchannel_dcolor * dpx2;
                                dL_dalpha *= T;
                                last_alpha = alpha;
                                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                                const float dL_dG = q1.y * dL_dalpha;
                                const float gdx = G * dx, gdy = G * dy;
                                const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
                                const float dG_ddely = -gdy * q1.x - gdx * q0.w;
                                v0 = dL_dG * dG_ddelx * ddelx_dx;
                                v1 = dL_dG * dG_ddely * ddely_dy;
                                v2 = -0.5f * gdx * dx * dL_dG;
                                v3 = -0.5f * gdx * dy * dL_dG;
                                v4 = -0.5f * gdy * dy * dL_dG;
                                v8 = G * dL_dalpha;
                            }
                        }
                    }
                    if (__any_sync(FULL, contrib)) {
                        // 8 values: three halving exchange steps, then two plain steps; lane 4*i ends with sum of v_i
                        const float w0 = bfly_pair(v0, v4, 16, lane), w1 = bfly_pair(v1, v5, 16, lane);
                        const float w2 = bfly_pair(v2, v6, 16, lane), w3 = bfly_pair(v3, v7, 16, lane);
                        const float u0 = bfly_pair(w0, w2, 8, lane), u1 = bfly_pair(w1, w3, 8, lane);
                        float t = bfly_pair(u0, u1, 4, lane);
                        t += __shfl_xor_sync(FULL, t, 2);
                        t += __shfl_xor_sync(FULL, t, 1);
                        v8 += __shfl_xor_sync(FULL, v8, 16);
                        v8 += __shfl_xor_sync(FULL, v8, 8);
                        v8 += __shfl_xor_sync(FULL, v8, 4);
                        v8 += __shfl_xor_sync(FULL, v8, 2);
                        v8 += __shfl_xor_sync(FULL, v8, 1);
                        float* dst = p.gacc + (size_t)ids[slot] * GSTAR_GACC;
                        if ((lane & 3) == 0) atomicAdd(dst + (lane >> 2), t);
                        else if (lane == 1) atomicAdd(dst + 8, v8);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with stage b&1 before batch b+2 overwrites it
    }
}

void launch_blend_fwd(const BlendParams& p, cudaStream_t s) { k_blend_fwd<<<p.gx * p.gy, 256, 0, s>>>(p); }
void launch_blend_bwd(const BlendParams& p, cudaStream_t s) { k_blend_bwd<<<p.gx * p.gy, 256, 0, s>>>(p); }

}  // namespace gstar
