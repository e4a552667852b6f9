// preprocess.cu -- per-Gaussian kernels: K1 preprocess_fwd, K8 preprocess_bwd, mark_visible, unpack.
//
// K1 replaces preprocessCUDA (DGR/cuda_rasterizer/forward.cu:155-256) and also counts, per 16x16
// tile, how many Gaussians touch it (the histogram that replaces the reference's Gaussian-major
// InclusiveSum, rasterizer_impl.cu:277, as the first half of the MSD tile|depth sort).
// K8 fuses computeCov2DCUDA (backward.cu:144-274) and preprocessCUDA-backward (backward.cu:346-396).
//
// Bit-exactness: every scalar that decides a sort key or a tile assignment (view depth, pixel
// centre, cov2D, det, conic, radius, rect) is computed with explicit round-to-nearest intrinsics in
// the order the reference's sm_100a SASS performs them, so neither NVVM nor ptxas can contract
// differently here than they did for the reference (see DESIGN.md "bit-exact keys").
#include <stdio.h>

#include "gstar_common.cuh"
#include "gstar_kernels.h"

namespace gstar {

__constant__ float c_SH_C0 = 0.28209479177387814f;  // auxiliary.h:22-39
__constant__ float c_SH_C1 = 0.4886025119029199f;
__constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                                 0.5462742152960396f};
__constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

// auxiliary.h:58-77: m0*x + m4*y + m8*z + m12  ==  fma(z,m8, fma(x,m0, y*m4)) + m12 in the reference SASS
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z)
{
    return __fadd_rn(__fmaf_rn(z, m[8 + r], __fmaf_rn(x, m[r], __fmul_rn(y, m[4 + r]))), m[12 + r]);
}

// forward.cu:118-152 (quaternion used as given)
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float4 q, float* c6)
{
    sx = __fmul_rn(mod, sx); sy = __fmul_rn(mod, sy); sz = __fmul_rn(mod, sz);
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float two = 2.0f;
    const float R00 = __fsub_rn(1.0f, __fmul_rn(two, __fadd_rn(yy, zz)));
    const float R01 = __fmul_rn(two, __fmaf_rn(x, y, -__fmul_rn(r, z)));
    const float R02 = __fmul_rn(two, __fmaf_rn(r, y, __fmul_rn(x, z)));
    const float R10 = __fmul_rn(two, __fmaf_rn(x, y, __fmul_rn(r, z)));
    const float R11 = __fsub_rn(1.0f, __fmul_rn(two, __fmaf_rn(x, x, zz)));
    const float R12 = __fmul_rn(two, __fmaf_rn(y, z, -__fmul_rn(r, x)));
    const float R20 = __fmul_rn(two, __fmaf_rn(-r, y, __fmul_rn(x, z)));
    const float R21 = __fmul_rn(two, __fmaf_rn(y, z, __fmul_rn(r, x)));
    const float R22 = __fsub_rn(1.0f, __fmul_rn(two, __fmaf_rn(x, x, yy)));
    const float m00 = __fmul_rn(sx, R00), m01 = __fmul_rn(sy, R01), m02 = __fmul_rn(sz, R02);
    const float m10 = __fmul_rn(sx, R10), m11 = __fmul_rn(sy, R11), m12 = __fmul_rn(sz, R12);
    const float m20 = __fmul_rn(sx, R20), m21 = __fmul_rn(sy, R21), m22 = __fmul_rn(sz, R22);
    c6[0] = __fmaf_rn(m02, m02, __fmaf_rn(m00, m00, __fmul_rn(m01, m01)));
    c6[1] = __fmaf_rn(m12, m02, __fmaf_rn(m10, m00, __fmul_rn(m11, m01)));
    c6[2] = __fmaf_rn(m22, m02, __fmaf_rn(m20, m00, __fmul_rn(m21, m01)));
    c6[3] = __fmaf_rn(m12, m12, __fmaf_rn(m10, m10, __fmul_rn(m11, m11)));
    c6[4] = __fmaf_rn(m22, m12, __fmaf_rn(m20, m10, __fmul_rn(m21, m11)));
    c6[5] = __fmaf_rn(m22, m22, __fmaf_rn(m20, m20, __fmul_rn(m21, m21)));
}

struct Cov2D {
    float Tx[3], Ty[3], Ax[3], Ay[3];
    float a, b, c;
    float tx, ty, tz, txtz, tytz;
};

// forward.cu:74-113
__device__ __forceinline__ void cov2d_from_cov3d(float px, float py, float pz, float fx, float fy, float tan_fovx, float tan_fovy,
                                                 const float* c6, const float* vm, Cov2D& o)
{
    float tx = xform_row(vm, 0, px, py, pz);
    float ty = xform_row(vm, 1, px, py, pz);
    const float tz = xform_row(vm, 2, px, py, pz);
    const float limx = __fmul_rn(1.3f, tan_fovx), limy = __fmul_rn(1.3f, tan_fovy);
    const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
    tx = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), tz);
    ty = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), tz);
    const float J00 = __fdiv_rn(fx, tz), J11 = __fdiv_rn(fy, tz);
    const float tz2 = __fmul_rn(tz, tz);
    const float J02 = __fdiv_rn(-__fmul_rn(fx, tx), tz2);
    const float J12 = __fdiv_rn(-__fmul_rn(fy, ty), tz2);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o.Tx[k] = __fmaf_rn(vm[2 + 4 * k], J02, __fmul_rn(vm[4 * k], J00));
        o.Ty[k] = __fmaf_rn(vm[2 + 4 * k], J12, __fmul_rn(vm[1 + 4 * k], J11));
    }
    const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o.Ax[k] = __fmaf_rn(o.Tx[2], V[2][k], __fmaf_rn(o.Tx[0], V[0][k], __fmul_rn(o.Tx[1], V[1][k])));
        o.Ay[k] = __fmaf_rn(o.Ty[2], V[2][k], __fmaf_rn(o.Ty[0], V[0][k], __fmul_rn(o.Ty[1], V[1][k])));
    }
    const float c00 = __fmaf_rn(o.Tx[2], o.Ax[2], __fmaf_rn(o.Tx[0], o.Ax[0], __fmul_rn(o.Tx[1], o.Ax[1])));
    const float c01 = __fmaf_rn(o.Tx[2], o.Ay[2], __fmaf_rn(o.Tx[0], o.Ay[0], __fmul_rn(o.Tx[1], o.Ay[1])));
    const float c11 = __fmaf_rn(o.Ty[2], o.Ay[2], __fmaf_rn(o.Ty[0], o.Ay[0], __fmul_rn(o.Ty[1], o.Ay[1])));
    o.a = __fadd_rn(c00, 0.3f);
    o.b = c01;
    o.c = __fadd_rn(c11, 0.3f);
    o.tx = tx; o.ty = ty; o.tz = tz; o.txtz = txtz; o.tytz = tytz;
}

// auxiliary.h:41-44, evaluated in double like the reference: fma(v+1, S, -1) * 0.5
__device__ __forceinline__ float ndc2pix(float v, int S)
{
    return __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}

// auxiliary.h:46-56
__device__ __forceinline__ void get_rect(float px, float py, int radius, int gx, int gy, uint32_t& minx, uint32_t& miny, uint32_t& maxx,
                                         uint32_t& maxy)
{
    const float r = (float)radius;
    minx = min((uint32_t)gx, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(px, r), 0.0625f)));
    miny = min((uint32_t)gy, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(py, r), 0.0625f)));
    maxx = min((uint32_t)gx, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), 1.0f), 0.0625f)));
    maxy = min((uint32_t)gy, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), 1.0f), 0.0625f)));
}

// ---- SH rows staged through shared memory ------------------------------------------------------------
// A Gaussian's SH row is 12*M contiguous bytes (192 B for degree 3).  One thread per Gaussian reading its
// own row touches 32 different cache lines per load instruction; instead each warp copies the 32 rows of
// its Gaussians with fully coalesced loads into padded shared-memory rows (stride 3M+4 floats when rows
// are 16-byte aligned, else 3M+1: both conflict-free for one-row-per-lane access) and every lane then
// reads -- and in the backward pass overwrites with dL_dsh -- only its own row.
constexpr int SH_MAX_M = 16;
constexpr int SH_ROW_STRIDE_MAX = 3 * SH_MAX_M + 4;                 // floats
constexpr int SH_WARP_FLOATS = 32 * SH_ROW_STRIDE_MAX;              // per warp
constexpr int SH_SMEM_BYTES = 8 * SH_WARP_FLOATS * (int)sizeof(float);  // per 256-thread CTA

struct ShStage {
    float* rows;   // this warp's shared-memory region
    int stride;    // floats between rows
    bool vec;      // rows are 16-byte aligned multiples of 16 bytes
};

__device__ __forceinline__ ShStage sh_stage_make(float* smem_cta, const float* shs, int M)
{
    ShStage st;
    st.rows = smem_cta + (threadIdx.x >> 5) * SH_WARP_FLOATS;
    st.vec = ((3 * M) % 4 == 0) && ((((size_t)shs) & 15) == 0);
    st.stride = st.vec ? 3 * M + 4 : 3 * M + 1;
    return st;
}

// coalesced global -> shared copy of the rows selected by need_mask (bit g = row of lane g)
__device__ __forceinline__ void sh_stage_load(const ShStage& st, const float* __restrict__ shs, int M, long long first_row, int P, unsigned need_mask)
{
    const int lane = threadIdx.x & 31;
    const int rowf = 3 * M;
    const float* base = shs + (size_t)first_row * rowf;
    if (st.vec && M == 16) {  // degree-3 rows: 12 float4 per row, padded to 13 -- constants let the division fold
        const float4* b4 = reinterpret_cast<const float4*>(base);
        float4* r4 = reinterpret_cast<float4*>(st.rows);
        float4 v[12];
#pragma unroll
        for (int it = 0; it < 12; it++) {  // all 12 loads of the lane in flight before the first store
            const int q = lane + 32 * it, g = q / 12;
            v[it] = (((need_mask >> g) & 1u) && first_row + g < P) ? ldg_nc_f4(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 12; it++) {
            const int q = lane + 32 * it, g = q / 12, j = q - g * 12;
            r4[g * 13 + j] = v[it];
        }
    } else if (st.vec) {
        const int row4 = rowf >> 2, str4 = st.stride >> 2;
        const float4* b4 = reinterpret_cast<const float4*>(base);
        float4* r4 = reinterpret_cast<float4*>(st.rows);
        for (int q = lane; q < 32 * row4; q += 32) {
            const int g = q / row4, j = q - g * row4;
            if (((need_mask >> g) & 1u) && first_row + g < P) r4[g * str4 + j] = ldg_nc_f4(b4 + q);
        }
    } else {
        for (int q = lane; q < 32 * rowf; q += 32) {
            const int g = q / rowf, j = q - g * rowf;
            if (((need_mask >> g) & 1u) && first_row + g < P) st.rows[g * st.stride + j] = __ldg(base + q);
        }
    }
    __syncwarp();
}

// coalesced shared -> global copy of the rows selected by row_mask; accumulate 1: global += shared (plain read-modify-write: one
// backward at a time per array), 2: the same with reductions at L2 (red.global.add: several backward calls on different streams
// may add into the same array at once)
__device__ __forceinline__ void sh_stage_store(const ShStage& st, float* __restrict__ dst, int M, long long first_row, int P, unsigned row_mask,
                                               int accumulate)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int rowf = 3 * M;
    float* base = dst + (size_t)first_row * rowf;
    if (st.vec && M == 16 && ((((size_t)dst) & 15) == 0)) {
        float4* b4 = reinterpret_cast<float4*>(base);
        const float4* r4 = reinterpret_cast<const float4*>(st.rows);
        float4 o[12];
        if (accumulate == 2) {
#pragma unroll
            for (int it = 0; it < 12; it++) {
                const int q = lane + 32 * it, g = q / 12, j = q - g * 12;
                if (((row_mask >> g) & 1u) && first_row + g < P) {
                    const float4 v = r4[g * 13 + j];
                    red_add_v4(reinterpret_cast<float*>(b4 + q), v.x, v.y, v.z, v.w);
                }
            }
            return;
        }
        if (accumulate) {
#pragma unroll
            for (int it = 0; it < 12; it++) {
                const int q = lane + 32 * it, g = q / 12;
                o[it] = (((row_mask >> g) & 1u) && first_row + g < P) ? b4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int it = 0; it < 12; it++) {
            const int q = lane + 32 * it, g = q / 12, j = q - g * 12;
            if (((row_mask >> g) & 1u) && first_row + g < P) {
                float4 v = r4[g * 13 + j];
                if (accumulate) { v.x += o[it].x; v.y += o[it].y; v.z += o[it].z; v.w += o[it].w; }
                b4[q] = v;
            }
        }
    } else if (st.vec && ((((size_t)dst) & 15) == 0)) {
        const int row4 = rowf >> 2, str4 = st.stride >> 2;
        float4* b4 = reinterpret_cast<float4*>(base);
        const float4* r4 = reinterpret_cast<const float4*>(st.rows);
        for (int q = lane; q < 32 * row4; q += 32) {
            const int g = q / row4, j = q - g * row4;
            if (((row_mask >> g) & 1u) && first_row + g < P) {
                float4 v = r4[g * str4 + j];
                if (accumulate == 2) { red_add_v4(reinterpret_cast<float*>(b4 + q), v.x, v.y, v.z, v.w); continue; }
                if (accumulate) {
                    const float4 o = b4[q];
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                b4[q] = v;
            }
        }
    } else {
        for (int q = lane; q < 32 * rowf; q += 32) {
            const int g = q / rowf, j = q - g * rowf;
            if (((row_mask >> g) & 1u) && first_row + g < P) {
                const float v = st.rows[g * st.stride + j];
                if (accumulate == 2) atomicAdd(base + q, v);
                else base[q] = accumulate ? base[q] + v : v;
            }
        }
    }
}

// Floats [4*q0, 4*q1) of a coefficient row into v (static indices after unrolling: registers).  VEC: the row is a
// 16-byte aligned shared-memory row of the staging area (stride 3M+4 floats): one conflict-free LDS.128 per four floats,
// where scalar reads of the same element of 32 rows would be 4-way bank conflicted (rows are float4-aligned).
template <bool VEC>
__device__ __forceinline__ void sh_row_load(const float* c, int q0, int q1, float* v)
{
#pragma unroll
    for (int q = q0; q < q1; q++) {
        if (VEC) {
            const float4 t = reinterpret_cast<const float4*>(c)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) v[4 * q + i] = c[4 * q + i];
        }
    }
}

// forward.cu:20-71.  `c` = this Gaussian's coefficients [M][3] (shared-memory row or global memory).
template <bool VEC>
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* c, float3 pos, float3 campos, uint32_t& clamp_bits)
{
    float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
    float v[48];
#define SHK(k) make_float3(v[3 * (k)], v[3 * (k) + 1], v[3 * (k) + 2])
#define ACC(w, k) { const float w_ = (w); const float3 s_ = SHK(k); res.x += w_ * s_.x; res.y += w_ * s_.y; res.z += w_ * s_.z; }
    if (VEC || deg > 0) sh_row_load<VEC>(c, 0, 1, v);
    else { v[0] = c[0]; v[1] = c[1]; v[2] = c[2]; }  // a degree-0 row in global memory may be only three floats long
    float3 res = {c_SH_C0 * v[0], c_SH_C0 * v[1], c_SH_C0 * v[2]};
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        sh_row_load<VEC>(c, 1, 3, v);  // coefficients 1-3 = floats 3..11
        ACC(-c_SH_C1 * y, 1);
        ACC(c_SH_C1 * z, 2);
        ACC(-c_SH_C1 * x, 3);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            if (VEC) sh_row_load<VEC>(c, 3, 7, v);  // coefficients 4-8 = floats 12..26 (+ one of coefficient 9)
            else {
#pragma unroll
                for (int i = 12; i < 27; i++) v[i] = c[i];
            }
            ACC(c_SH_C2[0] * xy, 4);
            ACC(c_SH_C2[1] * yz, 5);
            ACC(c_SH_C2[2] * (2.0f * zz - xx - yy), 6);
            ACC(c_SH_C2[3] * xz, 7);
            ACC(c_SH_C2[4] * (xx - yy), 8);
            if (deg > 2) {
                if (VEC) sh_row_load<VEC>(c, 7, 12, v);  // coefficients 9-15 = floats 27..47
                else {
#pragma unroll
                    for (int i = 27; i < 48; i++) v[i] = c[i];
                }
                ACC(c_SH_C3[0] * y * (3.0f * xx - yy), 9);
                ACC(c_SH_C3[1] * xy * z, 10);
                ACC(c_SH_C3[2] * y * (4.0f * zz - xx - yy), 11);
                ACC(c_SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), 12);
                ACC(c_SH_C3[4] * x * (4.0f * zz - xx - yy), 13);
                ACC(c_SH_C3[5] * z * (xx - yy), 14);
                ACC(c_SH_C3[6] * x * (xx - 3.0f * yy), 15);
            }
        }
    }
#undef ACC
#undef SHK
    res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
    clamp_bits = (res.x < 0 ? 1u : 0u) | (res.y < 0 ? 2u : 0u) | (res.z < 0 ? 4u : 0u);
    return make_float3(fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f));
}

__global__ void __launch_bounds__(256, 3) k_preprocess_fwd(PreFwdParams p)
{
    extern __shared__ __align__(16) float s_sh[];
    __shared__ float s_vm[16], s_pm[16];
    if (threadIdx.x < 16) s_vm[threadIdx.x] = p.viewmatrix[threadIdx.x];
    else if (threadIdx.x < 32) s_pm[threadIdx.x - 16] = p.projmatrix[threadIdx.x - 16];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const bool valid = idx < p.P;
    GRec rec;
    GAux aux;
    {
        float4 z = {0.f, 0.f, 0.f, 0.f};
        float4* r4 = reinterpret_cast<float4*>(&rec);
        r4[0] = z; r4[1] = z; r4[2] = z;
        *reinterpret_cast<float4*>(&aux) = z;
    }
    uint32_t minx = 0, miny = 0, maxx = 0, maxy = 0;
    bool visible = false;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (valid) {
        do {
            px = p.means3D[3 * (size_t)idx]; py = p.means3D[3 * (size_t)idx + 1]; pz = p.means3D[3 * (size_t)idx + 2];
            const float pvz = xform_row(s_vm, 2, px, py, pz);  // auxiliary.h:152-154
            if (pvz <= 0.2f) {
                if (p.prefiltered) {
                    printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                    __trap();
                }
                break;
            }
            const float hx = xform_row(s_pm, 0, px, py, pz);
            const float hy = xform_row(s_pm, 1, px, py, pz);
            const float hw = xform_row(s_pm, 3, px, py, pz);
            const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
            const float projx = __fmul_rn(hx, p_w), projy = __fmul_rn(hy, p_w);
            float c6[6];
            if (p.cov3D_precomp) {
#pragma unroll
                for (int k = 0; k < 6; k++) c6[k] = p.cov3D_precomp[6 * (size_t)idx + k];
            } else {
                const float4 q = *reinterpret_cast<const float4*>(p.rotations + 4 * (size_t)idx);
                cov3d_from_scale_rot(p.scales[3 * (size_t)idx], p.scales[3 * (size_t)idx + 1], p.scales[3 * (size_t)idx + 2], p.scale_modifier, q,
                                     c6);
            }
            Cov2D cv;
            cov2d_from_cov3d(px, py, pz, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, c6, s_vm, cv);
            const float det = __fmaf_rn(cv.a, cv.c, -__fmul_rn(cv.b, cv.b));  // forward.cu:219
            if (det == 0.0f) break;
            const float det_inv = __frcp_rn(det);
            const float conx = __fmul_rn(cv.c, det_inv), cony = __fmul_rn(-cv.b, det_inv), conz = __fmul_rn(cv.a, det_inv);
            const float mid = __fmul_rn(0.5f, __fadd_rn(cv.a, cv.c));
            const float sq = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));  // forward.cu:230-231
            const float lam = fmaxf(__fadd_rn(mid, sq), __fsub_rn(mid, sq));
            const float my_radius = ceilf(__fmul_rn(3.f, __fsqrt_rn(lam)));  // forward.cu:232
            const float pix = ndc2pix(projx, p.W), piy = ndc2pix(projy, p.H);
            const int radius = (int)my_radius;
            get_rect(pix, piy, radius, p.gx, p.gy, minx, miny, maxx, maxy);
            if ((maxx - minx) * (maxy - miny) == 0) break;
            visible = true;
            const float opac = p.opacities[idx];
            // Exact-conservative pixel bounds of {alpha >= 1/255}: alpha = o*exp(-q/2) with q >= dx^2/cov_xx, so a pixel
            // with |dx| > sqrt(2 ln(255 o) cov_xx) can never pass the reference's alpha test (forward.cu:343-345).
            // Margins (1e-3 on the log, 5e-4 relative + 2e-3 px on the extent) dwarf every fp32 rounding involved.
            int bx0 = -32768, bx1 = 32767, by0 = -32768, by1 = 32767;
            if (opac < (1.0f / 255.0f)) {
                bx0 = 1; bx1 = 0; by0 = 1; by1 = 0;  // o*G <= o < 1/255 : never blended
            } else if (det > 0.0f && cv.a > 0.0f && cv.c > 0.0f && opac <= 1e30f) {
                const float tau = logf(255.0f * opac) * 1.001f + 0.001f;
                const float ex = sqrtf(2.0f * tau * cv.a) * 1.0005f + 0.002f;
                const float ey = sqrtf(2.0f * tau * cv.c) * 1.0005f + 0.002f;
                bx0 = (int)fmaxf(-32768.f, fminf(32767.f, ceilf(pix - ex)));
                bx1 = (int)fmaxf(-32768.f, fminf(32767.f, floorf(pix + ex)));
                by0 = (int)fmaxf(-32768.f, fminf(32767.f, ceilf(piy - ey)));
                by1 = (int)fmaxf(-32768.f, fminf(32767.f, floorf(piy + ey)));
            }
            rec.x = pix; rec.y = piy; rec.A = conx; rec.B = cony; rec.C = conz; rec.o = opac;
            rec.bbox_x = ((uint32_t)bx0 & 0xffffu) | ((uint32_t)bx1 << 16);
            rec.bbox_y = ((uint32_t)by0 & 0xffffu) | ((uint32_t)by1 << 16);
            aux.depth = pvz;
            aux.rect_min = minx | (miny << 16);
            aux.rect_max = maxx | (maxy << 16);
            aux.radius = radius;
        } while (0);
    }
    // colour: colors_precomp, or SH -> RGB with the warp's SH rows staged through shared memory
    if (p.colors_precomp) {
        if (visible) {
            rec.r = p.colors_precomp[3 * (size_t)idx]; rec.g = p.colors_precomp[3 * (size_t)idx + 1]; rec.b = p.colors_precomp[3 * (size_t)idx + 2];
        }
    } else {
        const unsigned need = __ballot_sync(0xffffffffu, visible);
        const int nfl = 3 * (p.D + 1) * (p.D + 1);
        const bool staged = p.M <= SH_MAX_M && 2 * nfl >= 3 * p.M;  // reading whole rows pays off only if most of a row is used
        const float3 campos = make_float3(p.campos[0], p.campos[1], p.campos[2]);
        uint32_t clamp_bits = 0;
        float3 col = {0.f, 0.f, 0.f};
        if (staged) {
            if (need) {
                const ShStage st = sh_stage_make(s_sh, p.shs, p.M);
                sh_stage_load(st, p.shs, p.M, (long long)idx - lane, p.P, need);
                if (visible) {
                    const float* row = st.rows + lane * st.stride;
                    col = st.vec ? sh_to_rgb<true>(p.D, row, make_float3(px, py, pz), campos, clamp_bits)
                                 : sh_to_rgb<false>(p.D, row, make_float3(px, py, pz), campos, clamp_bits);
                }
            }
        } else if (visible) {
            float c[48];
            const float* sh = p.shs + (size_t)3 * p.M * idx;
            for (int i = 0; i < 48; i++)
                if (i < nfl) c[i] = __ldg(sh + i);
            col = sh_to_rgb<false>(p.D, c, make_float3(px, py, pz), campos, clamp_bits);
        }
        if (visible) { rec.r = col.x; rec.g = col.y; rec.b = col.z; rec.flags = clamp_bits; }
    }
    if (valid) {
        float4* dst = reinterpret_cast<float4*>(p.recs + idx);
        const float4* src = reinterpret_cast<const float4*>(&rec);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        *reinterpret_cast<float4*>(p.aux + idx) = *reinterpret_cast<const float4*>(&aux);
        p.radii[idx] = aux.radius;
    }
    // per-tile histogram (first digit of the MSD tile|depth sort): one atomic per distinct tile per warp
    uint32_t* cnt = p.tile_count;
    const int gx = p.gx;
    const uint32_t w = maxx - minx, h = maxy - miny;
    const uint32_t nsmall = (visible && w * h <= 4) ? w * h : 0u;
    const uint32_t rounds = __reduce_max_sync(0xffffffffu, nsmall);
    uint32_t tx = minx, ty = miny;  // walks the rect row by row (no integer division in the loop)
    for (uint32_t t = 0; t < rounds; t++) {
        const bool act = t < nsmall;
        const uint32_t tile = act ? ty * gx + tx : 0xffffffffu;
        if (++tx == maxx) { tx = minx; ty++; }
        const unsigned peers = __match_any_sync(0xffffffffu, tile);
        if (act && (int)lane == __ffs(peers) - 1) atomicAdd(cnt + tile, (uint32_t)__popc(peers));
    }
    unsigned big = __ballot_sync(0xffffffffu, visible && w * h > 4);
    while (big) {  // large rects: the whole warp walks one Gaussian's rect at a time
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bx = __shfl_sync(0xffffffffu, minx, src), by = __shfl_sync(0xffffffffu, miny, src);
        const uint32_t bw = __shfl_sync(0xffffffffu, w, src), ba = bw * __shfl_sync(0xffffffffu, h, src);
        for (uint32_t t = lane; t < ba; t += 32) atomicAdd(cnt + (by + t / bw) * gx + (bx + t % bw), 1u);
    }
}

// rasterizer_impl.cu:54-66 checkFrustum
__global__ void k_mark_visible(int P, const float* __restrict__ means3D, const float* __restrict__ vm, unsigned char* __restrict__ present)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = __ldg(vm + i);
    present[idx] = xform_row(m, 2, means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]) > 0.2f;
}

__global__ void k_geom_unpack(const GRec* __restrict__ recs, const GAux* __restrict__ auxs, int P, float* depths, float* means2D, float* conic_opacity, float* rgb,
                              uint32_t* tiles_touched, unsigned char* clamped)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const GRec r = recs[idx];
    const GAux a = auxs[idx];
    if (depths) depths[idx] = a.depth;
    if (means2D) { means2D[2 * idx] = r.x; means2D[2 * idx + 1] = r.y; }
    if (conic_opacity) { conic_opacity[4 * idx] = r.A; conic_opacity[4 * idx + 1] = r.B; conic_opacity[4 * idx + 2] = r.C; conic_opacity[4 * idx + 3] = r.o; }
    if (rgb) { rgb[3 * idx] = r.r; rgb[3 * idx + 1] = r.g; rgb[3 * idx + 2] = r.b; }
    if (tiles_touched) {
        const uint32_t w = (a.rect_max & 0xffff) - (a.rect_min & 0xffff), h = (a.rect_max >> 16) - (a.rect_min >> 16);
        tiles_touched[idx] = a.radius > 0 ? w * h : 0;
    }
    if (clamped) { clamped[3 * idx] = r.flags & 1; clamped[3 * idx + 1] = (r.flags >> 1) & 1; clamped[3 * idx + 2] = (r.flags >> 2) & 1; }
}

// ------------------------------------------------------------------------------------------------
// K8: fused per-Gaussian backward.  backward.cu:144-274 (cov2D), :346-396 (projection, SH, cov3D).
// ------------------------------------------------------------------------------------------------
// `sh` and `dL_dsh` may alias (the staged shared-memory row is overwritten in place: every coefficient is
// read before the first gradient is written)
template <bool VEC>
__device__ __forceinline__ void sh_backward(int deg, int M, const float* sh, float3 pos, float3 campos, uint32_t clamp_bits,
                                            float3 dL_dcolor, float3& dmean_add, float* dL_dsh, int acc_out)
{
    const float3 dorig = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
    const float len = sqrtf(dorig.x * dorig.x + dorig.y * dorig.y + dorig.z * dorig.z);
    const float x = dorig.x / len, y = dorig.y / len, z = dorig.z / len;
    float3 g = dL_dcolor;  // backward.cu:31-34
    if (clamp_bits & 1) g.x = 0;
    if (clamp_bits & 2) g.y = 0;
    if (clamp_bits & 4) g.z = 0;
    float3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
    float w[16];
    float v[48];  // the coefficient row (VEC: read with conflict-free LDS.128, see sh_row_load)
#define SHK(k) make_float3(v[3 * (k)], v[3 * (k) + 1], v[3 * (k) + 2])
#define AXPY(d, a, v) { const float a_ = (a); d.x += a_ * v.x; d.y += a_ * v.y; d.z += a_ * v.z; }
#pragma unroll
    for (int k = 1; k < 16; k++) w[k] = 0.f;
    w[0] = c_SH_C0;
    if (deg > 0) {
        sh_row_load<VEC>(sh, 0, 3, v);
        w[1] = -c_SH_C1 * y; w[2] = c_SH_C1 * z; w[3] = -c_SH_C1 * x;
        const float3 s1 = SHK(1), s2 = SHK(2), s3 = SHK(3);
        AXPY(dx, -c_SH_C1, s3);
        AXPY(dy, -c_SH_C1, s1);
        AXPY(dz, c_SH_C1, s2);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            w[4] = c_SH_C2[0] * xy; w[5] = c_SH_C2[1] * yz; w[6] = c_SH_C2[2] * (2.f * zz - xx - yy);
            w[7] = c_SH_C2[3] * xz; w[8] = c_SH_C2[4] * (xx - yy);
            if (VEC) sh_row_load<VEC>(sh, 3, 7, v);
            else {
#pragma unroll
                for (int i = 12; i < 27; i++) v[i] = sh[i];
            }
            const float3 s4 = SHK(4), s5 = SHK(5), s6 = SHK(6), s7 = SHK(7), s8 = SHK(8);
            AXPY(dx, c_SH_C2[0] * y, s4); AXPY(dx, c_SH_C2[2] * 2.f * -x, s6); AXPY(dx, c_SH_C2[3] * z, s7); AXPY(dx, c_SH_C2[4] * 2.f * x, s8);
            AXPY(dy, c_SH_C2[0] * x, s4); AXPY(dy, c_SH_C2[1] * z, s5); AXPY(dy, c_SH_C2[2] * 2.f * -y, s6); AXPY(dy, c_SH_C2[4] * 2.f * -y, s8);
            AXPY(dz, c_SH_C2[1] * y, s5); AXPY(dz, c_SH_C2[2] * 2.f * 2.f * z, s6); AXPY(dz, c_SH_C2[3] * x, s7);
            if (deg > 2) {
                w[9] = c_SH_C3[0] * y * (3.f * xx - yy); w[10] = c_SH_C3[1] * xy * z; w[11] = c_SH_C3[2] * y * (4.f * zz - xx - yy);
                w[12] = c_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); w[13] = c_SH_C3[4] * x * (4.f * zz - xx - yy);
                w[14] = c_SH_C3[5] * z * (xx - yy); w[15] = c_SH_C3[6] * x * (xx - 3.f * yy);
                if (VEC) sh_row_load<VEC>(sh, 7, 12, v);
                else {
#pragma unroll
                    for (int i = 27; i < 48; i++) v[i] = sh[i];
                }
                const float3 s9 = SHK(9), s10 = SHK(10), s11 = SHK(11), s12 = SHK(12), s13 = SHK(13), s14 = SHK(14), s15 = SHK(15);
                AXPY(dx, c_SH_C3[0] * 3.f * 2.f * xy, s9); AXPY(dx, c_SH_C3[1] * yz, s10); AXPY(dx, c_SH_C3[2] * -2.f * xy, s11);
                AXPY(dx, c_SH_C3[3] * -3.f * 2.f * xz, s12); AXPY(dx, c_SH_C3[4] * (-3.f * xx + 4.f * zz - yy), s13);
                AXPY(dx, c_SH_C3[5] * 2.f * xz, s14); AXPY(dx, c_SH_C3[6] * 3.f * (xx - yy), s15);
                AXPY(dy, c_SH_C3[0] * 3.f * (xx - yy), s9); AXPY(dy, c_SH_C3[1] * xz, s10); AXPY(dy, c_SH_C3[2] * (-3.f * yy + 4.f * zz - xx), s11);
                AXPY(dy, c_SH_C3[3] * -3.f * 2.f * yz, s12); AXPY(dy, c_SH_C3[4] * -2.f * xy, s13); AXPY(dy, c_SH_C3[5] * -2.f * yz, s14);
                AXPY(dy, c_SH_C3[6] * -3.f * 2.f * xy, s15);
                AXPY(dz, c_SH_C3[1] * xy, s10); AXPY(dz, c_SH_C3[2] * 4.f * 2.f * yz, s11); AXPY(dz, c_SH_C3[3] * 3.f * (2.f * zz - xx - yy), s12);
                AXPY(dz, c_SH_C3[4] * 4.f * 2.f * xz, s13); AXPY(dz, c_SH_C3[5] * (xx - yy), s14);
            }
        }
    }
#undef AXPY
#undef SHK
    // dL_dsh[k] = basis_k * dL_dRGB; coefficients above the active degree get zero gradient (w[k] == 0 there)
    if (VEC) {  // staged row, M in {4, 8, 12, 16}: float4 stores (the row was read completely above)
        const float gc[3] = {g.x, g.y, g.z};
#pragma unroll
        for (int q = 0; q < 12; q++) {
            if (4 * q < 3 * M) {
                float4 t;
                t.x = w[(4 * q) / 3] * gc[(4 * q) % 3];
                t.y = w[(4 * q + 1) / 3] * gc[(4 * q + 1) % 3];
                t.z = w[(4 * q + 2) / 3] * gc[(4 * q + 2) % 3];
                t.w = w[(4 * q + 3) / 3] * gc[(4 * q + 3) % 3];
                reinterpret_cast<float4*>(dL_dsh)[q] = t;
            }
        }
    } else {
        const int ncoef = (deg + 1) * (deg + 1);
        for (int k = 0; k < M; k++) {
            float wk = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) wk = (j == k && j < ncoef) ? w[j] : wk;
            if (acc_out == 2) {  // direct-to-global path in atomic accumulate mode
                atomicAdd(dL_dsh + 3 * k, wk * g.x); atomicAdd(dL_dsh + 3 * k + 1, wk * g.y); atomicAdd(dL_dsh + 3 * k + 2, wk * g.z);
            } else if (acc_out) {  // direct-to-global path in accumulate mode
                dL_dsh[3 * k] += wk * g.x; dL_dsh[3 * k + 1] += wk * g.y; dL_dsh[3 * k + 2] += wk * g.z;
            } else {
                dL_dsh[3 * k] = wk * g.x; dL_dsh[3 * k + 1] = wk * g.y; dL_dsh[3 * k + 2] = wk * g.z;
            }
        }
    }
    const float ddx = dx.x * g.x + dx.y * g.y + dx.z * g.z;
    const float ddy = dy.x * g.x + dy.y * g.y + dy.z * g.z;
    const float ddz = dz.x * g.x + dz.y * g.y + dz.z * g.z;
    // auxiliary.h:107-117 dnormvdv
    const float sum2 = dorig.x * dorig.x + dorig.y * dorig.y + dorig.z * dorig.z;
    const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean_add.x = ((+sum2 - dorig.x * dorig.x) * ddx - dorig.y * dorig.x * ddy - dorig.z * dorig.x * ddz) * inv;
    dmean_add.y = (-dorig.x * dorig.y * ddx + (sum2 - dorig.y * dorig.y) * ddy - dorig.z * dorig.y * ddz) * inv;
    dmean_add.z = (-dorig.x * dorig.z * ddx - dorig.y * dorig.z * ddy + (sum2 - dorig.z * dorig.z) * ddz) * inv;
}

__global__ void __launch_bounds__(256, 3) k_preprocess_bwd(PreBwdParams p)
{
    extern __shared__ __align__(16) float s_sh[];
    __shared__ float s_vm[16], s_pm[16];
    if (threadIdx.x < 16) s_vm[threadIdx.x] = p.viewmatrix[threadIdx.x];
    else if (threadIdx.x < 32) s_pm[threadIdx.x - 16] = p.projmatrix[threadIdx.x - 16];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const bool valid = idx < p.P;
    const size_t i = (size_t)(valid ? idx : 0);
    const int M = p.M;
    const float* vm = s_vm;
    const float* proj = s_pm;
    const int radius = valid ? (p.radii ? p.radii[idx] : p.aux[idx].radius) : 0;
    // blend_bwd accumulated the raw moments [S, Sx, Sy, Sxx, Sxy, Syy, cr, cg, cb] of every Gaussian (see blend.cu).
    // A Gaussian no pixel blended (culled, hidden behind the saturated front layer, ...) still has the all-zero row the
    // caller's memset left: all of its gradients are exactly zero, so nothing below needs its parameters or SH row --
    // in accumulate mode it is not touched at all.  On a closed surface that is about half of the Gaussians.
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float m8 = 0.f;
    if (valid && radius > 0) {
        const float4* a4 = reinterpret_cast<const float4*>(p.gacc + i * GSTAR_GACC);
        a0 = a4[0]; a1 = a4[1];
        m8 = p.gacc[i * GSTAR_GACC + 8];
    }
    const bool vis = radius > 0 && (a0.x != 0.f || a0.y != 0.f || a0.z != 0.f || a0.w != 0.f || a1.x != 0.f || a1.y != 0.f || a1.z != 0.f ||
                                    a1.w != 0.f || m8 != 0.f);
    // SH rows of the warp staged in shared memory: coalesced loads now, coalesced dL_dsh stores at the end
    const bool sh_staged = p.shs && p.dL_dsh && M > 0 && M <= SH_MAX_M;
    ShStage st;
    st.rows = nullptr; st.stride = 0; st.vec = false;
    if (sh_staged) {
        st = sh_stage_make(s_sh, p.shs, M);
        const unsigned need = __ballot_sync(0xffffffffu, valid && vis && p.D > 0);
        if (need) sh_stage_load(st, p.shs, M, (long long)idx - lane, p.P, need);
    }
    if (valid) {
    // turn the moments into the reference's blend-stage gradients (backward.cu:523-554) with the per-Gaussian factors
    float acc[9];
    {
        const GRec* rc = p.recs + i;
        const float4 r0 = *reinterpret_cast<const float4*>(rc);       // x y A B
        const float2 r1 = *(reinterpret_cast<const float2*>(rc) + 2);  // C o
        const float A = r0.z, B = r0.w, Cc = r1.x, o = r1.y;
        const float S = a0.x, Sx = a0.y, Sy = a0.z, Sxx = a0.w, Sxy = a1.x, Syy = a1.y;
        acc[0] = -(0.5f * p.W) * (A * Sx + B * Sy);   // dL_dmean2D.x  (ddelx_dx = 0.5 W)
        acc[1] = -(0.5f * p.H) * (Cc * Sy + B * Sx);  // dL_dmean2D.y
        acc[2] = -0.5f * Sxx;                         // dL_dconic.x
        acc[3] = -0.5f * Sxy;                         // dL_dconic.y
        acc[4] = -0.5f * Syy;                         // dL_dconic.w
        acc[5] = a1.z; acc[6] = a1.w; acc[7] = m8;    // dL_dcolor
        acc[8] = (vis && o != 0.f) ? S / o : 0.f;  // dL_dopacity = sum G*dL_dalpha = S / o
    }
    // blend-stage gradients in the reference's layouts (also outputs of the op / parity intermediates)
    p.dL_dmean2D[3 * i] = acc[0]; p.dL_dmean2D[3 * i + 1] = acc[1]; p.dL_dmean2D[3 * i + 2] = 0.f;
    if (p.dL_dconic) { p.dL_dconic[4 * i] = acc[2]; p.dL_dconic[4 * i + 1] = acc[3]; p.dL_dconic[4 * i + 2] = 0.f; p.dL_dconic[4 * i + 3] = acc[4]; }
    if (p.dL_dcolor) { p.dL_dcolor[3 * i] = acc[5]; p.dL_dcolor[3 * i + 1] = acc[6]; p.dL_dcolor[3 * i + 2] = acc[7]; }
    if (p.accumulate == 2) { if (vis) atomicAdd(p.dL_dopacity + i, acc[8]); }
    else if (p.accumulate) { if (vis) p.dL_dopacity[i] += acc[8]; } else p.dL_dopacity[i] = acc[8];
    float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    if (vis) {
        const float mx = p.means3D[3 * i], my = p.means3D[3 * i + 1], mz = p.means3D[3 * i + 2];
        float c6[6];
        float4 q = {0, 0, 0, 0};
        float sc[3] = {0, 0, 0};
        if (p.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = p.cov3D_precomp[6 * i + k];
        } else {
            q = *reinterpret_cast<const float4*>(p.rotations + 4 * i);
            sc[0] = p.scales[3 * i]; sc[1] = p.scales[3 * i + 1]; sc[2] = p.scales[3 * i + 2];
            cov3d_from_scale_rot(sc[0], sc[1], sc[2], p.scale_modifier, q, c6);
        }
        Cov2D cv;
        cov2d_from_cov3d(mx, my, mz, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, c6, vm, cv);
        const float a = cv.a, b = cv.b, c = cv.c;
        const float dcx = acc[2], dcy = acc[3], dcz = acc[4];
        const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
        const float x_grad_mul = (cv.txtz < -limx || cv.txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (cv.tytz < -limy || cv.tytz > limy) ? 0.f : 1.f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float* Tx = cv.Tx; const float* Ty = cv.Ty;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (Tx[0] * Tx[0] * dL_da + Tx[0] * Ty[0] * dL_db + Ty[0] * Ty[0] * dL_dc);
            dcov[3] = (Tx[1] * Tx[1] * dL_da + Tx[1] * Ty[1] * dL_db + Ty[1] * Ty[1] * dL_dc);
            dcov[5] = (Tx[2] * Tx[2] * dL_da + Tx[2] * Ty[2] * dL_db + Ty[2] * Ty[2] * dL_dc);
            dcov[1] = 2 * Tx[0] * Tx[1] * dL_da + (Tx[0] * Ty[1] + Tx[1] * Ty[0]) * dL_db + 2 * Ty[0] * Ty[1] * dL_dc;
            dcov[2] = 2 * Tx[0] * Tx[2] * dL_da + (Tx[0] * Ty[2] + Tx[2] * Ty[0]) * dL_db + 2 * Ty[0] * Ty[2] * dL_dc;
            dcov[4] = 2 * Tx[2] * Tx[1] * dL_da + (Tx[1] * Ty[2] + Tx[2] * Ty[1]) * dL_db + 2 * Ty[1] * Ty[2] * dL_dc;
        }
        // dL/dT (backward.cu:237-248):  T.V rows are exactly the forward's Ax, Ay
        const float dT00 = 2 * cv.Ax[0] * dL_da + cv.Ay[0] * dL_db, dT01 = 2 * cv.Ax[1] * dL_da + cv.Ay[1] * dL_db,
                    dT02 = 2 * cv.Ax[2] * dL_da + cv.Ay[2] * dL_db;
        const float dT10 = 2 * cv.Ay[0] * dL_dc + cv.Ax[0] * dL_db, dT11 = 2 * cv.Ay[1] * dL_dc + cv.Ax[1] * dL_db,
                    dT12 = 2 * cv.Ay[2] * dL_dc + cv.Ax[2] * dL_db;
        const float dJ00 = vm[0] * dT00 + vm[4] * dT01 + vm[8] * dT02;
        const float dJ02 = vm[2] * dT00 + vm[6] * dT01 + vm[10] * dT02;
        const float dJ11 = vm[1] * dT10 + vm[5] * dT11 + vm[9] * dT12;
        const float dJ12 = vm[2] * dT10 + vm[6] * dT11 + vm[10] * dT12;
        const float itz = 1.f / cv.tz, tz2 = itz * itz, tz3 = tz2 * itz;
        const float h_x = p.focal_x, h_y = p.focal_y;
        const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
        const float dty = y_grad_mul * -h_y * tz2 * dJ12;
        const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * cv.tx) * tz3 * dJ02 + (2 * h_y * cv.ty) * tz3 * dJ12;
        dmean[0] = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;  // auxiliary.h:89-97
        dmean[1] = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
        dmean[2] = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;
        // projection part (backward.cu:370-387)
        const float m_hw = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
        const float m_w = 1.0f / (m_hw + 0.0000001f);
        const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        const float g2x = acc[0], g2y = acc[1];
        dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        if (p.shs) {
            float3 add;
            const float* sh_in = sh_staged ? st.rows + lane * st.stride : p.shs + (size_t)3 * M * i;
            float* sh_out = sh_staged ? st.rows + lane * st.stride : p.dL_dsh + (size_t)3 * M * i;
            if (sh_staged && st.vec)
                sh_backward<true>(p.D, M, sh_in, make_float3(mx, my, mz), make_float3(p.campos[0], p.campos[1], p.campos[2]), p.recs[idx].flags,
                                  make_float3(acc[5], acc[6], acc[7]), add, sh_out, false);
            else
                sh_backward<false>(p.D, M, sh_in, make_float3(mx, my, mz), make_float3(p.campos[0], p.campos[1], p.campos[2]), p.recs[idx].flags,
                                   make_float3(acc[5], acc[6], acc[7]), add, sh_out, sh_staged ? 0 : p.accumulate);
            dmean[0] += add.x; dmean[1] += add.y; dmean[2] += add.z;
        }
        if (p.scales) {
            // backward.cu:278-341.  M_glm[c][r] = s_r * R_std(c,r);  dL_dM = 2 M dSigma
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                   {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                   {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};  // R[c][r] glm == R_std(c,r)
            const float s[3] = {p.scale_modifier * sc[0], p.scale_modifier * sc[1], p.scale_modifier * sc[2]};
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]}, {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3];  // glm [c][r]
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++) {
                    const float m0 = 2.0f * (s[rr] * R[0][rr]), m1 = 2.0f * (s[rr] * R[1][rr]), m2 = 2.0f * (s[rr] * R[2][rr]);
                    dM[cc][rr] = m0 * dS[cc][0] + m1 * dS[cc][1] + m2 * dS[cc][2];
                }
            // Rt[k][j] = R[j][k]; dMt[k][j] = dM[j][k]
#pragma unroll
            for (int k = 0; k < 3; k++) dscale[k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
            float D[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) D[k][j] = dM[j][k] * s[k];
            drot[0] = 2 * z * (D[0][1] - D[1][0]) + 2 * y * (D[2][0] - D[0][2]) + 2 * x * (D[1][2] - D[2][1]);
            drot[1] = 2 * y * (D[1][0] + D[0][1]) + 2 * z * (D[2][0] + D[0][2]) + 2 * r * (D[1][2] - D[2][1]) - 4 * x * (D[2][2] + D[1][1]);
            drot[2] = 2 * x * (D[1][0] + D[0][1]) + 2 * r * (D[2][0] - D[0][2]) + 2 * z * (D[1][2] + D[2][1]) - 4 * y * (D[2][2] + D[0][0]);
            drot[3] = 2 * r * (D[0][1] - D[1][0]) + 2 * x * (D[2][0] + D[0][2]) + 2 * y * (D[1][2] + D[2][1]) - 4 * z * (D[1][1] + D[0][0]);
        }
    } else if (p.dL_dsh && !p.accumulate) {
        float* d = sh_staged ? st.rows + lane * st.stride : p.dL_dsh + (size_t)3 * M * i;
        for (int k = 0; k < 3 * M; k++) d[k] = 0.f;
    }
    const bool acc_mode = p.accumulate != 0;
    if (!acc_mode) {
        p.dL_dmean3D[3 * i] = dmean[0]; p.dL_dmean3D[3 * i + 1] = dmean[1]; p.dL_dmean3D[3 * i + 2] = dmean[2];
        if (p.dL_dscale) { p.dL_dscale[3 * i] = dscale[0]; p.dL_dscale[3 * i + 1] = dscale[1]; p.dL_dscale[3 * i + 2] = dscale[2]; }
        if (p.dL_drot) *reinterpret_cast<float4*>(p.dL_drot + 4 * i) = make_float4(drot[0], drot[1], drot[2], drot[3]);
    } else if (vis && p.accumulate == 2) {  // several backward calls may be adding into these arrays at once
        atomicAdd(p.dL_dmean3D + 3 * i, dmean[0]); atomicAdd(p.dL_dmean3D + 3 * i + 1, dmean[1]); atomicAdd(p.dL_dmean3D + 3 * i + 2, dmean[2]);
        if (p.dL_dscale) { atomicAdd(p.dL_dscale + 3 * i, dscale[0]); atomicAdd(p.dL_dscale + 3 * i + 1, dscale[1]); atomicAdd(p.dL_dscale + 3 * i + 2, dscale[2]); }
        if (p.dL_drot) red_add_v4(p.dL_drot + 4 * i, drot[0], drot[1], drot[2], drot[3]);
    } else if (vis) {  // accumulate: invisible Gaussians contribute nothing and are not touched
        p.dL_dmean3D[3 * i] += dmean[0]; p.dL_dmean3D[3 * i + 1] += dmean[1]; p.dL_dmean3D[3 * i + 2] += dmean[2];
        if (p.dL_dscale) { p.dL_dscale[3 * i] += dscale[0]; p.dL_dscale[3 * i + 1] += dscale[1]; p.dL_dscale[3 * i + 2] += dscale[2]; }
        if (p.dL_drot) {
            float4* d4 = reinterpret_cast<float4*>(p.dL_drot + 4 * i);
            float4 o = *d4;
            o.x += drot[0]; o.y += drot[1]; o.z += drot[2]; o.w += drot[3];
            *d4 = o;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (p.dL_dcov3D) p.dL_dcov3D[6 * i + k] = dcov[k];
    }  // valid
    if (sh_staged) {
        const unsigned rows = p.accumulate ? __ballot_sync(0xffffffffu, valid && vis) : 0xffffffffu;
        sh_stage_store(st, p.dL_dsh, M, (long long)idx - lane, p.P, rows, p.accumulate);
    }
}

int preprocess_setup()
{
    cudaError_t e = cudaFuncSetAttribute(k_preprocess_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFuncSetAttribute(k_preprocess_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES);
}
void launch_preprocess_fwd(const PreFwdParams& p, cudaStream_t s)
{
    const int smem = (p.shs && !p.colors_precomp) ? SH_SMEM_BYTES : 0;
    k_preprocess_fwd<<<(p.P + 255) / 256, 256, smem, s>>>(p);
}
void launch_preprocess_bwd(const PreBwdParams& p, cudaStream_t s)
{
    const int smem = (p.shs && p.dL_dsh) ? SH_SMEM_BYTES : 0;
    k_preprocess_bwd<<<(p.P + 255) / 256, 256, smem, s>>>(p);
}
void launch_mark_visible(int P, const float* means3D, const float* vm, unsigned char* present, cudaStream_t s)
{
    k_mark_visible<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, vm, present);
}
void launch_geom_unpack(const GRec* recs, const GAux* aux, int P, float* depths, float* means2D, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                        unsigned char* clamped, cudaStream_t s)
{
    k_geom_unpack<<<(P + 255) / 256, 256, 0, s>>>(recs, aux, P, depths, means2D, conic_opacity, rgb, tiles_touched, clamped);
}

}  // namespace gstar
