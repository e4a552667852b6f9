// sugar_prologue.cu -- SURVEY 8f-4: the caller-side per-Gaussian prologue of a mesh-bound SuGaR model as ONE kernel pair.
//
// What gaustar_scene/sugar_model.py computes with a few dozen torch kernels on every render call (and differentiates with as
// many again): `points` (:417-435, barycentric combination of the face's vertices), `scaling` (:457-476, thickness | exp of the
// two learnt in-plane scales, optional clamps), `quaternions` (:478-508: face normal | first edge | their cross product as a
// frame, rotated in the triangle's plane by the learnt complex number, turned into a quaternion by pytorch3d's
// matrix_to_quaternion and normalised) and `strengths` (:443-447, sigmoid).  One thread per Gaussian; the backward pushes the
// four upstream gradients back to the mesh vertices (atomics: a vertex is shared by ~36 Gaussians), the scale and complex
// parameters and the densities.  Not on the reference's operator boundary: reached through gaustar_b200/sugar.py, which can
// patch a SuGaR instance's properties (nothing in GauSTAR is edited).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gstar_kernels.h"

namespace gstar {

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ f3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
// torch.nn.functional.normalize(x, eps): x / max(|x|, eps)
__device__ __forceinline__ f3 normalize_eps(f3 a, float eps, float& len)
{
    len = fmaxf(sqrtf(dot(a, a)), eps);
    return (1.0f / len) * a;
}
// backward of y = x / |x| (|x| above eps): g_x = (g_y - y (y . g_y)) / |x|
__device__ __forceinline__ f3 normalize_bwd(f3 y, float len, f3 gy) { return (1.0f / len) * (gy - dot(y, gy) * y); }

struct Frame {
    f3 v0, v1, v2, p, n, R0, u1, u2, R1, R2;
    float len_nraw, len_n, len_a, len_w, len_c, c0, c1;
};

__device__ __forceinline__ void face_vertices(const SugarParams& s, int f, int& i0, int& i1, int& i2)
{
    if (s.faces64) { i0 = (int)s.faces64[3 * f]; i1 = (int)s.faces64[3 * f + 1]; i2 = (int)s.faces64[3 * f + 2]; }
    else { i0 = s.faces32[3 * f]; i1 = s.faces32[3 * f + 1]; i2 = s.faces32[3 * f + 2]; }
}

__device__ __forceinline__ Frame make_frame(const SugarParams& s, int g, int i0, int i1, int i2)
{
    Frame F;
    const int k = g % s.K;
    F.v0 = ld3(s.verts + 3 * (size_t)i0); F.v1 = ld3(s.verts + 3 * (size_t)i1); F.v2 = ld3(s.verts + 3 * (size_t)i2);
    const float b0 = s.bary[3 * k], b1 = s.bary[3 * k + 1], b2 = s.bary[3 * k + 2];
    F.p = b0 * F.v0 + b1 * F.v1 + b2 * F.v2;                          // sugar_model.py:426-435
    const f3 nraw = cross(F.v1 - F.v0, F.v2 - F.v0);                   // pytorch3d Meshes.faces_normals_packed
    F.n = normalize_eps(nraw, 1e-6f, F.len_nraw);
    F.R0 = normalize_eps(F.n, 1e-12f, F.len_n);                        // :484
    F.u1 = normalize_eps(F.v0 - F.v1, 1e-12f, F.len_a);               // :487-488
    F.u2 = normalize_eps(cross(F.R0, F.u1), 1e-12f, F.len_w);          // :491
    const float q0 = s.cplx[2 * (size_t)g], q1 = s.cplx[2 * (size_t)g + 1];
    F.len_c = fmaxf(sqrtf(q0 * q0 + q1 * q1), 1e-12f);                 // :494
    F.c0 = q0 / F.len_c; F.c1 = q1 / F.len_c;
    F.R1 = F.c0 * F.u1 + F.c1 * F.u2;                                  // :495-496
    F.R2 = F.c0 * F.u2 - F.c1 * F.u1;
    return F;
}

// pytorch3d.transforms.matrix_to_quaternion (rotation_conversions.py, 0.7.4): four candidates, the best-conditioned one.
// Columns of the matrix: R0 R1 R2.  Returns the pivot index; t = the pivot's 1 +- m00 +- m11 +- m22.
__device__ __forceinline__ int mat_to_quat(const Frame& F, float q[4], float& qa)
{
    const float m00 = F.R0.x, m10 = F.R0.y, m20 = F.R0.z, m01 = F.R1.x, m11 = F.R1.y, m21 = F.R1.z, m02 = F.R2.x, m12 = F.R2.y, m22 = F.R2.z;
    const float t[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    float a[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = t[i] > 0.0f ? sqrtf(t[i]) : 0.0f; if (a[i] > a[best]) best = i; }
    qa = a[best];
    const float d = 1.0f / (2.0f * fmaxf(qa, 0.1f));
    float r[4];
    if (best == 0) { r[0] = qa * qa; r[1] = m21 - m12; r[2] = m02 - m20; r[3] = m10 - m01; }
    else if (best == 1) { r[0] = m21 - m12; r[1] = qa * qa; r[2] = m10 + m01; r[3] = m02 + m20; }
    else if (best == 2) { r[0] = m02 - m20; r[1] = m10 + m01; r[2] = qa * qa; r[3] = m12 + m21; }
    else { r[0] = m10 - m01; r[1] = m20 + m02; r[2] = m21 + m12; r[3] = qa * qa; }
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = r[i] * d;
    return best;
}

__global__ void __launch_bounds__(256) k_sugar_prologue_fwd(SugarParams s)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= s.P) return;
    int i0, i1, i2;
    face_vertices(s, g / s.K, i0, i1, i2);
    const Frame F = make_frame(s, g, i0, i1, i2);
    float q[4], qa;
    mat_to_quat(F, q, qa);
    const float ql = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);  // :508
    float* o = s.points + 3 * (size_t)g;
    o[0] = F.p.x; o[1] = F.p.y; o[2] = F.p.z;
    float sa = expf(s.scales[2 * (size_t)g]), sb = expf(s.scales[2 * (size_t)g + 1]);              // :462-466
    if (s.has_max) { sa = fminf(sa, s.max_scale); sb = fminf(sb, s.max_scale); }
    if (s.has_min) { sa = fmaxf(sa, s.min_scale); sb = fmaxf(sb, s.min_scale); }
    o = s.scaling + 3 * (size_t)g;
    o[0] = s.thickness; o[1] = sa; o[2] = sb;
    reinterpret_cast<float4*>(s.quats)[g] = make_float4(q[0] / ql, q[1] / ql, q[2] / ql, q[3] / ql);
    s.opac[g] = 1.0f / (1.0f + expf(-s.dens[g]));                                                  // :447
}

__global__ void __launch_bounds__(256) k_sugar_prologue_bwd(SugarParams s)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= s.P) return;
    int i0, i1, i2;
    face_vertices(s, g / s.K, i0, i1, i2);
    const Frame F = make_frame(s, g, i0, i1, i2);
    const int k = g % s.K;
    // ---- strengths, scaling
    if (s.d_dens) {
        const float o = 1.0f / (1.0f + expf(-s.dens[g]));
        s.d_dens[g] = (s.g_opac ? s.g_opac[g] : 0.0f) * o * (1.0f - o);
    }
    if (s.d_scales) {
        float ga = 0.f, gb = 0.f;
        if (s.g_scaling) {
            const float ea = expf(s.scales[2 * (size_t)g]), eb = expf(s.scales[2 * (size_t)g + 1]);
            const bool ca = (s.has_max && ea > s.max_scale) || (s.has_min && fminf(ea, s.has_max ? s.max_scale : ea) < s.min_scale);
            const bool cb = (s.has_max && eb > s.max_scale) || (s.has_min && fminf(eb, s.has_max ? s.max_scale : eb) < s.min_scale);
            ga = ca ? 0.f : s.g_scaling[3 * (size_t)g + 1] * ea;
            gb = cb ? 0.f : s.g_scaling[3 * (size_t)g + 2] * eb;
        }
        s.d_scales[2 * (size_t)g] = ga; s.d_scales[2 * (size_t)g + 1] = gb;
    }
    // ---- quaternion -> rotation columns
    f3 gR0 = {0, 0, 0}, gR1 = {0, 0, 0}, gR2 = {0, 0, 0};
    if (s.g_quats) {
        float q[4], qa;
        const int best = mat_to_quat(F, q, qa);
        const float ql = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
        const float4 gq4 = reinterpret_cast<const float4*>(s.g_quats)[g];
        const float gy[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
        float yn[4], gq[4], yd = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) { yn[i] = q[i] / ql; yd += yn[i] * gy[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) gq[i] = (gy[i] - yn[i] * yd) / ql;   // through the final normalisation
        // q_j = r_j / (2 qa) for j != best, q_best = qa / 2 (qa >= 1 for a rotation matrix: the 0.1 floor is inactive)
        const float d = 1.0f / (2.0f * fmaxf(qa, 0.1f));
        float gr[4], gqa = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i == best) { gr[i] = 0.f; gqa += 0.5f * gq[i]; }
            else { gr[i] = gq[i] * d; gqa -= gq[i] * q[i] / fmaxf(qa, 0.1f); }
        }
        const float gt = qa > 0.f ? gqa / (2.0f * qa) : 0.f;  // qa = sqrt(t)
        float gm00 = 0, gm01 = 0, gm02 = 0, gm10 = 0, gm11 = 0, gm12 = 0, gm20 = 0, gm21 = 0, gm22 = 0;
        if (best == 0) { gm00 += gt; gm11 += gt; gm22 += gt; gm21 += gr[1]; gm12 -= gr[1]; gm02 += gr[2]; gm20 -= gr[2]; gm10 += gr[3]; gm01 -= gr[3]; }
        else if (best == 1) { gm00 += gt; gm11 -= gt; gm22 -= gt; gm21 += gr[0]; gm12 -= gr[0]; gm10 += gr[2]; gm01 += gr[2]; gm02 += gr[3]; gm20 += gr[3]; }
        else if (best == 2) { gm00 -= gt; gm11 += gt; gm22 -= gt; gm02 += gr[0]; gm20 -= gr[0]; gm10 += gr[1]; gm01 += gr[1]; gm12 += gr[3]; gm21 += gr[3]; }
        else { gm00 -= gt; gm11 -= gt; gm22 += gt; gm10 += gr[0]; gm01 -= gr[0]; gm20 += gr[1]; gm02 += gr[1]; gm21 += gr[2]; gm12 += gr[2]; }
        gR0 = {gm00, gm10, gm20}; gR1 = {gm01, gm11, gm21}; gR2 = {gm02, gm12, gm22};
    }
    // ---- R1 = c0 u1 + c1 u2, R2 = c0 u2 - c1 u1
    const float gc0 = dot(gR1, F.u1) + dot(gR2, F.u2), gc1 = dot(gR1, F.u2) - dot(gR2, F.u1);
    f3 gu1 = F.c0 * gR1 - F.c1 * gR2, gu2 = F.c1 * gR1 + F.c0 * gR2;
    if (s.d_cplx) {
        const float yd = F.c0 * gc0 + F.c1 * gc1;
        s.d_cplx[2 * (size_t)g] = (gc0 - F.c0 * yd) / F.len_c;
        s.d_cplx[2 * (size_t)g + 1] = (gc1 - F.c1 * yd) / F.len_c;
    }
    if (!s.d_verts) return;
    // ---- u2 = normalize(R0 x u1), u1 = normalize(v0 - v1), R0 = normalize(normalize((v1 - v0) x (v2 - v0)))
    const f3 gw = normalize_bwd(F.u2, F.len_w, gu2);
    gR0 = gR0 + cross(F.u1, gw);
    gu1 = gu1 + cross(gw, F.R0);
    const f3 ga = normalize_bwd(F.u1, F.len_a, gu1);
    const f3 gn = normalize_bwd(F.R0, F.len_n, gR0);
    const f3 gnraw = normalize_bwd(F.n, F.len_nraw, gn);
    const f3 e1 = F.v1 - F.v0, e2 = F.v2 - F.v0;
    const f3 ge1 = cross(e2, gnraw), ge2 = cross(gnraw, e1);
    f3 gp = {0, 0, 0};
    if (s.g_points) gp = ld3(s.g_points + 3 * (size_t)g);
    const float b0 = s.bary[3 * k], b1 = s.bary[3 * k + 1], b2 = s.bary[3 * k + 2];
    const f3 gv0 = b0 * gp + ga - ge1 - ge2, gv1 = b1 * gp - ga + ge1, gv2 = b2 * gp + ge2;
    float* d0 = s.d_verts + 3 * (size_t)i0; float* d1 = s.d_verts + 3 * (size_t)i1; float* d2 = s.d_verts + 3 * (size_t)i2;
    atomicAdd(d0, gv0.x); atomicAdd(d0 + 1, gv0.y); atomicAdd(d0 + 2, gv0.z);
    atomicAdd(d1, gv1.x); atomicAdd(d1 + 1, gv1.y); atomicAdd(d1 + 2, gv1.z);
    atomicAdd(d2, gv2.x); atomicAdd(d2 + 1, gv2.y); atomicAdd(d2 + 2, gv2.z);
}

void launch_sugar_prologue_fwd(const SugarParams& s, cudaStream_t st) { k_sugar_prologue_fwd<<<(s.P + 255) / 256, 256, 0, st>>>(s); }
void launch_sugar_prologue_bwd(const SugarParams& s, cudaStream_t st) { k_sugar_prologue_bwd<<<(s.P + 255) / 256, 256, 0, st>>>(s); }

}  // namespace gstar
