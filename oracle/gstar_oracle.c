/*
 * gstar_oracle.c -- CPU restatement of the diff-gaussian-rasterization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gaustar_b200/, the
 * C-ABI library, the torch shim) may include, link, import or execute this
 * file.  It is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py)
 * against tests/golden/*.npz, which hold outputs of the UNMODIFIED reference
 * CUDA rasterizer (oracle/_ref, built from /root/reference by
 * oracle/build_ref.py) executed on a B200 by tests/golden/make_golden.py.
 *
 * DGR = gaussian_splatting/submodules/diff-gaussian-rasterization (reference).
 * Every function cites the reference file:line it restates.  Arithmetic is
 * IEEE fp32.  The reference is compiled by nvcc with -fmad=true, and which
 * mul/add pairs become FMAs is decided by NVVM *and* ptxas, so the placement
 * of fmaf() below follows the SASS of the reference kernels compiled for
 * sm_100a with nvcc 12.9 (dumped with cuobjdump; see DESIGN.md "bit-exact
 * keys").  Compile with -ffp-contract=off so gcc adds no contraction of its
 * own:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC.
 *
 * Only the key-determining chain (depth, radius, pixel centre, tile rect,
 * conic, and the per-pixel alpha/T thresholds) is transcribed FMA-exactly;
 * colours and all gradients are written as plain fp32 expressions (gradients
 * accumulate in fp64) and are compared to a tolerance, because the reference
 * itself sums them with order-nondeterministic atomics.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BLOCK_X 16 /* DGR/cuda_rasterizer/config.h:16 */
#define ORC_BLOCK_Y 16 /* DGR/cuda_rasterizer/config.h:17 */

/* DGR/cuda_rasterizer/auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

typedef struct {
    int P, D, M, W, H;
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    float scale_modifier;
    const float* viewmatrix; /* [16] flat, see auxiliary.h:58-77 */
    const float* projmatrix; /* [16] */
    const float* campos;     /* [3] */
    const float* bg;         /* [3] */
    float tan_fovx, tan_fovy;
} orc_scene;

/* auxiliary.h:58-77 rows; SASS: fma(z,m8,fma(x,m0,y*m4)) + m12 */
static inline float xform_row(const float* m, int r, float x, float y, float z)
{
    return fmaf(z, m[8 + r], fmaf(x, m[r], y * m[4 + r])) + m[12 + r];
}

/* DGR/rasterizer_impl.cu:35-50 */
uint32_t orc_higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb)
            msb += step;
        else
            msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* forward.cu:118-152 computeCov3D (quaternion used as given, :127) */
static void cov3d_from_scale_rot(const float* s3, float mod, const float* q4, float* cov6)
{
    const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
    const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
    /* R_std entries; contraction as in the sm_100a SASS */
    const float yy = y * y, zz = z * z;
    const float R00 = 1.0f - 2.0f * (yy + zz);
    const float R01 = 2.0f * fmaf(x, y, -(r * z));
    const float R02 = 2.0f * fmaf(r, y, x * z);
    const float R10 = 2.0f * fmaf(x, y, r * z);
    const float R11 = 1.0f - 2.0f * fmaf(x, x, zz);
    const float R12 = 2.0f * fmaf(y, z, -(r * x));
    const float R20 = 2.0f * fmaf(-r, y, x * z);
    const float R21 = 2.0f * fmaf(y, z, r * x);
    const float R22 = 1.0f - 2.0f * fmaf(x, x, yy);
    /* M = S*R (glm): m_ak = s_k * R_ak */
    const float m00 = sx * R00, m01 = sy * R01, m02 = sz * R02;
    const float m10 = sx * R10, m11 = sy * R11, m12 = sz * R12;
    const float m20 = sx * R20, m21 = sy * R21, m22 = sz * R22;
    /* Sigma = M^T M : fma(m_a2,m_b2, fma(m_a0,m_b0, m_a1*m_b1)) */
    cov6[0] = fmaf(m02, m02, fmaf(m00, m00, m01 * m01));
    cov6[1] = fmaf(m12, m02, fmaf(m10, m00, m11 * m01));
    cov6[2] = fmaf(m22, m02, fmaf(m20, m00, m21 * m01));
    cov6[3] = fmaf(m12, m12, fmaf(m10, m10, m11 * m11));
    cov6[4] = fmaf(m22, m12, fmaf(m20, m10, m21 * m11));
    cov6[5] = fmaf(m22, m22, fmaf(m20, m20, m21 * m21));
}

/* forward.cu:74-113 computeCov2D; also returns T rows (glm T[0][*], T[1][*]) and A rows */
typedef struct {
    float Tx[3], Ty[3], Ax[3], Ay[3];
    float a, b, c; /* cov2D incl. +0.3 */
    float tx, ty, tz, txtz, tytz;
} orc_cov2d;

static void cov2d_from_cov3d(float px, float py, float pz, float fx, float fy, float tan_fovx, float tan_fovy,
                             const float* c6, const float* vm, orc_cov2d* o)
{
    float tx = xform_row(vm, 0, px, py, pz);
    float ty = xform_row(vm, 1, px, py, pz);
    const float tz = xform_row(vm, 2, px, py, pz);
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float J00 = fx / tz, J11 = fy / tz;
    const float J02 = -(fx * tx) / (tz * tz);
    const float J12 = -(fy * ty) / (tz * tz);
    for (int k = 0; k < 3; k++) {
        o->Tx[k] = fmaf(vm[2 + 4 * k], J02, vm[4 * k] * J00);
        o->Ty[k] = fmaf(vm[2 + 4 * k], J12, vm[1 + 4 * k] * J11);
    }
    /* symmetric Vrk columns */
    const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    for (int k = 0; k < 3; k++) {
        o->Ax[k] = fmaf(o->Tx[2], V[2][k], fmaf(o->Tx[0], V[0][k], o->Tx[1] * V[1][k]));
        o->Ay[k] = fmaf(o->Ty[2], V[2][k], fmaf(o->Ty[0], V[0][k], o->Ty[1] * V[1][k]));
    }
    const float c00 = fmaf(o->Tx[2], o->Ax[2], fmaf(o->Tx[0], o->Ax[0], o->Tx[1] * o->Ax[1]));
    const float c01 = fmaf(o->Tx[2], o->Ay[2], fmaf(o->Tx[0], o->Ay[0], o->Tx[1] * o->Ay[1]));
    const float c11 = fmaf(o->Ty[2], o->Ay[2], fmaf(o->Ty[0], o->Ay[0], o->Ty[1] * o->Ay[1]));
    o->a = c00 + 0.3f;
    o->b = c01;
    o->c = c11 + 0.3f;
    o->tx = tx; o->ty = ty; o->tz = tz; o->txtz = txtz; o->tytz = tytz;
}

/* auxiliary.h:41-44 ndc2Pix, evaluated in double (SASS: DADD, DFMA, DMUL) */
static inline float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

/* auxiliary.h:46-56 getRect */
static void get_rect(float pxf, float pyf, int radius, int gx, int gy, uint32_t* rmin, uint32_t* rmax)
{
    const float r = (float)radius;
    int v;
    v = (int)((pxf - r) * 0.0625f); v = v < 0 ? 0 : v; rmin[0] = (uint32_t)v < (uint32_t)gx ? (uint32_t)v : (uint32_t)gx;
    v = (int)((pyf - r) * 0.0625f); v = v < 0 ? 0 : v; rmin[1] = (uint32_t)v < (uint32_t)gy ? (uint32_t)v : (uint32_t)gy;
    v = (int)((((pxf + r) + 16.0f) - 1.0f) * 0.0625f); v = v < 0 ? 0 : v; rmax[0] = (uint32_t)v < (uint32_t)gx ? (uint32_t)v : (uint32_t)gx;
    v = (int)((((pyf + r) + 16.0f) - 1.0f) * 0.0625f); v = v < 0 ? 0 : v; rmax[1] = (uint32_t)v < (uint32_t)gy ? (uint32_t)v : (uint32_t)gy;
}

/* forward.cu:20-71 computeColorFromSH */
static void sh_to_rgb(int deg, int M, const float* mean, const float* campos, const float* sh /*[M][3]*/, float* rgb,
                      uint8_t* clamped)
{
    float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
    const float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    const float x = dx / len, y = dy / len, z = dz / len;
    float res[3];
    (void)M;
    for (int c = 0; c < 3; c++) {
        float v = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            v = v - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                v = v + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    v = v + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] + SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] +
                        SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                }
            }
        }
        v += 0.5f;
        clamped[c] = (v < 0);
        res[c] = v > 0.0f ? v : 0.0f;
    }
    rgb[0] = res[0]; rgb[1] = res[1]; rgb[2] = res[2];
}

/* forward.cu:155-256 preprocessCUDA.  Outputs sized P (zero-initialised here like
 * rasterize_points.cu:67 does for radii; untouched entries of the float arrays stay 0). */
void orc_preprocess(const orc_scene* s, float* depths, int32_t* radii, float* means2D, float* cov3D, float* conic_opacity,
                    float* rgb, uint8_t* clamped, uint32_t* tiles_touched)
{
    const int P = s->P;
    const float focal_y = s->H / (2.0f * s->tan_fovy); /* rasterizer_impl.cu:222-223 */
    const float focal_x = s->W / (2.0f * s->tan_fovx);
    const int gx = (s->W + ORC_BLOCK_X - 1) / ORC_BLOCK_X, gy = (s->H + ORC_BLOCK_Y - 1) / ORC_BLOCK_Y;
    const float* vm = s->viewmatrix;
    const float* pm = s->projmatrix;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        depths[i] = 0.f;
        means2D[2 * i] = means2D[2 * i + 1] = 0.f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.f;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }
        const float px = s->means3D[3 * i], py = s->means3D[3 * i + 1], pz = s->means3D[3 * i + 2];
        /* auxiliary.h:139-164 in_frustum: only the near test survives */
        const float pvz = xform_row(vm, 2, px, py, pz);
        if (pvz <= 0.2f) continue;
        const float hx = xform_row(pm, 0, px, py, pz);
        const float hy = xform_row(pm, 1, px, py, pz);
        const float hw = xform_row(pm, 3, px, py, pz);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;
        float c6[6];
        if (s->cov3D_precomp) {
            for (int k = 0; k < 6; k++) c6[k] = s->cov3D_precomp[6 * i + k];
        } else {
            cov3d_from_scale_rot(s->scales + 3 * i, s->scale_modifier, s->rotations + 4 * i, c6);
            for (int k = 0; k < 6; k++) cov3D[6 * i + k] = c6[k];
        }
        orc_cov2d cv;
        cov2d_from_cov3d(px, py, pz, focal_x, focal_y, s->tan_fovx, s->tan_fovy, c6, vm, &cv);
        const float det = fmaf(cv.a, cv.c, -(cv.b * cv.b)); /* forward.cu:219, SASS FFMA */
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conx = cv.c * det_inv, cony = -cv.b * det_inv, conz = cv.a * det_inv;
        const float mid = 0.5f * (cv.a + cv.c);
        const float disc = fmaxf(0.1f, fmaf(mid, mid, -det)); /* forward.cu:230 */
        const float sq = sqrtf(disc);
        const float lambda1 = mid + sq, lambda2 = mid - sq;
        const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        const float pix = ndc2pix(projx, s->W), piy = ndc2pix(projy, s->H);
        uint32_t rmin[2], rmax[2];
        get_rect(pix, piy, (int)my_radius, gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (!s->colors_precomp) sh_to_rgb(s->D, s->M, s->means3D + 3 * i, s->campos, s->shs + (size_t)3 * s->M * i, rgb + 3 * i, clamped + 3 * i);
        depths[i] = pvz;
        radii[i] = (int32_t)my_radius;
        means2D[2 * i] = pix;
        means2D[2 * i + 1] = piy;
        conic_opacity[4 * i + 0] = conx;
        conic_opacity[4 * i + 1] = cony;
        conic_opacity[4 * i + 2] = conz;
        conic_opacity[4 * i + 3] = s->opacities[i];
        tiles_touched[i] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
    }
}

/* rasterizer_impl.cu:54-66 checkFrustum / markVisible */
void orc_mark_visible(int P, const float* means3D, const float* vm, uint8_t* present)
{
    for (int i = 0; i < P; i++) present[i] = xform_row(vm, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) > 0.2f;
}

/* rasterizer_impl.cu:277 InclusiveSum; returns num_rendered (:281) */
int64_t orc_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets)
{
    uint32_t acc = 0;
    for (int i = 0; i < P; i++) { acc += tiles_touched[i]; offsets[i] = acc; }
    return P ? (int64_t)acc : 0;
}

/* rasterizer_impl.cu:70-111 duplicateWithKeys */
void orc_duplicate_with_keys(int P, int W, int H, const float* means2D, const float* depths, const uint32_t* offsets,
                             const int32_t* radii, uint64_t* keys, uint32_t* values)
{
    const int gx = (W + ORC_BLOCK_X - 1) / ORC_BLOCK_X, gy = (H + ORC_BLOCK_Y - 1) / ORC_BLOCK_Y;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t off = (i == 0) ? 0 : offsets[i - 1];
        uint32_t rmin[2], rmax[2];
        get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, rmin, rmax);
        uint32_t dbits;
        memcpy(&dbits, &depths[i], 4);
        for (uint32_t y = rmin[1]; y < rmax[1]; y++)
            for (uint32_t x = rmin[0]; x < rmax[0]; x++) {
                uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
                key <<= 32;
                key |= dbits;
                keys[off] = key;
                values[off] = (uint32_t)i;
                off++;
            }
    }
}

/* rasterizer_impl.cu:303-308: stable LSD radix sort (cub::DeviceRadixSort::SortPairs,
 * CUDA toolkit CCCL; integer, stable, version independent) on bits [0, end_bit). */
void orc_sort_pairs(int64_t n, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out, int end_bit)
{
    uint64_t* ka = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
    uint32_t* va = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
    uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    memcpy(ka, keys_in, sizeof(uint64_t) * (size_t)n);
    memcpy(va, vals_in, sizeof(uint32_t) * (size_t)n);
    for (int shift = 0; shift < end_bit; shift += 8) {
        const int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        const uint64_t mask = ((uint64_t)1 << bits) - 1;
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (int64_t i = 0; i < n; i++) hist[((ka[i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (int64_t i = 0; i < n; i++) {
            const size_t p = hist[(ka[i] >> shift) & mask]++;
            kb[p] = ka[i];
            vb[p] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    memcpy(keys_out, ka, sizeof(uint64_t) * (size_t)n);
    memcpy(vals_out, va, sizeof(uint32_t) * (size_t)n);
    free(ka); free(va); free(kb); free(vb);
}

/* rasterizer_impl.cu:310 (memset) + :116-138 identifyTileRanges. ranges [T][2] */
void orc_tile_ranges(int64_t L, const uint64_t* sorted_keys, int num_tiles, uint32_t* ranges)
{
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)num_tiles);
    for (int64_t i = 0; i < L; i++) {
        const uint32_t cur = (uint32_t)(sorted_keys[i] >> 32);
        if (i == 0)
            ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(sorted_keys[i - 1] >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)i;
                ranges[2 * cur] = (uint32_t)i;
            }
        }
        if (i == L - 1) ranges[2 * cur + 1] = (uint32_t)L;
    }
}

/* forward.cu:261-374 renderCUDA.  colors [P,3]; out_color [3,H,W]; final_T, n_contrib [H*W] */
void orc_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                       const float* colors, const float* conic_opacity, const float* bg, float* out_color, float* final_T,
                       uint32_t* n_contrib)
{
    const int gx = (W + ORC_BLOCK_X - 1) / ORC_BLOCK_X, gy = (H + ORC_BLOCK_Y - 1) / ORC_BLOCK_Y;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < ORC_BLOCK_Y; ly++)
            for (int lx = 0; lx < ORC_BLOCK_X; lx++) {
                const int pxi = tx * ORC_BLOCK_X + lx, pyi = ty * ORC_BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const float pxf = (float)pxi, pyf = (float)pyi;
                float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = r0; k < r1; k++) {
                    const uint32_t id = point_list[k];
                    contributor++;
                    const float dx = means2D[2 * id] - pxf, dy = means2D[2 * id + 1] - pyf;
                    const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], Cc = conic_opacity[4 * id + 2],
                                o = conic_opacity[4 * id + 3];
                    /* forward.cu:335; SASS: fma(fma(dx,A*dx,(C*dy)*dy), -0.5, -((B*dx)*dy)) */
                    const float q = fmaf(dx, dx * A, dy * (dy * Cc));
                    const float power = fmaf(q, -0.5f, -(dy * (dx * B)));
                    if (power > 0.0f) continue;
                    float alpha = o * expf(power);
                    alpha = alpha < 0.99f ? alpha : 0.99f; /* min(0.99f, .) */
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break; /* done = true */
                    C0 = fmaf(T, alpha * colors[3 * id + 0], C0);
                    C1 = fmaf(T, alpha * colors[3 * id + 1], C1);
                    C2 = fmaf(T, alpha * colors[3 * id + 2], C2);
                    T = test_T;
                    last = contributor;
                }
                const size_t pid = (size_t)W * pyi + pxi;
                final_T[pid] = T;
                n_contrib[pid] = last;
                out_color[0 * (size_t)H * W + pid] = fmaf(bg[0], T, C0);
                out_color[1 * (size_t)H * W + pid] = fmaf(bg[1], T, C1);
                out_color[2 * (size_t)H * W + pid] = fmaf(bg[2], T, C2);
            }
    }
}

/* backward.cu:399-557 renderCUDA (backward).  Accumulators are fp64 (the reference
 * uses fp32 atomics in nondeterministic order).  dL_dmean2D [P,3], dL_dconic [P,4],
 * dL_dopacity [P], dL_dcolors [P,3], all double, caller-zeroed. */
void orc_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg, const float* means2D,
                        const float* conic_opacity, const float* colors, const float* final_Ts, const uint32_t* n_contrib,
                        const float* dL_dpixels, double* dL_dmean2D, double* dL_dconic, double* dL_dopacity, double* dL_dcolors)
{
    const int gx = (W + ORC_BLOCK_X - 1) / ORC_BLOCK_X, gy = (H + ORC_BLOCK_Y - 1) / ORC_BLOCK_Y;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    /* serial over tiles: accumulators are shared between tiles */
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const uint32_t todo = r1 - r0;
        for (int ly = 0; ly < ORC_BLOCK_Y; ly++)
            for (int lx = 0; lx < ORC_BLOCK_X; lx++) {
                const int pxi = tx * ORC_BLOCK_X + lx, pyi = ty * ORC_BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const size_t pid = (size_t)W * pyi + pxi;
                const float pxf = (float)pxi, pyf = (float)pyi;
                const float T_final = final_Ts[pid];
                float T = T_final;
                uint32_t contributor = todo;
                const uint32_t last_contributor = n_contrib[pid];
                float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
                const float dpx[3] = {dL_dpixels[0 * (size_t)H * W + pid], dL_dpixels[1 * (size_t)H * W + pid],
                                      dL_dpixels[2 * (size_t)H * W + pid]};
                for (uint32_t k = 0; k < todo; k++) {
                    const uint32_t id = point_list[r1 - k - 1];
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const float dx = means2D[2 * id] - pxf, dy = means2D[2 * id + 1] - pyf;
                    const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], Cc = conic_opacity[4 * id + 2],
                                o = conic_opacity[4 * id + 3];
                    const float q = fmaf(dx, dx * A, dy * (dy * Cc));
                    const float power = fmaf(q, -0.5f, -(dy * (dx * B)));
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    float alpha = o * G;
                    alpha = alpha < 0.99f ? alpha : 0.99f;
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = colors[3 * id + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dpx[ch];
                        dL_dcolors[3 * (size_t)id + ch] += (double)(dchannel_dcolor * dpx[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0;
                    for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dpx[ch];
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = o * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * A - gdy * B;
                    const float dG_ddely = -gdy * Cc - gdx * B;
                    dL_dmean2D[3 * (size_t)id + 0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                    dL_dmean2D[3 * (size_t)id + 1] += (double)(dL_dG * dG_ddely * ddely_dy);
                    dL_dconic[4 * (size_t)id + 0] += (double)(-0.5f * gdx * dx * dL_dG);
                    dL_dconic[4 * (size_t)id + 1] += (double)(-0.5f * gdx * dy * dL_dG);
                    dL_dconic[4 * (size_t)id + 3] += (double)(-0.5f * gdy * dy * dL_dG);
                    dL_dopacity[id] += (double)(G * dL_dalpha);
                }
            }
    }
}

/* glm-convention 3x3 helpers: m[c][r] (column c, row r); type_mat3x3.inl:486-518 */
typedef struct { float m[3][3]; } gmat3;
static gmat3 gmul(const gmat3* A, const gmat3* B)
{
    gmat3 R;
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) R.m[c][r] = A->m[0][r] * B->m[c][0] + A->m[1][r] * B->m[c][1] + A->m[2][r] * B->m[c][2];
    return R;
}
static gmat3 gtranspose(const gmat3* A)
{
    gmat3 R;
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) R.m[c][r] = A->m[r][c];
    return R;
}

/* backward.cu:20-139 computeColorFromSH (backward) */
static void sh_backward(int deg, int M, const float* mean, const float* campos, const float* sh, const uint8_t* clamped,
                        const float* dL_dcolor3, float* dL_dmean_add, float* dL_dsh)
{
    const float dox = mean[0] - campos[0], doy = mean[1] - campos[1], doz = mean[2] - campos[2];
    const float len = sqrtf(dox * dox + doy * doy + doz * doz);
    const float x = dox / len, y = doy / len, z = doz / len;
    float dRGB[3];
    for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor3[c] * (clamped[c] ? 0.f : 1.f);
    float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
    (void)M;
#define SHV(k, c) sh[(k) * 3 + (c)]
#define DSH(k, w) for (int c = 0; c < 3; c++) dL_dsh[(k) * 3 + c] = (w) * dRGB[c]
    DSH(0, SH_C0);
    if (deg > 0) {
        DSH(1, -SH_C1 * y);
        DSH(2, SH_C1 * z);
        DSH(3, -SH_C1 * x);
        for (int c = 0; c < 3; c++) {
            dRGBdx[c] = -SH_C1 * SHV(3, c);
            dRGBdy[c] = -SH_C1 * SHV(1, c);
            dRGBdz[c] = SH_C1 * SHV(2, c);
        }
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, SH_C2[0] * xy);
            DSH(5, SH_C2[1] * yz);
            DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
            DSH(7, SH_C2[3] * xz);
            DSH(8, SH_C2[4] * (xx - yy));
            for (int c = 0; c < 3; c++) {
                dRGBdx[c] += SH_C2[0] * y * SHV(4, c) + SH_C2[2] * 2.f * -x * SHV(6, c) + SH_C2[3] * z * SHV(7, c) + SH_C2[4] * 2.f * x * SHV(8, c);
                dRGBdy[c] += SH_C2[0] * x * SHV(4, c) + SH_C2[1] * z * SHV(5, c) + SH_C2[2] * 2.f * -y * SHV(6, c) + SH_C2[4] * 2.f * -y * SHV(8, c);
                dRGBdz[c] += SH_C2[1] * y * SHV(5, c) + SH_C2[2] * 2.f * 2.f * z * SHV(6, c) + SH_C2[3] * x * SHV(7, c);
            }
            if (deg > 2) {
                DSH(9, SH_C3[0] * y * (3.f * xx - yy));
                DSH(10, SH_C3[1] * xy * z);
                DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy));
                DSH(14, SH_C3[5] * z * (xx - yy));
                DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
                for (int c = 0; c < 3; c++) {
                    dRGBdx[c] += (SH_C3[0] * SHV(9, c) * 3.f * 2.f * xy + SH_C3[1] * SHV(10, c) * yz + SH_C3[2] * SHV(11, c) * -2.f * xy +
                                  SH_C3[3] * SHV(12, c) * -3.f * 2.f * xz + SH_C3[4] * SHV(13, c) * (-3.f * xx + 4.f * zz - yy) +
                                  SH_C3[5] * SHV(14, c) * 2.f * xz + SH_C3[6] * SHV(15, c) * 3.f * (xx - yy));
                    dRGBdy[c] += (SH_C3[0] * SHV(9, c) * 3.f * (xx - yy) + SH_C3[1] * SHV(10, c) * xz +
                                  SH_C3[2] * SHV(11, c) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHV(12, c) * -3.f * 2.f * yz +
                                  SH_C3[4] * SHV(13, c) * -2.f * xy + SH_C3[5] * SHV(14, c) * -2.f * yz + SH_C3[6] * SHV(15, c) * -3.f * 2.f * xy);
                    dRGBdz[c] += (SH_C3[1] * SHV(10, c) * xy + SH_C3[2] * SHV(11, c) * 4.f * 2.f * yz +
                                  SH_C3[3] * SHV(12, c) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHV(13, c) * 4.f * 2.f * xz +
                                  SH_C3[5] * SHV(14, c) * (xx - yy));
                }
            }
        }
    }
#undef SHV
#undef DSH
    const float ddx = dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2];
    const float ddy = dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2];
    const float ddz = dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2];
    /* auxiliary.h:107-117 dnormvdv */
    const float sum2 = dox * dox + doy * doy + doz * doz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dL_dmean_add[0] = ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
    dL_dmean_add[1] = (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
    dL_dmean_add[2] = (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
}

/* backward.cu:144-274 computeCov2DCUDA + :346-396 preprocessCUDA(bwd) + :278-341 computeCov3D(bwd).
 * Inputs: blend-stage gradients as fp32 [P,3]/[P,4]/[P,3]; cov3D = forward cov3D (or precomp).
 * Outputs [P,*] fp32, zero for radii<=0 (rasterize_points.cu:150-158 zero-inits). */
void orc_preprocess_backward(const orc_scene* s, const int32_t* radii, const float* cov3Ds, const uint8_t* clamped,
                             const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor, float* dL_dmean3D,
                             float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot)
{
    const int P = s->P, M = s->M;
    const float h_y = s->H / (2.0f * s->tan_fovy), h_x = s->W / (2.0f * s->tan_fovx);
    const float* vm = s->viewmatrix;
    const float* proj = s->projmatrix;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dL_dmean3D[3 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
        if (dL_dscale) for (int k = 0; k < 3; k++) dL_dscale[3 * i + k] = 0.f;
        if (dL_drot) for (int k = 0; k < 4; k++) dL_drot[4 * i + k] = 0.f;
        if (dL_dsh) for (int k = 0; k < 3 * M; k++) dL_dsh[(size_t)3 * M * i + k] = 0.f;
        if (!(radii[i] > 0)) continue;
        const float* c6 = cov3Ds + 6 * i;
        const float mx = s->means3D[3 * i], my = s->means3D[3 * i + 1], mz = s->means3D[3 * i + 2];
        const float dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
        /* --- computeCov2DCUDA --- */
        float tx = vm[0] * mx + vm[4] * my + vm[8] * mz + vm[12];
        float ty = vm[1] * mx + vm[5] * my + vm[9] * mz + vm[13];
        const float tz = vm[2] * mx + vm[6] * my + vm[10] * mz + vm[14];
        const float limx = 1.3f * s->tan_fovx, limy = 1.3f * s->tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        gmat3 J = {{{h_x / tz, 0.0f, -(h_x * tx) / (tz * tz)}, {0.0f, h_y / tz, -(h_y * ty) / (tz * tz)}, {0, 0, 0}}};
        gmat3 Wm = {{{vm[0], vm[4], vm[8]}, {vm[1], vm[5], vm[9]}, {vm[2], vm[6], vm[10]}}};
        gmat3 Vrk = {{{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}}};
        gmat3 T = gmul(&Wm, &J);
        gmat3 Tt = gtranspose(&T), Vt = gtranspose(&Vrk);
        gmat3 tmp = gmul(&Tt, &Vt);
        gmat3 cov2D = gmul(&tmp, &T);
        const float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcv = dL_dcov3D + 6 * i;
#define TT(c_, r_) T.m[c_][r_]
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcv[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
            dcv[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
            dcv[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
            dcv[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 1) * dL_dc;
            dcv[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 2) * dL_dc;
            dcv[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db + 2 * TT(1, 1) * TT(1, 2) * dL_dc;
        }
#define VV(c_, r_) Vrk.m[c_][r_]
        const float dL_dT00 = 2 * (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_da + (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_db;
        const float dL_dT01 = 2 * (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_da + (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_db;
        const float dL_dT02 = 2 * (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_da + (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_db;
        const float dL_dT10 = 2 * (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_dc + (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_db;
        const float dL_dT11 = 2 * (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_dc + (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_db;
        const float dL_dT12 = 2 * (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_dc + (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_db;
#undef VV
#undef TT
        const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
        const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
        const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
        const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;
        const float itz = 1.f / tz, tz2 = itz * itz, tz3 = tz2 * itz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * tx) * tz3 * dL_dJ02 + (2 * h_y * ty) * tz3 * dL_dJ12;
        /* auxiliary.h:89-97 transformVec4x3Transpose; assignment (backward.cu:273) */
        float dm[3] = {vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz, vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
                       vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz};
        /* --- preprocessCUDA (bwd) backward.cu:370-387 --- */
        const float m_hw = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
        const float m_w = 1.0f / (m_hw + 0.0000001f);
        const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
        dm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        if (s->shs) {
            float add[3];
            sh_backward(s->D, M, s->means3D + 3 * i, s->campos, s->shs + (size_t)3 * M * i, clamped + 3 * i, dL_dcolor + 3 * i, add,
                        dL_dsh + (size_t)3 * M * i);
            dm[0] += add[0]; dm[1] += add[1]; dm[2] += add[2];
        }
        dL_dmean3D[3 * i] = dm[0]; dL_dmean3D[3 * i + 1] = dm[1]; dL_dmean3D[3 * i + 2] = dm[2];
        if (s->scales) {
            /* backward.cu:278-341 */
            const float* q = s->rotations + 4 * i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            gmat3 R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                        {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                        {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
            const float sv[3] = {s->scale_modifier * s->scales[3 * i], s->scale_modifier * s->scales[3 * i + 1], s->scale_modifier * s->scales[3 * i + 2]};
            gmat3 S = {{{sv[0], 0, 0}, {0, sv[1], 0}, {0, 0, sv[2]}}};
            gmat3 Mm = gmul(&S, &R);
            gmat3 dSig = {{{dcv[0], 0.5f * dcv[1], 0.5f * dcv[2]}, {0.5f * dcv[1], dcv[3], 0.5f * dcv[4]}, {0.5f * dcv[2], 0.5f * dcv[4], dcv[5]}}};
            gmat3 M2;
            for (int cc = 0; cc < 3; cc++) for (int rr = 0; rr < 3; rr++) M2.m[cc][rr] = 2.0f * Mm.m[cc][rr];
            gmat3 dL_dM = gmul(&M2, &dSig);
            gmat3 Rt = gtranspose(&R), dMt = gtranspose(&dL_dM);
            for (int k = 0; k < 3; k++) dL_dscale[3 * i + k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
            for (int k = 0; k < 3; k++) for (int rr = 0; rr < 3; rr++) dMt.m[k][rr] *= sv[k];
            float* dq = dL_drot + 4 * i;
#define D(c_, r_) dMt.m[c_][r_]
            dq[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
            dq[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
            dq[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * y * (D(2, 2) + D(0, 0));
            dq[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
#undef D
        }
    }
}
