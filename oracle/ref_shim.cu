// ref_shim.cu -- C-ABI shim around the UNMODIFIED reference rasterizer (test infrastructure only).
//
// Compiled by oracle/build_ref.py together with the reference's own
//   DGR/cuda_rasterizer/{rasterizer_impl,forward,backward}.cu
// straight from /root/reference (nothing is copied into this repo) into oracle/_ref/libref_dgr.so.
// It calls CudaRasterizer::Rasterizer::{forward,backward,markVisible} (rasterizer.h:24-84) and uses
// the reference's own GeometryState/BinningState/ImageState::fromChunk (rasterizer_impl.h:33-68) to
// expose every intermediate buffer, so that the parity tests can compare them bit for bit.
// Never linked into, imported by, or executed from the product path.
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

#include "rasterizer.h"
#include "rasterizer_impl.h"

namespace {
struct Buf {
    char* ptr = nullptr;
    size_t cap = 0;
    char* get(size_t n)
    {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            cudaMalloc((void**)&ptr, n + 256);
            cap = n;
        }
        return ptr;
    }
};
Buf g_geom, g_bin, g_img;
int g_P = 0, g_R = 0, g_W = 0, g_H = 0;
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_forward(int P, int D, int M, const float* background, int W, int H, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color, int* radii, int debug)
{
    std::function<char*(size_t)> gf = [](size_t n) { return g_geom.get(n); };
    std::function<char*(size_t)> bf = [](size_t n) { return g_bin.get(n); };
    std::function<char*(size_t)> imf = [](size_t n) { return g_img.get(n); };
    cudaMemset(out_color, 0, sizeof(float) * 3 * (size_t)W * H);  // rasterize_points.cu:66
    cudaMemset(radii, 0, sizeof(int) * (size_t)P);                // rasterize_points.cu:67
    int R = 0;
    try {
        R = CudaRasterizer::Rasterizer::forward(gf, bf, imf, P, D, M, background, W, H, means3D, shs, colors_precomp, opacities, scales,
                                                scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
                                                tan_fovy, prefiltered != 0, out_color, radii, debug != 0);
    } catch (...) {
        return -1;
    }
    g_P = P; g_R = R; g_W = W; g_H = H;
    return cudaDeviceSynchronize() == cudaSuccess ? R : -2;
}

// all outputs must be zero-filled by the caller (rasterize_points.cu:150-158)
REF_API int ref_backward(int P, int D, int M, int R, const float* background, int W, int H, const float* means3D, const float* shs,
                         const float* colors_precomp, const float* scales, float scale_modifier, const float* rotations,
                         const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos,
                         float tan_fovx, float tan_fovy, const int* radii, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                         float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                         float* dL_drot, int debug)
{
    try {
        CudaRasterizer::Rasterizer::backward(P, D, M, R, background, W, H, means3D, shs, colors_precomp, scales, scale_modifier, rotations,
                                             cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, g_geom.ptr, g_bin.ptr,
                                             g_img.ptr, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
                                             dL_dscale, dL_drot, debug != 0);
    } catch (...) {
        return -1;
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

REF_API int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

#define COPY(dst, src, n)                                                          \
    if (dst) cudaMemcpy(dst, src, (n), cudaMemcpyDeviceToDevice)

// Copy the intermediates of the last ref_forward into caller device arrays (any may be NULL).
REF_API int ref_get_geometry(float* depths, float* means2D, float* cov3D, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                             uint32_t* point_offsets, unsigned char* clamped)
{
    char* chunk = g_geom.ptr;
    auto gs = CudaRasterizer::GeometryState::fromChunk(chunk, g_P);
    const size_t P = g_P;
    COPY(depths, gs.depths, P * 4);
    COPY(means2D, gs.means2D, P * 8);
    COPY(cov3D, gs.cov3D, P * 24);
    COPY(conic_opacity, gs.conic_opacity, P * 16);
    COPY(rgb, gs.rgb, P * 12);
    COPY(tiles_touched, gs.tiles_touched, P * 4);
    COPY(point_offsets, gs.point_offsets, P * 4);
    COPY(clamped, gs.clamped, P * 3);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

REF_API int ref_get_binning(uint64_t* keys_unsorted, uint32_t* values_unsorted, uint64_t* keys_sorted, uint32_t* point_list)
{
    char* chunk = g_bin.ptr;
    auto bs = CudaRasterizer::BinningState::fromChunk(chunk, g_R);
    const size_t R = g_R;
    COPY(keys_unsorted, bs.point_list_keys_unsorted, R * 8);
    COPY(values_unsorted, bs.point_list_unsorted, R * 4);
    COPY(keys_sorted, bs.point_list_keys, R * 8);
    COPY(point_list, bs.point_list, R * 4);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

REF_API int ref_get_image(float* accum_alpha, uint32_t* n_contrib, uint32_t* ranges /* [T][2] */)
{
    char* chunk = g_img.ptr;
    const size_t N = (size_t)g_W * g_H;
    auto is = CudaRasterizer::ImageState::fromChunk(chunk, N);
    const size_t T = (size_t)((g_W + 15) / 16) * ((g_H + 15) / 16);
    COPY(accum_alpha, is.accum_alpha, N * 4);
    COPY(n_contrib, is.n_contrib, N * 4);
    COPY(ranges, is.ranges, T * 8);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}
