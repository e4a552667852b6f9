// knn.cu -- mean squared distance to the three nearest neighbours (the co-requisite `simple_knn._C.distCUDA2`,
// SURVEY 8b).  Replaces SimpleKNN::knn (gaussian_splatting/submodules/simple-knn/simple_knn.cu:188-220): same result --
// for every point the mean of the squared distances to its 3 nearest OTHER points (simple_knn.cu:139-186: the point's
// own index is skipped, coincident points count with distance 0, missing neighbours stay FLT_MAX) -- computed on a
// uniform grid instead of Morton-sorted boxes: the host side (simple_knn/_C.py) bins the points into cells of about
// two points each, this kernel walks the cell shells around a point until the third-best distance is closer than the
// unsearched space.  The distance expression is the reference's (simple_knn.cu:125-127, FMA-contracted by nvcc).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/gstar_raster.h"

namespace {

__device__ __forceinline__ void update3(float dist, float best[3])
{
#pragma unroll
    for (int j = 0; j < 3; j++) {  // simple_knn.cu:128-136
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
        }
    }
}

__global__ void __launch_bounds__(256) k_knn3(int P, const float* __restrict__ pts, const int* __restrict__ order, const int* __restrict__ cell_start,
                                              int nx, int ny, int nz, float ox, float oy, float oz, float cell, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;  // position in cell order: neighbouring threads search the same cells
    if (k >= P) return;
    const int me = order[k];
    const float px = pts[3 * (size_t)me], py = pts[3 * (size_t)me + 1], pz = pts[3 * (size_t)me + 2];
    const float inv = 1.0f / cell;
    const int cx = min(nx - 1, max(0, (int)floorf((px - ox) * inv)));
    const int cy = min(ny - 1, max(0, (int)floorf((py - oy) * inv)));
    const int cz = min(nz - 1, max(0, (int)floorf((pz - oz) * inv)));
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    const int rmax = max(max(max(cx, nx - 1 - cx), max(cy, ny - 1 - cy)), max(cz, nz - 1 - cz));
    for (int r = 0; r <= rmax; r++) {
        // every point outside the cube of cells within Chebyshev radius r-1 is farther than (r-1)*cell (with a margin
        // for the rounding of the cell coordinates)
        if (r >= 2) {
            const float safe = (float)(r - 1) * cell * 0.999f;
            if (best[2] <= safe * safe) break;
        }
        const int z0 = max(0, cz - r), z1 = min(nz - 1, cz + r);
        const int y0 = max(0, cy - r), y1 = min(ny - 1, cy + r);
        const int x0 = max(0, cx - r), x1 = min(nx - 1, cx + r);
        for (int z = z0; z <= z1; z++) {
            const bool zshell = (z == cz - r) || (z == cz + r);
            for (int y = y0; y <= y1; y++) {
                const bool yshell = zshell || (y == cy - r) || (y == cy + r);
                // on a shell face every x of the row belongs to the shell; otherwise only the two end cells
                const int xstep = (yshell || r == 0) ? 1 : max(1, 2 * r);
                for (int x = (yshell || r == 0) ? x0 : cx - r; x <= x1; x += xstep) {
                    if (x < x0) continue;
                    const int c = (z * ny + y) * nx + x;
                    const int b = cell_start[c], e = cell_start[c + 1];
                    for (int i = b; i < e; i++) {
                        const int o = order[i];
                        if (o == me) continue;  // simple_knn.cu:171-172
                        const float dx = pts[3 * (size_t)o] - px, dy = pts[3 * (size_t)o + 1] - py, dz = pts[3 * (size_t)o + 2] - pz;
                        update3(dx * dx + dy * dy + dz * dz, best);
                    }
                }
            }
        }
    }
    out[me] = (best[0] + best[1] + best[2]) / 3.0f;  // simple_knn.cu:185
}

}  // namespace

extern "C" int gstar_knn3_mean_dist2(int P, const float* points, const int* order, const int* cell_start, int nx, int ny, int nz, float ox, float oy,
                                     float oz, float cell, float* out, void* stream)
{
    if (P <= 0) return 0;
    if (!points || !order || !cell_start || !out || nx <= 0 || ny <= 0 || nz <= 0 || !(cell > 0.0f)) return GSTAR_ERR_INVALID;
    k_knn3<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, points, order, cell_start, nx, ny, nz, ox, oy, oz, cell, out);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : GSTAR_ERR_CUDA;
}
