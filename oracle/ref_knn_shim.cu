// TEST INFRASTRUCTURE ONLY.  C entry point over the UNMODIFIED reference simple_knn (compiled where it lies by oracle/build_ref.py into
// oracle/_ref/libref_knn.so): SimpleKNN::knn, gaussian_splatting/submodules/simple-knn/simple_knn.cu:188-220, the function the
// reference's own binding (spatial.cu:15-26, distCUDA2) calls.  Device pointers in, device pointer out; returns the CUDA error code.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_knn_mean_dist2(int P, const float* points_dev, float* mean_dists_dev)
{
    SimpleKNN::knn(P, reinterpret_cast<float3*>(const_cast<float*>(points_dev)), mean_dists_dev);
    return (int)cudaDeviceSynchronize();
}
