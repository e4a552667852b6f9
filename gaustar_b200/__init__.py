"""gaustar_b200 -- B200-native differentiable surface-Gaussian rasterizer.

One hot path, nothing else: the operator behind
``diff_gaussian_rasterization.GaussianRasterizer`` of eth-ait/GauSTAR
(preprocess -> tile binning / sort -> per-tile alpha compositing, forward and
backward) as hand-written sm_100a CUDA kernels behind a C ABI
(``include/gstar_raster.h``, ``gaustar_b200/lib/libgstar_raster.so``) and a thin
torch shim (``gaustar_b200/_C``).  There is no CPU fallback: importing the
rasterizer without the built extension raises.
"""
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
    _C,
    set_grad_accumulation_fusion,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "set_grad_accumulation_fusion"]
