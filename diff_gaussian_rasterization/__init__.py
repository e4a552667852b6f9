"""Drop-in import name.

GauSTAR's callers do ``from diff_gaussian_rasterization import
GaussianRasterizationSettings, GaussianRasterizer``
(gaustar_scene/sugar_model.py:10, gaussian_splatting/gaussian_renderer/__init__.py:14).
Putting this repository on PYTHONPATH ahead of the reference's submodule makes
those imports resolve to the B200-native implementation in ``gaustar_b200``.
"""
from gaustar_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    rasterize_gaussians,
    set_geometry_cache,
    shared_geometry,
    release_shared_geometry,
    set_deterministic_backward,
    set_pass_fusion,
    _C,
)
