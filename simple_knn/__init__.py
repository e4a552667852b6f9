"""Import shim for the reference's ``simple_knn`` extension (gaussian_splatting/submodules/simple-knn): GauSTAR imports
``from simple_knn._C import distCUDA2`` at module import time (gaustar_scene/sugar_model.py:9,
gaussian_splatting/scene/gaussian_model.py:20).  The function is implemented in ``libgstar_raster.so`` (csrc/knn.cu)."""
