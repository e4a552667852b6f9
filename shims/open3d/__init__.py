"""Import-only stand-in for `open3d` (`sugar_model.py:3`: mesh IO and TSDF fusion live in GauSTAR's mesh-surgery code, outside the
rasterizer hot path).  A surface mesh handed to `SuGaR(surface_mesh_to_bind=...)` only needs `.vertices`, `.triangles` and
`.vertex_colors` (`sugar_model.py:175,228-237`), which `TriangleMeshLike` provides for tests and synthetic data."""
import numpy as np


class TriangleMeshLike:
    def __init__(self, vertices, triangles, vertex_colors=None):
        self.vertices = np.asarray(vertices, np.float64)
        self.triangles = np.asarray(triangles, np.int32)
        self.vertex_colors = [] if vertex_colors is None else [tuple(c) for c in np.asarray(vertex_colors, np.float64)]


def __getattr__(name):
    raise NotImplementedError(f"shims/open3d: open3d.{name} is not available (open3d is not installed in this image)")
