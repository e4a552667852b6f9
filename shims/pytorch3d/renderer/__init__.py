"""Stand-in for the `pytorch3d.renderer` names GauSTAR imports (`sugar_model.py:4,8`, `cameras.py:9-10`).

`FoVPerspectiveCameras` is functional for exactly what GauSTAR does with it -- a container of R / T / K / znear / zfar that can
be indexed, moved between devices, and asked for its camera centres (`cameras.py:229-330,537-548`, `sugar_model.py:1113-1162`);
the mesh-rasterisation classes are import-only (GauSTAR constructs them only in texture-baking and densification paths that
`refine.py` does not take with `use_densifier=False`).  Restated from the published pytorch3d 0.7.4 behaviour; parity unpinned.
"""
import torch

from . import cameras  # noqa: F401
from .cameras import FoVPerspectiveCameras  # noqa: F401


class _ImportOnly:
    _what = "pytorch3d.renderer"

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"shims/pytorch3d: {type(self).__name__} is an import-only stand-in ({self._what} is not installed in this image)")


class TexturesVertex:
    """Per-vertex features of a batch of meshes (pytorch3d/renderer/mesh/textures.py): a container here -- `SuGaR.surface_mesh`
    (`sugar_model.py:568-576`) attaches one to every `Meshes` it builds; sampling it needs the mesh rasteriser (not provided)."""

    def __init__(self, verts_features):
        if torch.is_tensor(verts_features):
            if verts_features.dim() != 3:
                raise ValueError("Expected verts_features to be of shape (N, V, D)")
            self._verts_features_padded = verts_features
            self._verts_features_list = None
        else:
            self._verts_features_list = list(verts_features)
            self._verts_features_padded = None

    def verts_features_padded(self):
        if self._verts_features_padded is None:
            self._verts_features_padded = torch.nn.utils.rnn.pad_sequence(self._verts_features_list, batch_first=True)
        return self._verts_features_padded

    def verts_features_list(self):
        if self._verts_features_list is None:
            self._verts_features_list = list(self._verts_features_padded)
        return self._verts_features_list

    def verts_features_packed(self):
        return torch.cat(self.verts_features_list(), dim=0)


class TexturesUV:
    """UV-mapped textures: a record of its constructor arguments (sampling needs the mesh rasteriser, not provided)."""

    def __init__(self, maps, faces_uvs, verts_uvs, padding_mode="border", align_corners=True, sampling_mode="bilinear"):
        self.__dict__.update(locals())
        del self.__dict__["self"]


class MeshRasterizer(_ImportOnly):
    pass


class RasterizationSettings:
    """Plain record of the settings (pytorch3d/renderer/mesh/rasterizer.py); constructing it is harmless."""

    def __init__(self, image_size=256, blur_radius=0.0, faces_per_pixel=1, bin_size=None, max_faces_per_bin=None, perspective_correct=None,
                 clip_barycentric_coords=None, cull_backfaces=False, z_clip_value=None, cull_to_frustum=False):
        self.__dict__.update(locals())
        del self.__dict__["self"]
