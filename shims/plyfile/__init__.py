"""Import-only stand-in for `plyfile` (imported at module load by the reference's `scene/gaussian_model.py:18`; used only by its
`save_ply` / `load_ply`).  `gaustar_b200/synth_dataset.py` writes PLY files itself."""


class _ImportOnly:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("shims/plyfile: import-only stand-in (plyfile is not installed in this image)")

    @classmethod
    def read(cls, *args, **kwargs):
        raise NotImplementedError("shims/plyfile: import-only stand-in (plyfile is not installed in this image)")

    @classmethod
    def describe(cls, *args, **kwargs):
        raise NotImplementedError("shims/plyfile: import-only stand-in (plyfile is not installed in this image)")


class PlyData(_ImportOnly):
    pass


class PlyElement(_ImportOnly):
    pass
