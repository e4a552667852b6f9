"""Partial stand-in for pytorch3d 0.7.4 (only `pytorch3d.transforms`; see shims/README.md).  Parity unpinned."""
__version__ = "0.7.4+gaustar_b200.shim"
