"""Import-only stand-in (see shims/plotly/__init__.py)."""


def __getattr__(name):
    raise NotImplementedError(f"shims/plotly: graph_objs.{name} is not available (plotly is not installed in this image)")
