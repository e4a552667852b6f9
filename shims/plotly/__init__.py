"""Import-only stand-in for `plotly` (`gaustar_scene/gs_model.py:5` imports `plotly.graph_objs` for an optional viewer)."""
