from . import embedding, fc_map, sample, loss, render  # noqa: F401
