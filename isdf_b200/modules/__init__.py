from . import embedding, fc_map, sample, loss, render, trainer  # noqa: F401  (all registered under isdf.modules.* by the alias package)
