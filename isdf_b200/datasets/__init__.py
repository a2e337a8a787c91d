from . import data_util, dataset  # noqa: F401
