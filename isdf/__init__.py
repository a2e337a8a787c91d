"""Drop-in alias: `from isdf.modules import trainer` (reference train.py:16) resolves to isdf_b200.

Put this repo's root on sys.path INSTEAD of the reference's and the reference drivers
(isdf/train/train.py, train_vis.py, batch_train) import the B200 implementation unchanged."""
import sys

import isdf_b200
from isdf_b200 import modules, geometry, datasets, eval  # noqa: F401,A004

for _name, _mod in (("modules", modules), ("geometry", geometry), ("datasets", datasets), ("eval", eval)):
    sys.modules["isdf." + _name] = _mod
    for _sub in dir(_mod):
        _obj = getattr(_mod, _sub)
        if getattr(_obj, "__name__", "").startswith("isdf_b200." + _name + "."):
            sys.modules["isdf.%s.%s" % (_name, _sub)] = _obj
__version__ = isdf_b200.__version__
