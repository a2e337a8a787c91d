"""TEST INFRASTRUCTURE ONLY -- import shim for the real reference (/root/reference).

The reference iSDF package imports GUI / mesh libraries (trimesh, open3d, ...)
at module scope; none of them is touched by the training step.  This shim
installs a meta-path finder that hands out MagicMock packages for the missing
ones and puts /root/reference on sys.path so that the UNMODIFIED reference
modules can be imported on CPU in the build container.

It is used by
  * tests/golden/make_golden.py (generates the committed fixtures), and
  * tests that are skipped when /root/reference is absent (GPU box).
Nothing in isdf_b200/ may import this file.
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """The read-only checkout in the build container; on the GPU box the verbatim copy under oracle/_ref
    (oracle/make_ref.py).  ISDF_REFERENCE_ROOT overrides both."""
    env = os.environ.get("ISDF_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if os.path.isdir(os.path.join(cand, "isdf", "modules")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()

_ABSENT = {
    "trimesh", "imgviz", "matplotlib", "open3d", "skimage", "pyglet",
    "imageio", "urdfpy", "rospy", "sensor_msgs", "geometry_msgs",
    "orb_slam3_ros_wrapper", "cv_bridge",
}


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = MagicMock(name=spec.name)
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__name__ = spec.name
        return mod

    def exec_module(self, module):
        return None


class _MockFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        root = name.split(".")[0]
        if root not in _ABSENT:
            return None
        try:  # only mock what is really missing
            for finder in sys.meta_path:
                if finder is self:
                    continue
                spec = finder.find_spec(name, path, target) if hasattr(finder, "find_spec") else None
                if spec is not None:
                    return None
        except Exception:
            pass
        return importlib.machinery.ModuleSpec(name, _MockLoader(), is_package=True)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "isdf", "modules"))


def load():
    """Return the reference's `isdf.modules` sub-modules as a dict.

    The repo ships its own top-level `isdf` compatibility package; to make sure
    the REAL reference is imported, any already-imported `isdf*` modules are
    evicted and /root/reference is placed first on sys.path for the import.
    The evicted modules are restored afterwards so both can coexist.
    """
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not any(isinstance(f, _MockFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _MockFinder())

    saved = {k: v for k, v in sys.modules.items() if k == "isdf" or k.startswith("isdf.")}
    for k in saved:
        del sys.modules[k]
    # the repo's own regular `isdf` package would shadow the reference's namespace package
    # regardless of path order -> hide every path entry that holds it while importing
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hidden = [(i, e) for i, e in enumerate(sys.path)
              if os.path.isfile(os.path.join(os.path.abspath(e or "."), "isdf", "__init__.py"))]
    for _, e in hidden:
        sys.path.remove(e)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from isdf.modules import trainer, fc_map, embedding, sample, loss, render  # noqa
            from isdf.geometry import transform  # noqa
            from isdf.datasets import data_util  # noqa
        ref = dict(trainer=trainer, fc_map=fc_map, embedding=embedding, sample=sample,
                   loss=loss, render=render, transform=transform, data_util=data_util)
        assert trainer.__file__.startswith(REFERENCE_ROOT), trainer.__file__
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for i, e in hidden:
            sys.path.insert(min(i, len(sys.path)), e)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "isdf" or k.startswith("isdf.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref
