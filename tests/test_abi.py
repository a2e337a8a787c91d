"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
symbols include/isdf_b200.h declares (no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def header_symbols():
    src = open(os.path.join(ROOT, "include", "isdf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isdfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for need in ("isdfb_create", "isdfb_destroy", "isdfb_pack_weights", "isdfb_gather_rays", "isdfb_sample_rays",
                 "isdfb_mlp_forward", "isdfb_mlp_forward_grad", "isdfb_train_fwd_bwd", "isdfb_frame_bins",
                 "isdfb_adamw"):
        assert need in syms


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (isdfb_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported, set(header_symbols()) - exported


def test_ctypes_binding_covers_header(built):
    from isdf_b200 import _lib
    lib = _lib.load()
    assert set(_lib.SIGNATURES) == set(header_symbols())
    for name in header_symbols():
        assert hasattr(lib, name)


def test_sass_is_sm100a(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "isdf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt or "import oracle" not in txt and "from oracle" not in txt, f


def test_no_cpu_fallback():
    import torch
    from isdf_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(torch.device("cpu"), 6, 256, 2, 0.05, 0.14)
