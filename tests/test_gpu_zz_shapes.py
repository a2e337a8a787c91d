"""GPU parity at the model shapes of the other shipped / BASELINE configs: wide MLP 512x(4+4)
(BASELINE configs[4]), realsense E = 381 (n_embed_funcs 8), franka E = 465 with hidden_layers_block 3.  All three on the
CUDA-core fp32 path; the two wide embeddings also on the tcgen05 path (two embedding halves); hidden = 512 must be refused
loudly by the tcgen05 path.  Checked against the fp64 oracle like the default shape.  (File name sorts last on purpose: the default-shape suites
run first.)"""
import pytest
import torch

from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SHAPES = [("wide_512x8", 6, 512, 4), ("realsense_E381", 9, 256, 2), ("franka_E465_block3", 11, 256, 3)]


def _case(n_freqs, hidden, block, R=48, S=16):
    cfg = O.default_cfg(n_freqs=n_freqs, hidden=hidden, block=block, noise_std=0.05, n_strat=S - 8, n_surf=8)
    sd = C.golden_weights(5, E=3 + 42 * n_freqs, H=hidden, block=block, gain=1.2)
    batch, noise = C.loss_batch(9, R, S=S)
    return cfg, sd, batch, noise


@pytest.mark.parametrize("tag,n_freqs,hidden,block", SHAPES, ids=[s[0] for s in SHAPES])
def test_fp32_path_matches_oracle_at_other_shapes(tag, n_freqs, hidden, block):
    cfg, sd, batch, noise = _case(n_freqs, hidden, block)
    eng = P.make_engine(DEV, cfg, "fp32", max_points=512)          # 768 samples: two internal chunks
    assert eng.embedding_size == 3 + 42 * n_freqs and eng.n_params == sum(v.numel() for v in sd.values())
    out = P.run_train(eng, sd, batch, noise, cfg, DEV)
    ref = P.oracle_train(sd, batch, noise, cfg)
    e = P.compare_train(out, ref)
    # Bound = 3x the error the reference's OWN fp32 arithmetic makes against fp64 at this shape (the same restatement
    # evaluated in torch fp32), never tighter than the default-shape bound.  With 11 octaves, sin(2^10 x) in fp32 alone
    # puts the reference 5.9e-6 away from fp64 (measured: 1.5e-6 / 1.5e-6 / 5.9e-6 for the three shapes), so a fixed
    # 5e-6 would ask the kernel to beat the arithmetic it is compared with.
    ref32 = P.oracle_train(sd, batch, noise, cfg, dtype=torch.float32)
    floor_sdf, floor_g = P.rel(ref32["sdf"], ref["sdf"]), P.rel(ref32["g"], ref["g"])
    tol_sdf, tol_g = max(5e-6, 3 * floor_sdf), max(1e-4, 3 * floor_g)
    assert e["sdf"] < tol_sdf and e["g"] < tol_g, (e, floor_sdf, floor_g)
    assert e["total_loss"] < 5e-5 and e["sdf_loss"] < 5e-5, e
    assert e["grad_max_rel_fro"] < 1e-3, e
    # forward-only and forward + input gradient on ragged sizes
    x = batch["pc"].reshape(-1, 3)[:301].to(DEV).contiguous()
    layers = [(w.double(), b.double()) for w, b in O.layers_from_state_dict(sd, block)]
    sdf_ref = O.sdf_forward(layers, x.cpu().double(), cfg)
    sdf, g = eng.forward(x, want_grad=True)
    assert P.rel(sdf.cpu(), sdf_ref) < tol_sdf and P.rel(eng.forward(x).cpu(), sdf_ref) < tol_sdf
    assert P.rel(g.cpu(), out["g"].reshape(-1, 3)[:301]) < 1e-4


@pytest.mark.parametrize("mode", ["bf16x3", "bf16x3g"])
@pytest.mark.parametrize("tag,n_freqs,hidden,block", SHAPES[1:], ids=[s[0] for s in SHAPES[1:]])
def test_tensor_core_path_wide_embeddings_match_oracle(tag, n_freqs, hidden, block, mode):
    """E = 381 / 465 (n_embed_funcs 8 / 10 of the realsense / franka configs, block 3) on the tcgen05 path: the padded
    embedding is two halves of 256 internal columns, every embedding-fed product the sum of two 128x256x256 products."""
    cfg, sd, batch, noise = _case(n_freqs, hidden, block, R=60, S=16)          # 960 samples: 8 tiles, 2 chunks
    eng = P.make_engine(DEV, cfg, mode, max_points=512)
    assert eng.embedding_size == 3 + 42 * n_freqs
    out = P.run_train(eng, sd, batch, noise, cfg, DEV)
    ref = P.oracle_train(sd, batch, noise, cfg)
    e = P.compare_train(out, ref)
    assert e["sdf"] < 1e-4 and e["g"] < 1e-3, e
    assert e["total_loss"] < 1e-3 and e["sdf_loss"] < 1e-3, e
    assert e["grad_max_rel_fro"] < 2e-2, e                                   # small batch: see TOL gw_small
    x = batch["pc"].reshape(-1, 3)[:301].to(DEV).contiguous()
    layers = [(w.double(), b.double()) for w, b in O.layers_from_state_dict(sd, block)]
    sdf_ref = O.sdf_forward(layers, x.cpu().double(), cfg)
    sdf, g = eng.forward(x, want_grad=True)
    assert P.rel(sdf.cpu(), sdf_ref) < 1e-4 and P.rel(eng.forward(x).cpu(), sdf_ref) < 1e-4
    assert P.rel(g.cpu(), out["g"].reshape(-1, 3)[:301]) < 1e-3
    # chunk invariance at this shape
    out2 = P.run_train(P.make_engine(DEV, cfg, mode, max_points=4096), sd, batch, noise, cfg, DEV)
    assert P.rel(out2["sdf"], out["sdf"]) < 5e-5
    assert max(P.rel_fro(a, b) for a, b in zip(out2["grads"], out["grads"])) < 2e-3


def test_tensor_core_path_refuses_other_widths_loudly():
    from isdf_b200 import _lib
    tag, n_freqs, hidden, block = SHAPES[0]
    cfg, _, _, _ = _case(n_freqs, hidden, block)
    with pytest.raises(_lib.IsdfbError, match="tensor-core path supports hidden=256"):
        P.make_engine(DEV, cfg, "bf16x3", max_points=512)
