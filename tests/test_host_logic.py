"""CPU tests of the host-side mirror (no kernels): render / grid helpers and the checkpoint format against
fixtures produced by the unmodified reference (tests/golden/make_golden_infer.py)."""
import io
import os

import torch

from tests.golden import common as C
describe = C.describe

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_sdf_render_depth_matches_reference():
    from isdf_b200.modules import render
    g = load("infer.pt")["render_depth"]
    assert torch.equal(render.sdf_render_depth(g["z"], g["sdf"]), g["depth"])


def test_make_3d_grid_matches_reference():
    from isdf_b200.geometry import transform
    g = load("infer.pt")["grid"]
    T_box = C.rigid_transform(14)
    scale = g["extents"] / (2.0 * 0.9)
    pc = transform.make_3D_grid([-1.0, 1.0], 12, "cpu", transform=torch.inverse(T_box), scale=scale).view(-1, 3)
    assert torch.allclose(pc, g["grid_pc"], atol=1e-6)


def test_checkpoint_structure_matches_reference_files():
    """Row N4: what Trainer.save_checkpoint writes has exactly the nested keys / shapes / dtypes of the file the
    reference driver writes (train.py:207-219), so either side loads the other's checkpoints."""
    from isdf_b200.modules import embedding, fc_map
    from isdf_b200.modules.trainer import FusedAdamW
    gold = load("infer.pt")["checkpoint"]
    pe = embedding.PostionalEncoding(min_deg=0, max_deg=5, scale=0.05937489)
    m = fc_map.SDFMap(pe, 256, 2, 0.14)
    m.load_state_dict(C.golden_weights(73))
    opt = FusedAdamW(m, lr=0.0013, weight_decay=0.012)
    opt.step_count = 2
    ck = {"step": 7.5, "model_state_dict": m.state_dict(), "optimizer_state_dict": opt.state_dict(), "loss": 0.25}
    buf = io.BytesIO()
    torch.save(ck, buf)
    mine = describe(torch.load(io.BytesIO(buf.getvalue()), weights_only=False))
    assert mine == gold["structure"]
    # and torch's own AdamW accepts it (what the reference's optimiser.load_state_dict would do)
    ref_opt = torch.optim.AdamW(m.parameters(), lr=0.5, weight_decay=0.5)
    ref_opt.load_state_dict(ck["optimizer_state_dict"])
    assert ref_opt.param_groups[0]["lr"] == 0.0013 and float(ref_opt.state[list(m.parameters())[0]]["step"]) == 2.0
