"""Golden fixtures for the forward-only rows N1 / N4 of SURVEY.md 8f, from the UNMODIFIED reference on CPU.
   python tests/golden/make_golden_infer.py   ->  tests/golden/infer.pt

 * grid: transform.make_3D_grid + fc_map.chunks over a 12^3 lattice with an oriented box (trainer.py:103-156,
   1426-1444);
 * render: render.sdf_render_depth on crafted sdf profiles, render.render_normals (render.py:12-57);
 * sample: sample.sample_along_rays with gt_depth=None (the render passes, trainer.py:1087-1128);
 * checkpoint: structure of the dict the reference driver saves (train.py:207-219) after two reference AdamW
   steps, with digests of the tensors (the checkpoint itself is rebuilt by the test from seeded inputs)."""
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import common as C  # noqa: E402
import make_golden as G  # noqa: E402

ref = G.ref
fc_map, render, sample, transform = ref["fc_map"], ref["render"], ref["sample"], ref["transform"]


describe = C.describe


def main():
    out = {}
    # ---- grid ----------------------------------------------------------------------------------------------
    T_box = C.rigid_transform(14)                       # world -> box (what trimesh.bounds.oriented_bounds returns)
    extents = torch.tensor([6.0, 2.5, 4.0])
    m = G.build_ref_map(C.golden_weights(71, gain=1.2), transform=T_box)
    dim = 12
    scale = extents / (2.0 * 0.9)
    grid_pc = transform.make_3D_grid([-1.0, 1.0], dim, "cpu", transform=torch.inverse(T_box), scale=scale).view(-1, 3)
    with torch.no_grad():
        sdf = fc_map.chunks(grid_pc, 500, m)
    out["grid"] = dict(grid_pc=grid_pc, sdf=sdf.view(dim, dim, dim), extents=extents)
    # ---- render ----------------------------------------------------------------------------------------------
    g = C.gen(72)
    z = torch.sort(torch.rand(16, 20, generator=g) * 4 + 0.1, dim=1).values
    s = torch.randn(16, 20, generator=g) * 0.3
    s[0] = s[0].abs() + 0.01           # never inside -> argmax 0
    s[1] = -s[1].abs() - 0.01          # always inside -> last index -> depth 0
    out["render_depth"] = dict(z=z, sdf=s, depth=render.sdf_render_depth(z, s))
    Tc = C.synthetic_pose(3)[None]
    dirs = transform.ray_dirs_C(1, 6, 8, 10.0, 10.0, 3.5, 2.5, "cpu", "z").view(1, -1, 3)
    depth = 1.0 + torch.rand(1, 48, generator=g)
    out["render_normals"] = dict(depth=depth, normals=render.render_normals(Tc, depth, m, dirs).detach())
    # ---- sample_along_rays without gt_depth ----------------------------------------------------------------------
    torch.manual_seed(5)
    pc1, z1 = sample.sample_along_rays(Tc, 0.07, 12.0, 20, 0, dirs, gt_depth=None)
    du = depth.view(-1)
    torch.manual_seed(6)
    pc2, z2 = sample.sample_along_rays(Tc, du - 0.1, du + 0.1, 12, 12, dirs)
    out["sample_render"] = dict(pc1=pc1, z1=z1, pc2=pc2, z2=z2)
    # ---- checkpoint ------------------------------------------------------------------------------------------
    m2 = G.build_ref_map(C.golden_weights(73))
    opt = torch.optim.AdamW(m2.parameters(), lr=0.0013, weight_decay=0.012)      # trainer.py:435-439
    gg = C.gen(74)
    for it in range(2):
        for p in m2.parameters():
            p.grad = torch.randn(p.shape, generator=gg) * 0.01
        opt.step()
    ck = {"step": 7.5, "model_state_dict": m2.state_dict(), "optimizer_state_dict": opt.state_dict(), "loss": 0.25}
    buf = io.BytesIO()
    torch.save(ck, buf)                                                          # train.py:207-219
    ck = torch.load(io.BytesIO(buf.getvalue()), weights_only=False)
    x = (torch.rand(8, 3, generator=gg) - 0.5) * 4
    with torch.no_grad():
        sdf_ck = m2(x)
    out["checkpoint"] = dict(
        structure=describe(ck),
        model_digest={k: (float(v.double().sum()), C.subsample(v, 997)) for k, v in ck["model_state_dict"].items()},
        opt_digest={i: {k: (float(v.double().sum()) if torch.is_tensor(v) else v) for k, v in st.items()}
                    for i, st in ck["optimizer_state_dict"]["state"].items()},
        x=x, sdf=sdf_ck)
    G.save("infer.pt", out)


if __name__ == "__main__":
    main()
