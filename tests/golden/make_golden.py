"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:   python tests/golden/make_golden.py
Writes tests/golden/*.pt (small tensors only).  The inputs are rebuilt by the tests from
tests/golden/common.py (seeded generators), so only OUTPUTS of the reference are stored.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
import common as C  # noqa: E402

torch.set_num_threads(8)
ref = ref_shim.load()
fc_map, embedding, sample, loss = ref["fc_map"], ref["embedding"], ref["sample"], ref["loss"]

CFG = dict(scale_input=0.05937489, max_deg=5, block=2, hidden=256, scale_output=0.14,
           trunc_weight=5.38344020, trunc_distance=0.29365022, eik_weight=0.268,
           eik_apply_dist=0.1, grad_weight=0.018)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print("wrote", path, os.path.getsize(path), "bytes")


def build_ref_map(sd, transform=None, block=2, hidden=256, max_deg=5):
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        pe = embedding.PostionalEncoding(min_deg=0, max_deg=max_deg, scale=CFG["scale_input"],
                                         transform=transform)
        m = fc_map.SDFMap(pe, hidden_size=hidden, hidden_layers_block=block,
                          scale_output=CFG["scale_output"])
    m.load_state_dict(sd)
    return m


# ---- G1: positional encoding ------------------------------------------------
def g_pe():
    import io, contextlib
    x = (torch.rand(64, 3, generator=C.gen(11)) - 0.5) * torch.tensor([12.0, 4.0, 12.0])
    out = {}
    for tag, tr in (("plain", None), ("rigid", C.rigid_transform(5))):
        with contextlib.redirect_stdout(io.StringIO()):
            pe = embedding.PostionalEncoding(min_deg=0, max_deg=5, scale=CFG["scale_input"], transform=tr)
        out[tag] = pe(x)
    with contextlib.redirect_stdout(io.StringIO()):
        pe8 = embedding.PostionalEncoding(min_deg=0, max_deg=8, scale=0.04, transform=None)
    out["deg8"] = pe8(x)
    save("pe.pt", out)


# ---- G2: SDFMap forward + input gradient; init parity ------------------------
def g_map():
    out = {}
    x = (torch.rand(96, 3, generator=C.gen(12)) - 0.5) * torch.tensor([12.0, 4.0, 12.0])
    for tag, seed, gain, tr in (("g1", 21, 1.0, None), ("g2_rigid", 22, 2.0, C.rigid_transform(6))):
        m = build_ref_map(C.golden_weights(seed, gain=gain), transform=tr)
        xx = x.clone().requires_grad_(True)
        sdf = m(xx)
        g = fc_map.gradient(xx, sdf)
        out[tag] = dict(sdf=sdf.detach(), grad=g.detach())
    # init parity: the reference's own constructor under torch.manual_seed(0)
    import io, contextlib
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        pe = embedding.PostionalEncoding(min_deg=0, max_deg=5, scale=CFG["scale_input"])
        m0 = fc_map.SDFMap(pe, 256, 2, 0.14)
    sd0 = m0.state_dict()
    out["init_seed0"] = {k: dict(shape=tuple(v.shape), sum=v.double().sum(), abs=v.double().abs().sum(),
                                 head=v.reshape(-1)[:4].clone()) for k, v in sd0.items()}
    xk = torch.tensor([[1., -2., 3.], [0., 0., 0.], [-4., 0.5, 2.5]], requires_grad=True)
    sk = m0(xk)
    out["init_seed0_kat"] = dict(sdf=sk.detach(), grad=fc_map.gradient(xk, sk).detach())
    save("sdfmap.pt", out)


# ---- G3: ray / depth sampling ------------------------------------------------
def g_sample():
    F, H, W = 3, 32, 48
    depth = torch.stack([C.synthetic_depth(k, H, W, invalid_frac=0.15) for k in range(F)])
    T = torch.stack([C.synthetic_pose(k) for k in range(F)])
    nrm = torch.stack([C.synthetic_normals(H, W, 0.1, 40 + k) for k in range(F)])
    cam = dict(fx=40.0, fy=42.0, cx=23.5, cy=15.5)
    dirs_C = ref["transform"].ray_dirs_C(1, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"], "cpu", "z")
    n_rays = 50
    torch.manual_seed(7)
    ib, ih, iw = sample.sample_pixels(n_rays, F, H, W, "cpu")
    (dC, dS, nS, TS, masks, ib2, ih2, iw2) = sample.get_batch_data(depth, T, dirs_C, ib, ih, iw,
                                                                   norm_batch=nrm, get_masks=True)
    torch.manual_seed(8)
    pc, z = sample.sample_along_rays(TS, 0.07, dS + 0.1, 19, 8, dC, gt_depth=dS, grad=False)
    torch.manual_seed(8)
    u = torch.rand(dS.shape[0], 19)
    nn_ = torch.normal(torch.zeros(dS.shape[0], 7), 0.1)
    save("sample.pt", dict(ib=ib, ih=ih, iw=iw, ib2=ib2, ih2=ih2, iw2=iw2, dirs_C=dC, depth=dS,
                           norm=nS, T=TS, masks_sum=masks.sum(dim=(1, 2)), pc=pc, z=z, u=u, n_near=nn_))


# ---- G4: losses, double back-prop gradients ----------------------------------
def ref_loss_and_grads(m, batch, noise, noise_std, loss_type="L1"):
    """Mirrors Trainer.sdf_eval_and_loss + backward using the reference's own functions."""
    cosSim = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
    pc = batch["pc"].clone().requires_grad_(True)
    # SDFMap.forward draws its own noise; reproduce by monkeypatching randn for this call
    real_randn = torch.randn
    torch.randn = lambda *a, **k: noise.clone()[..., None]
    try:
        sdf = m(pc, noise_std=noise_std)
    finally:
        torch.randn = real_randn
    g = fc_map.gradient(pc, sdf)
    bounds, grad_vec = loss.bounds("ray", batch["dirs_C_sample"], batch["depth_sample"],
                                   batch["T_WC_sample"], batch["z_vals"], pc,
                                   CFG["trunc_distance"], batch["norm_sample"], do_grad=True)
    sdf_loss_mat, free_ixs = loss.sdf_loss(sdf, bounds, CFG["trunc_distance"], loss_type=loss_type)
    eik = torch.abs(g.norm(2, dim=-1) - 1)
    surf = 1 - cosSim(g[:, 0], batch["norm_sample"])
    gl = 1 - cosSim(grad_vec, g[:, 1:])
    gl = torch.cat((surf[:, None], gl), dim=1)
    tot, tot_mat, losses = loss.tot_loss(sdf_loss_mat, gl, eik, free_ixs, bounds, CFG["eik_apply_dist"],
                                         CFG["trunc_weight"], CFG["grad_weight"], CFG["eik_weight"])
    for p in m.parameters():
        p.grad = None
    tot.backward()
    return sdf.detach(), g.detach(), bounds, tot_mat.detach(), losses, m


def g_step():
    out = {}
    cases = (("c1", 31, 1.0, None, 48, 0.25, "L1"),
             ("c2_rigid_gain2", 32, 2.0, C.rigid_transform(9), 40, 0.04, "L1"),
             ("c3_L2", 33, 1.5, None, 24, 0.0, "L2"))
    for tag, seed, gain, tr, R, nstd, lt in cases:
        sd = C.golden_weights(seed, gain=gain)
        m = build_ref_map(sd, transform=tr)
        batch, noise = C.loss_batch(seed + 100, R)
        sdf, g, bounds, tot_mat, losses, m = ref_loss_and_grads(m, batch, noise, nstd, lt)
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        rec = dict(sdf=sdf, grad=g, bounds=bounds, total_mat=tot_mat,
                   losses={k: (float(v) if not torch.is_tensor(v) else float(v.item())) for k, v in losses.items()},
                   grad_norm={k: v.double().norm() for k, v in grads.items()},
                   grad_sub={k: (C.subsample(v) if v.numel() > 4096 else v.clone()) for k, v in grads.items()})
        # frame average on a fake pixel assignment (with duplicates)
        F, H, W = 4, 16, 24
        gq = C.gen(seed + 200)
        ib = torch.randint(0, F, (R,), generator=gq).sort().values
        ih = torch.randint(0, H, (R,), generator=gq)
        iw = torch.randint(0, W, (R,), generator=gq)
        ih[1], iw[1], ib[1] = ih[0], iw[0], ib[0]          # forced duplicate pixel
        masks = torch.zeros(F, H, W)
        masks[ib, ih, iw] = 1
        la, fa = loss.frame_avg(tot_mat, torch.zeros(F, H, W), ib, ih, iw, W, H, 8, masks)
        rec.update(frame_ib=ib, frame_ih=ih, frame_iw=iw, loss_approx=la, frame_avg=fa)
        out[tag] = rec
    save("step.pt", out)


# ---- G6: AdamW ---------------------------------------------------------------
def g_adamw():
    g = C.gen(51)
    p = torch.nn.Parameter(torch.randn(1000, generator=g))
    opt = torch.optim.AdamW([p], lr=0.0013, weight_decay=0.012)
    traj = []
    for it in range(3):
        p.grad = torch.randn(1000, generator=g) * 0.01
        opt.step()
        traj.append(p.detach().clone())
    save("adamw.pt", dict(traj=torch.stack(traj)))


if __name__ == "__main__":
    g_pe(); g_map(); g_sample(); g_step(); g_adamw()
