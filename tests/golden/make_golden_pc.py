"""Golden fixture for the 'pc' (batch-distance) bound, row N2 of SURVEY.md 8f: runs the UNMODIFIED
reference (loss.bounds('pc'), sdf_loss, tot_loss, backward) on CPU.   python tests/golden/make_golden_pc.py
Writes tests/golden/step_pc.pt; inputs are rebuilt by the tests from common.loss_batch_pc."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import common as C  # noqa: E402
import make_golden as G  # noqa: E402  (imports the reference through oracle/ref_shim.py)

loss, fc_map = G.loss, G.fc_map


def ref_loss_and_grads_pc(m, batch, noise, noise_std, loss_type):
    """Trainer.sdf_eval_and_loss (trainer.py:768-836) with bounds_method='pc', then backward."""
    cosSim = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
    pc = batch["pc"].clone().requires_grad_(True)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: noise.clone()[..., None]
    try:
        sdf = m(pc, noise_std=noise_std)
    finally:
        torch.randn = real_randn
    g = fc_map.gradient(pc, sdf)
    bounds, grad_vec = loss.bounds("pc", batch["dirs_C_sample"], batch["depth_sample"], batch["T_WC_sample"],
                                   batch["z_vals"], pc, G.CFG["trunc_distance"], batch["norm_sample"], do_grad=True)
    raw_vec = grad_vec.clone()
    sdf_loss_mat, free_ixs = loss.sdf_loss(sdf, bounds, G.CFG["trunc_distance"], loss_type=loss_type)
    eik = torch.abs(g.norm(2, dim=-1) - 1)
    surf = 1 - cosSim(g[:, 0], batch["norm_sample"])
    nan = torch.where(grad_vec[..., 0].isnan())
    grad_vec[nan] = batch["norm_sample"][nan[0]]
    gl = torch.cat((surf[:, None], 1 - cosSim(grad_vec, g[:, 1:])), dim=1)
    tot, tot_mat, losses = loss.tot_loss(sdf_loss_mat, gl, eik, free_ixs, bounds, G.CFG["eik_apply_dist"],
                                         G.CFG["trunc_weight"], G.CFG["grad_weight"], G.CFG["eik_weight"])
    for p in m.parameters():
        p.grad = None
    tot.backward()
    return sdf.detach(), g.detach(), bounds, raw_vec, tot_mat.detach(), losses, m


def main():
    out = {}
    for tag, seed, gain, tr, R, nstd, lt in (("p1", 61, 1.0, None, 48, 0.25, "L1"),
                                             ("p2_rigid_L2", 62, 1.5, C.rigid_transform(9), 32, 0.0, "L2")):
        m = G.build_ref_map(C.golden_weights(seed, gain=gain), transform=tr)
        batch, noise = C.loss_batch_pc(seed + 100, R)
        sdf, g, bounds, vec, tot_mat, losses, m = ref_loss_and_grads_pc(m, batch, noise, nstd, lt)
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        out[tag] = dict(sdf=sdf, grad=g, bounds=bounds, grad_vec=vec, total_mat=tot_mat,
                        losses={k: (float(v) if not torch.is_tensor(v) else float(v.item())) for k, v in losses.items()},
                        grad_norm={k: v.double().norm() for k, v in grads.items()},
                        grad_sub={k: (C.subsample(v) if v.numel() > 4096 else v.clone()) for k, v in grads.items()})
        print(tag, "NaN rows in grad_vec:", int(vec[..., 0].isnan().sum()), out[tag]["losses"])
    G.save("step_pc.pt", out)


if __name__ == "__main__":
    main()
