"""Loss helpers -- host-side mirror of reference isdf/modules/loss.py.

On the training path all of this is fused into K4 (isdfb_train_fwd_bwd); `frame_avg` runs K5.
The tensor-level helpers below exist so callers written against the reference's API (eval, vis)
keep working; they are plain torch expressions and are NOT used by Trainer.step()."""
import torch

from . import sample as _sample

cosSim = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)


def grad_ray(T_WC_sample, dirs_C_sample, n_samples):
    """Negative viewing direction per sample (loss.py:48-53)."""
    dirs_W = (T_WC_sample[:, :3, :3] * dirs_C_sample[..., None, :]).sum(dim=-1)
    return -dirs_W[:, None, :].repeat(1, n_samples, 1)


def bounds_ray(depth_sample, z_vals, dirs_C_sample, T_WC_sample, do_grad):
    """Upper bound on |sdf| along the ray: ||d_C|| (depth - z)  (loss.py:13-22)."""
    b = dirs_C_sample.norm(dim=-1)[:, None] * (depth_sample[:, None] - z_vals)
    return b, (grad_ray(T_WC_sample, dirs_C_sample, z_vals.shape[1] - 1) if do_grad else None)


def bounds_pc(pc, z_vals, depth_sample, do_grad=True):
    """'Batch distance' bound: distance to the closest surface sample of the batch (loss.py:56-89)."""
    with torch.no_grad():
        surf = pc[:, 0]
        diff = pc[:, :, None] - surf
        dists, closest = diff.norm(dim=-1).min(dim=-1)
        behind = z_vals > depth_sample[:, None]
        dists = torch.where(behind, -dists, dists)
        grad = None
        if do_grad:
            idx = closest[..., None, None].expand(-1, -1, 1, 3)
            grad = torch.gather(diff, 2, idx).squeeze(2)[:, 1:]
            grad = grad / grad.norm(dim=-1, keepdim=True)
            grad = torch.where(behind[:, 1:, None], -grad, grad)
    return dists, grad


def bounds(method, dirs_C_sample, depth_sample, T_WC_sample, z_vals, pc, normal_trunc_dist, norm_sample,
           do_grad=True):
    assert method in ["ray", "normal", "pc"]
    if method == "ray":
        return bounds_ray(depth_sample, z_vals, dirs_C_sample, T_WC_sample, do_grad)
    if method == "pc":
        return bounds_pc(pc, z_vals, depth_sample, do_grad)
    # the reference's bounds_normal calls bounds_ray with 3 of 5 arguments (loss.py:29) and raises
    raise TypeError("bounds_method 'normal' is broken in the reference (loss.py:29) and not supported")


def full_sdf_loss(sdf, target_sdf, free_space_factor=5.0):
    free = torch.maximum(torch.relu(sdf - target_sdf), torch.exp(-free_space_factor * sdf) - 1.)
    return free, sdf - target_sdf


def tsdf_loss(sdf, target_sdf, trunc_dist):
    """TSDF-style residuals: sdf - 1 in free space, sdf - target / trunc_dist in the band (loss.py:167-175)."""
    return sdf - torch.ones_like(sdf), sdf - target_sdf / trunc_dist


def approx_loss(full_loss, binary_masks, W, H, factor=8):
    """[F,H,W] loss image and pixel mask -> [F,factor,factor] block means over the sampled pixels
    (loss.py:208-218).  Dense-image form kept for callers that have such images; Trainer.step() uses K5."""
    hb, wb = H // factor, W // factor
    sums = full_loss.view(-1, factor, hb, factor, wb).sum(dim=(2, 4))
    n = binary_masks.view(-1, factor, hb, factor, wb).sum(dim=(2, 4))
    return sums / torch.where(n == 0, torch.ones_like(n), n)


def sdf_loss(sdf, bounds, t, loss_type="L1"):
    """Free-space loss where bounds > t, direct supervision inside the truncation band (loss.py:122-145)."""
    free, trunc = full_sdf_loss(sdf, bounds)
    free_space_ixs = bounds > t
    mat = torch.where(free_space_ixs, free, trunc)
    if loss_type == "L1":
        mat = mat.abs()
    elif loss_type == "L2":
        mat = mat.square()
    else:
        raise ValueError("Must be L1 or L2")
    return mat, free_space_ixs


def tot_loss(sdf_loss_mat, grad_loss_mat, eik_loss_mat, free_space_ixs, bounds, eik_apply_dist, trunc_weight,
             grad_weight, eik_weight):
    """Weighted total and the scalar report (loss.py:178-205)."""
    sdf_loss_mat = torch.where(free_space_ixs, sdf_loss_mat, sdf_loss_mat * trunc_weight)
    losses = {"sdf_loss": sdf_loss_mat.mean().item()}
    tot = sdf_loss_mat
    if grad_loss_mat is not None:
        tot = tot + grad_weight * grad_loss_mat
        losses["grad_loss"] = grad_loss_mat.mean().item()
    if eik_loss_mat is not None:
        eik = torch.where(bounds < eik_apply_dist, torch.zeros_like(eik_loss_mat), eik_loss_mat) * eik_weight
        tot = tot + eik
        losses["eikonal_loss"] = eik.mean().item()
    total = tot.mean()
    losses["total_loss"] = total
    return total, tot, losses


def frame_avg(total_loss_mat, depth_batch, indices_b, indices_h, indices_w, W, H, loss_approx_factor,
              binary_masks=None, engine=None):
    """Per-frame factor x factor loss histogram and its mean (loss.py:221-240) -- K5, no [F,H,W] images."""
    eng = engine or _sample._engine(total_loss_mat.device)
    n_frames = depth_batch.shape[0] if torch.is_tensor(depth_batch) else int(depth_batch)
    return eng.frame_bins(total_loss_mat.detach(), indices_b, indices_h, indices_w, n_frames, H, W,
                          loss_approx_factor)
