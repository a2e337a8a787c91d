"""Forward-only rendering helpers -- mirror of reference isdf/modules/render.py (keyframe test, vis).
Row N1 of SURVEY.md 8f: the SDF / gradient evaluations go through K2 / K3."""
import torch

from . import fc_map


def sdf_render_depth(z_vals, sdf):
    """Depth of the last inside->outside transition along sorted samples (render.py:12-35)."""
    n = sdf.size(1)
    weights = (sdf < 0) * torch.arange(n, 0, -1, device=sdf.device)
    ix = weights.argmax(dim=1)
    rows = torch.arange(z_vals.size(0), device=sdf.device)
    depths = z_vals[rows, ix] + sdf[rows, ix]
    return torch.where(ix == n - 1, torch.zeros_like(depths), depths)


def render_normals(T_WC, render_depth, sdf_map, dirs_C):
    """Camera-frame surface normals from the SDF gradient at the rendered depth (render.py:39-57)."""
    R_WC = T_WC[:, :3, :3]
    dirs_W = (R_WC * dirs_C[..., None, :]).sum(dim=-1).view(-1, 3)
    origins = T_WC[:, :3, -1].view(-1, 3)
    pc = (origins + dirs_W * render_depth.flatten()[:, None]).requires_grad_()
    sdf = sdf_map(pc)
    g = fc_map.gradient(pc, sdf).detach()
    n_W = -g / (g.norm(dim=1, keepdim=True) + 1e-4)
    n_C = (R_WC.inverse() * n_W[..., None, :]).sum(dim=-1)
    return n_C.view(render_depth.shape[0], render_depth.shape[1], 3)


def render_weighted(weights, vals, dim=-1, normalise=False):
    """Weighted sum along `dim`, optionally divided by the number of samples (render.py:60-71)."""
    out = (weights * vals).sum(dim=dim)
    return out / weights.size(dim) if normalise else out
