"""Shared, deterministic input builders for golden-vector generation and for the tests.

Everything here is derived from seeded torch CPU generators so that the generating
script (make_golden.py, run once in the build container against the real reference)
and the tests (run anywhere) see identical inputs without storing them.
"""
import math

import torch


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def state_dict_names(block):
    names = ["in_layer.0"] + ["mid1.%d.0" % i for i in range(block)] + ["cat_layer.0"] \
        + ["mid2.%d.0" % i for i in range(block)] + ["out_alpha"]
    return names


def golden_weights(seed, E=255, H=256, block=2, gain=1.0, dtype=torch.float32):
    """Xavier-normal-like weights / small uniform biases from an explicit generator."""
    g = gen(seed)
    shapes = [(H, E)] + [(H, H)] * block + [(H, H + E)] + [(H, H)] * block + [(1, H)]
    sd = {}
    for name, (o, i) in zip(state_dict_names(block), shapes):
        std = gain * math.sqrt(2.0 / (o + i))
        sd[name + ".weight"] = (torch.randn(o, i, generator=g) * std).to(dtype)
        bound = 1.0 / math.sqrt(i)
        sd[name + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * bound).to(dtype)
    return sd


def rigid_transform(seed):
    """A reproducible 4x4 rigid transform (rotation via QR, |t| ~ 1 m)."""
    g = gen(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))[None, :]
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = q
    T[:3, 3] = torch.randn(3, generator=g, dtype=torch.float64)
    return T.float()


def synthetic_depth(k, H, W, invalid_frac=0.0, seed=1234):
    """Depth image of keyframe k in metres (SURVEY.md 8d synthetic stream)."""
    v = torch.arange(H, dtype=torch.float32)[:, None]
    u = torch.arange(W, dtype=torch.float32)[None, :]
    d = 2.0 + 0.5 * torch.sin(u / 80.0 + 0.1 * k) + 0.3 * torch.cos(v / 60.0)
    if invalid_frac > 0:
        m = torch.rand(H, W, generator=gen(seed + k)) < invalid_frac
        d = torch.where(m, torch.zeros_like(d), d)
    return d.contiguous()


def synthetic_pose(k, rot=True):
    T = torch.eye(4)
    if rot:
        ang = 0.05 * k
        c, s = math.cos(ang), math.sin(ang)
        T[0, 0], T[0, 2], T[2, 0], T[2, 2] = c, s, -s, c
    T[0, 3] = 0.05 * k
    T[1, 3] = 0.01 * k
    return T


def synthetic_normals(H, W, nan_frac, seed):
    """Unit 'normals' with a fraction of NaN rows (the reference marks invalid normals NaN)."""
    g = gen(seed)
    n = torch.randn(H, W, 3, generator=g)
    n = n / n.norm(dim=-1, keepdim=True)
    if nan_frac > 0:
        m = torch.rand(H, W, generator=g) < nan_frac
        n[m] = float("nan")
    return n


def loss_batch(seed, R, S=27, n_surf=8, min_depth=0.07, dist_behind=0.1):
    """A pc-level batch (what Trainer.sample_points returns) built without any image."""
    g = gen(seed)
    depth = 1.0 + 3.0 * torch.rand(R, generator=g)
    dirs_C = torch.stack([(torch.rand(R, generator=g) - 0.5) * 1.6,
                          (torch.rand(R, generator=g) - 0.5) * 1.0,
                          torch.ones(R)], dim=-1)
    T = torch.stack([synthetic_pose(int(k)) for k in torch.randint(0, 5, (R,), generator=g)])
    nrm = torch.randn(R, 3, generator=g)
    nrm = nrm / nrm.norm(dim=-1, keepdim=True)
    n_strat = S - n_surf
    far = depth + dist_behind
    u = torch.rand(R, n_strat, generator=g)
    edges = torch.linspace(0, 1, n_strat + 1)[None, :] * (far - min_depth)[:, None] + min_depth
    z_strat = edges[:, :-1] + u * ((far - min_depth)[:, None] / n_strat)
    near = depth[:, None] + 0.1 * torch.randn(R, n_surf - 1, generator=g)
    near = torch.minimum(torch.maximum(near, torch.full_like(near, min_depth)), far[:, None])
    z = torch.cat([depth[:, None], near, z_strat], dim=1)
    dirs_W = (T[:, :3, :3] * dirs_C[:, None, :]).sum(-1)
    pc = T[:, :3, 3][:, None, :] + dirs_W[:, None, :] * z[:, :, None]
    noise = torch.randn(R, S, generator=g)
    return dict(pc=pc, z_vals=z, depth_sample=depth, dirs_C_sample=dirs_C, dirs_W=dirs_W,
                T_WC_sample=T, norm_sample=nrm), noise


def subsample(t, stride=97):
    return t.reshape(-1)[::stride].clone()


def loss_batch_pc(seed, R, S=27):
    """loss_batch for the 'pc' (batch-distance) bound: rays share two poses so that surface points of
    neighbouring rays are the closest ones, and one sample coincides with a surface point (its bounds_pc
    direction is NaN -> replaced by the ray's normal, trainer.py:823-824)."""
    batch, noise = loss_batch(seed, R, S=S)
    g = gen(seed + 7)
    k = torch.randint(0, 2, (R,), generator=g)
    T = torch.stack([synthetic_pose(int(i)) for i in k])
    z = batch["z_vals"].clone()
    z[1, 3] = batch["depth_sample"][1]                  # sample (1,3) == surface sample of ray 1
    dirs_W = (T[:, :3, :3] * batch["dirs_C_sample"][:, None, :]).sum(-1)
    pc = T[:, :3, 3][:, None, :] + dirs_W[:, None, :] * z[:, :, None]
    pc[1, 3] = pc[1, 0]
    batch.update(T_WC_sample=T, z_vals=z, dirs_W=dirs_W, pc=pc)
    return batch, noise


def describe(obj):
    """Nested structure with tensor shapes / dtypes instead of values."""
    if torch.is_tensor(obj):
        return ("tensor", tuple(obj.shape), str(obj.dtype))
    if isinstance(obj, dict):
        return {k: describe(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [describe(v) for v in obj]
    return obj
