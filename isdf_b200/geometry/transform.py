"""Geometry helpers on the boundary of the hot path -- mirror of the used subset of reference
isdf/geometry/transform.py (ray_dirs_C 13-33, origin_dirs_W 36-41, transform_3D_grid 287-304,
pointcloud_from_depth_torch 169-196, estimate_pointcloud_normals 215-270).

The training kernels recompute ray directions from pixel indices; these torch versions serve the
callers around the step (frame ingest, visualisation) and are not used inside Trainer.step()."""
import numpy as np
import torch


def ray_dirs_C(B, H, W, fx, fy, cx, cy, device, depth_type='z'):
    """Per-pixel camera-frame directions [B,H,W,3]: ((c-cx)/fx, (r-cy)/fy, 1), optionally unit length."""
    c = torch.arange(W, device=device, dtype=torch.float32)
    r = torch.arange(H, device=device, dtype=torch.float32)
    x = ((c - cx) / fx)[None, :].expand(H, W)
    y = ((r - cy) / fy)[:, None].expand(H, W)
    dirs = torch.stack((x, y, torch.ones(H, W, device=device)), dim=-1)[None].repeat(B, 1, 1, 1)
    if depth_type == 'euclidean':
        dirs = dirs / dirs.norm(dim=3, keepdim=True)
    return dirs


def origin_dirs_W(T_WC, dirs_C):
    """World-frame ray origins and directions: t_WC and R_WC d_C."""
    dirs_W = (T_WC[:, :3, :3] * dirs_C[..., None, :]).sum(dim=-1)
    return T_WC[:, :3, -1], dirs_W


def normalize(x):
    """Unit vector of a 1-D numpy array (transform.py:44-46)."""
    assert x.ndim == 1, "x must be a vector (ndim: 1)"
    return x / np.linalg.norm(x)


def pointcloud_from_depth(depth, fx, fy, cx, cy, depth_type="z", skip=1):
    """numpy back-projection of a float depth image [H,W] -> [H/skip,W/skip,3]; NaN stays NaN (transform.py:141-166)."""
    assert depth_type in ["z", "euclidean"], "Unexpected depth_type"
    assert depth.dtype.kind == "f", "depth must be float and have meter values"
    rows, cols = depth.shape
    c = np.arange(cols, step=skip)[None, :]
    r = np.arange(rows, step=skip)[:, None]
    z = depth[::skip, ::skip]
    bad = np.isnan(z)
    x = np.where(bad, np.nan, z * (c - cx) / fx)
    y = np.where(bad, np.nan, z * (r - cy) / fy)
    pc = np.dstack((x, y, z))
    if depth_type == "euclidean":
        pc = pc * (z / np.linalg.norm(pc, axis=2))[:, :, None]
    return pc


def backproject_pointclouds(depths, fx, fy, cx, cy):
    """[B,H,W] depth images -> [B,H*W,3] camera-frame points (transform.py:127-138)."""
    return np.stack([pointcloud_from_depth(d, fx, fy, cx, cy).reshape(-1, 3) for d in depths], axis=0)


def pc_bounds(pc):
    """Axis-aligned extents and centre of an [N,3] point array (transform.py:199-212)."""
    lo, hi = pc[:, :3].min(axis=0), pc[:, :3].max(axis=0)
    return hi - lo, (hi + lo) / 2.0


def transform_3D_grid(grid_3d, transform=None, scale=None):
    """scale first, then the rigid transform R x + t (applies to the last dimension)."""
    if scale is not None:
        grid_3d = grid_3d * scale
    if transform is not None:
        tr = torch.as_tensor(transform, dtype=grid_3d.dtype, device=grid_3d.device)
        grid_3d = grid_3d @ tr[:3, :3].T + tr[:3, 3]
    return grid_3d


def make_3D_grid(grid_range, dim, device, transform=None, scale=None):
    t = torch.linspace(grid_range[0], grid_range[1], steps=dim, device=device)
    g = torch.stack(torch.meshgrid(t, t, t, indexing="ij"), dim=3)
    return transform_3D_grid(g, transform=transform, scale=scale)


def pointcloud_from_depth_torch(depth, fx, fy, cx, cy, depth_type="z", skip=1):
    """Back-project a depth image [H,W] to camera-frame points [H,W,3]; NaN depth stays NaN."""
    assert depth_type in ["z", "euclidean"], "Unexpected depth_type"
    depth = depth[::skip, ::skip]
    rows, cols = depth.shape
    c = torch.arange(0, cols * skip, skip, device=depth.device)[None, :]
    r = torch.arange(0, rows * skip, skip, device=depth.device)[:, None]
    z = depth
    x = z * (c - cx) / fx
    y = z * (r - cy) / fy
    pc = torch.dstack((x, y, z))
    if depth_type == "euclidean":
        pc = pc * (z / torch.linalg.norm(pc, dim=2))[:, :, None]
    return pc


def estimate_pointcloud_normals(points):
    """Pixel normals from an organised point cloud [H,W,3]: for each pixel pick, among the 8
    neighbour pairs (k, k+2) at distance 2, the pair closest to the anchor and cross their offsets."""
    assert points.shape[2] == 3
    d = 2
    H, W = points.shape[:2]
    nan = float('nan')
    padded = torch.nn.functional.pad(points, pad=(0, 0, d, d, d, d), mode="constant", value=nan)
    offs = [(-d, 0), (-d, d), (0, d), (d, d), (d, 0), (d, -d), (0, -d), (-d, -d)]

    def shifted(dy, dx):
        return padded[d + dy:d + dy + H, d + dx:d + dx + W]

    nb = torch.stack([shifted(dy, dx) for dy, dx in offs])              # [8,H,W,3]
    p2 = nb - points[None]
    p3 = torch.roll(nb, shifts=-2, dims=0) - points[None]
    cost = torch.linalg.norm(p2, dim=3) + torch.linalg.norm(p3, dim=3)
    cost = torch.where(torch.isnan(cost), torch.full_like(cost, float('inf')), cost)
    best = torch.argmin(cost, dim=0)[None, :, :, None].expand(1, H, W, 3)
    a = torch.gather(p2, 0, best)[0]
    b = torch.gather(p3, 0, best)[0]
    normals = torch.linalg.cross(a, b, dim=2)
    return normals / torch.linalg.norm(normals, dim=2, keepdim=True)
