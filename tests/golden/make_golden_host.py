"""Golden fixtures for the host-side helpers (plain torch / numpy, no kernels), from the UNMODIFIED reference on CPU.
   python tests/golden/make_golden_host.py  ->  tests/golden/host.pt
Covers transform.{ray_dirs_C, origin_dirs_W, pointcloud_from_depth(_torch), backproject_pointclouds, pc_bounds,
estimate_pointcloud_normals, normalize}, loss.{bounds_ray, grad_ray, sdf_loss, full_sdf_loss, tsdf_loss, tot_loss,
approx_loss}, render.render_weighted, sample.stratified_sample and the FrameData append / replace semantics."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import common as C  # noqa: E402
import make_golden as G  # noqa: E402

ref = G.ref
T, L, R, S, D = ref["transform"], ref["loss"], ref["render"], ref["sample"], ref["data_util"]


def main():
    out = {}
    g = C.gen(81)
    H, W = 12, 16
    cam = (20.0, 21.0, 7.5, 5.5)
    out["ray_dirs_z"] = T.ray_dirs_C(2, H, W, *cam, "cpu", "z")
    out["ray_dirs_e"] = T.ray_dirs_C(1, H, W, *cam, "cpu", "euclidean")
    dirs = out["ray_dirs_z"].view(2, -1, 3)[0, :40]                    # per-ray form [R,3] with [R,4,4]
    Tw = torch.stack([C.synthetic_pose(k % 5) for k in range(40)])
    out["origin_dirs_W"] = T.origin_dirs_W(Tw, dirs)
    out["origin_dirs_W_one_pose"] = T.origin_dirs_W(Tw[3:4], out["ray_dirs_z"].view(2, -1, 3)[:1])
    depth = C.synthetic_depth(2, H, W)
    depth[3, 4] = float("nan")
    out["pc_torch"] = T.pointcloud_from_depth_torch(depth, *cam)
    out["pc_torch_e_skip2"] = T.pointcloud_from_depth_torch(depth, *cam, depth_type="euclidean", skip=2)
    out["pc_np"] = T.pointcloud_from_depth(depth.numpy(), *cam)
    out["backproject"] = T.backproject_pointclouds(np.stack([depth.numpy(), 2 * depth.numpy()]), *cam)
    pts = torch.randn(50, 3, generator=g).numpy()
    out["pc_bounds"] = T.pc_bounds(pts)
    out["normals"] = T.estimate_pointcloud_normals(T.pointcloud_from_depth_torch(C.synthetic_depth(1, 24, 32), 30., 30., 15.5, 11.5))
    out["normalize"] = T.normalize(np.array([3.0, -4.0, 12.0]))
    batch, _ = C.loss_batch(82, 20)
    b, gv = L.bounds_ray(batch["depth_sample"], batch["z_vals"], batch["dirs_C_sample"], batch["T_WC_sample"], True)
    out["bounds_ray"] = (b, gv)
    sdf = torch.randn(20, 27, generator=g) * 0.3
    out["sdf"] = sdf
    for lt in ("L1", "L2"):
        out["sdf_loss_" + lt] = L.sdf_loss(sdf, b, 0.29365022, loss_type=lt)
    out["full_sdf_loss"] = L.full_sdf_loss(sdf, b)
    out["tsdf_loss"] = L.tsdf_loss(sdf, b, 0.3)
    gl = torch.rand(20, 27, generator=g)
    ek = torch.rand(20, 27, generator=g)
    mat, free = out["sdf_loss_L1"]
    tot, tot_mat, losses = L.tot_loss(mat.clone(), gl, ek, free, b, 0.1, 5.38344020, 0.018, 0.268)
    out["tot_loss"] = (tot, tot_mat, {k: float(v) for k, v in losses.items()})
    full = torch.rand(3, 16, 24, generator=g)
    masks = (torch.rand(3, 16, 24, generator=g) < 0.2).float()
    out["approx_loss"] = L.approx_loss(full * masks, masks.clone(), 24, 16, 8)
    w = torch.rand(5, 9, generator=g)
    v = torch.rand(5, 9, generator=g)
    out["render_weighted"] = (R.render_weighted(w, v), R.render_weighted(w, v, normalise=True))
    torch.manual_seed(9)
    out["strat_scalar"] = S.stratified_sample(0.07, 5.0, 6, "cpu", 11)
    torch.manual_seed(9)
    out["strat_tensor"] = S.stratified_sample(0.07, torch.linspace(1, 3, 6), 6, "cpu", 11)
    # FrameData append / replace (data_util.py:11-102)
    fd = D.FrameData()
    log = []
    for k, rep in ((0, False), (1, False), (2, True), (3, False), (4, True)):
        d_ = D.FrameData(frame_id=np.array([k]), im_batch=torch.full((1, 2, 3, 3), float(k)), im_batch_np=np.full((1, 2, 3, 3), k, np.uint8),
                         depth_batch=torch.full((1, 2, 3), float(k)), depth_batch_np=np.full((1, 2, 3), k, np.float32),
                         T_WC_batch=torch.eye(4)[None] * k, T_WC_batch_np=np.eye(4, dtype=np.float32)[None] * k,
                         normal_batch=torch.full((1, 2, 3, 3), float(k)))
        fd.add_frame_data(d_, replace=rep)
        log.append((len(fd), fd.frame_id.copy(), fd.depth_batch[:, 0, 0].clone(), fd.im_batch_np[:, 0, 0, 0].copy(),
                    fd.frame_avg_losses.clone(), fd.T_WC_batch_np[:, 0, 0].copy()))
    out["framedata"] = log
    G.save("host.pt", out)


if __name__ == "__main__":
    main()
