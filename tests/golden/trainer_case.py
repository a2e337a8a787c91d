"""Synthetic ReplicaCAD-format sequence + config shared by make_trainer_golden.py (reference, CPU)
and tests/test_gpu_trainer.py (isdf_b200, GPU).  Frames are written to disk so that both
implementations ingest byte-identical PNGs through their own readers."""
import json
import os

import numpy as np

H, W = 120, 160
N_FRAMES = 8
DEPTH_SCALE = 3276.75
CAM = dict(w=W, h=H, fx=100.0, fy=100.0, cx=79.5, cy=59.5)
SCHEDULE = [(k, 2) for k in range(7)]        # (add frame k, then run 2 steps)


def pose(k):
    T = np.eye(4)
    a = 0.04 * k
    c, s = np.cos(a), np.sin(a)
    T[0, 0], T[0, 2], T[2, 0], T[2, 2] = c, s, -s, c
    T[0, 3], T[1, 3], T[2, 3] = 0.05 * k, 0.01 * k, -0.02 * k
    return T


def depth_u16(k):
    v = np.arange(H, dtype=np.float64)[:, None]
    u = np.arange(W, dtype=np.float64)[None, :]
    d = 2.0 + 0.5 * np.sin(u / 20.0 + 0.3 * k) + 0.3 * np.cos(v / 15.0)
    rng = np.random.default_rng(1234 + k)
    d[rng.random((H, W)) < 0.05] = 0.0
    return np.round(d * DEPTH_SCALE).astype(np.uint16)


def write_sequence(root):
    import cv2
    seq = os.path.join(root, "synth_seq")
    res = os.path.join(seq, "results")
    os.makedirs(res, exist_ok=True)
    traj = []
    for k in range(N_FRAMES):
        cv2.imwrite(os.path.join(res, "depth%06d.png" % k), depth_u16(k))
        cv2.imwrite(os.path.join(res, "frame%06d.png" % k), np.full((H, W, 3), 90 + k, dtype=np.uint8))
        traj.append(pose(k).reshape(-1))
    np.savetxt(os.path.join(seq, "traj.txt"), np.array(traj))
    return seq + "/"


def config(seq_dir, base_json=None):
    cfg = {
        "dataset": {"format": "replicaCAD", "seq_dir": seq_dir, "noisy_depth": 0, "depth_scale": DEPTH_SCALE,
                    "fps": 30, "camera": CAM},
        "eval": {"do_vox_comparison": 0, "do_eval": 0, "eval_freq_s": 1, "sdf_eval": 1, "mesh_eval": 0},
        "save": {"save_period": 10, "save_checkpoints": 0, "save_slices": 0, "save_meshes": 0},
        "optimiser": {"lr": 0.0013, "weight_decay": 0.012},
        "trainer": {"steps": 100},
        "sample": {"n_rays": 40, "n_rays_is_kf": 80, "n_strat_samples": 19, "n_surf_samples": 8,
                   "depth_range": [0.07, 12.0], "dist_behind_surf": 0.1},
        "model": {"refine_poses": 0, "do_active": 0, "frac_time_perception": 1.0, "scale_output": 0.14,
                  "noise_std": 0.25, "noise_kf": 0.08, "noise_frame": 0.04, "window_size": 5,
                  "hidden_layers_block": 2, "hidden_feature_size": 256, "iters_per_kf": 60, "iters_per_frame": 10,
                  "kf_dist_th": 0.1, "kf_pixel_ratio": 0.65,
                  "embedding": {"scale_input": 0.05937489, "n_embed_funcs": 5, "gauss_embed": 0,
                                "gauss_embed_std": 11, "optim_embedding": 0}},
        "loss": {"bounds_method": "ray", "loss_type": "L1", "trunc_weight": 5.38344020,
                 "trunc_distance": 0.29365022, "eik_weight": 0.268, "eik_apply_dist": 0.1, "grad_weight": 0.018,
                 "orien_loss": 0},
        "pose_refine": {"pose_lr": 0.0004},
    }
    return cfg


def run_schedule(trainer_cls, device, cfg_path, probe, **kw):
    """Drive a Trainer like train.py does (add frame, optimise) and record what step() returns."""
    import torch
    np.random.seed(1)
    torch.manual_seed(1)
    tr = trainer_cls(device, cfg_path, incremental=True, **kw)
    rec = []
    for k, n_steps in SCHEDULE:
        tr.last_is_keyframe = True
        tr.add_data(tr.get_data([k]))
        for _ in range(n_steps):
            losses, _ = tr.step()
            rec.append({n: float(v) for n, v in losses.items()})
    with torch.no_grad():
        sdf = tr.sdf_map(probe.to(device)).cpu()
    sums = {n: float(p.detach().double().sum()) for n, p in tr.sdf_map.named_parameters()}
    favg = tr.frames.frame_avg_losses.detach().cpu().clone()
    return dict(losses=rec, probe_sdf=sdf, param_sums=sums, frame_avg_losses=favg)
