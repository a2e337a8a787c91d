"""TEST / BENCH INFRASTRUCTURE ONLY -- time the UNMODIFIED reference `Trainer.step()` on the host CPU.

The reference package (oracle/_ref/isdf, a verbatim copy made by oracle/make_ref.py, or /root/reference itself in the
build container) is imported through oracle/ref_shim.py, which only stubs the GUI / mesh libraries that the training
step never touches.  The Trainer is driven the way isdf/train/train.py:102-136 drives it: a ReplicaCAD-format sequence
on disk (written here with the same synthetic depth / pose formulas as isdf_b200.datasets.dataset.SyntheticStream),
`get_data` -> `add_data` per keyframe, then `step()`.  Used by bench.py (`--impl reference`, `cpu_baseline`) only."""
import contextlib
import io
import json
import math
import os
import tempfile
import time

import numpy as np
import torch

from . import ref_shim


def write_sequence(root, H, W, n_frames, depth_scale=1000.0):
    """<root>/seq/results/depth%06d.png (uint16), frame%06d.png, <root>/seq/traj.txt (reference dataset.py:20-71)."""
    import cv2
    seq = os.path.join(root, "seq")
    res = os.path.join(seq, "results")
    os.makedirs(res, exist_ok=True)
    v = np.arange(H, dtype=np.float32)[:, None]
    u = np.arange(W, dtype=np.float32)[None, :]
    traj = []
    for k in range(n_frames):
        d = (2.0 + 0.5 * np.sin(u / 80.0 + 0.1 * k) + 0.3 * np.cos(v / 60.0)).astype(np.float32)
        cv2.imwrite(os.path.join(res, "depth%06d.png" % k), np.round(d * depth_scale).astype(np.uint16))
        cv2.imwrite(os.path.join(res, "frame%06d.png" % k), np.full((H, W, 3), 128, dtype=np.uint8))
        T = np.eye(4)
        a = 0.05 * k
        c, s = math.cos(a), math.sin(a)
        T[0, 0], T[0, 2], T[2, 0], T[2, 2] = c, s, -s, c
        T[0, 3], T[1, 3] = 0.05 * k, 0.01 * k
        traj.append(T.reshape(-1))
    np.savetxt(os.path.join(seq, "traj.txt"), np.array(traj))
    return seq + "/"


class RefTrainerStepper:
    """The reference Trainer on device 'cpu' with `n_keyframes` resident keyframes of the given workload config."""

    def __init__(self, cfg, n_keyframes=8, seed=1, n_rays=None):
        ref = ref_shim.load()
        self.tmp = tempfile.mkdtemp(prefix="isdf_ref_")
        cam = cfg["dataset"]["camera"]
        cfg = json.loads(json.dumps(cfg))
        cfg.pop("b200", None)
        cfg["dataset"].update(format="replicaCAD", noisy_depth=0, depth_scale=1000.0,
                              seq_dir=write_sequence(self.tmp, cam["h"], cam["w"], n_keyframes))
        if n_rays is not None:
            cfg["sample"]["n_rays"] = int(n_rays)
        path = os.path.join(self.tmp, "cfg.json")
        json.dump(cfg, open(path, "w"))
        np.random.seed(seed)
        torch.manual_seed(seed)                       # the drivers' seeding (train.py:285-287)
        with contextlib.redirect_stdout(io.StringIO()):
            self.tr = ref["trainer"].Trainer("cpu", path, incremental=True)
            for k in range(n_keyframes):
                self.tr.last_is_keyframe = True
                self.tr.add_data(self.tr.get_data([k]))
                # one optimisation step per new keyframe, as the driver loop interleaves them (train.py:122-136): the
                # reference's loss-weighted window draw divides by the sum of the per-keyframe losses, which are zero
                # until a step has scored the frame (trainer.py:652-669)
                self.tr.step()
        self.trainer_file = ref["trainer"].__file__
        self.points_per_step = (cfg["sample"]["n_rays"] * min(n_keyframes, cfg["model"]["window_size"]) *
                                (cfg["sample"]["n_strat_samples"] + cfg["sample"]["n_surf_samples"]))

    def step(self):
        with contextlib.redirect_stdout(io.StringIO()):
            losses, _ = self.tr.step()
        return float(losses["total_loss"].detach()), self.points_per_step


def time_steps(stepper, warmup, steps):
    """-> (seconds of the timed steps, points, per-step seconds)"""
    for _ in range(warmup):
        stepper.step()
    per, pts = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        _, n = stepper.step()
        per.append(time.perf_counter() - t0)
        pts += n
    return sum(per), pts, per
