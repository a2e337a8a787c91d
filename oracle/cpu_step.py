"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's Trainer.step() restated for the host CPU.

Same op sequence as reference isdf/modules/trainer.py:951-1016 on torch CPU tensors: pixel sampling,
gather + compaction, stratified / surface depths, PE + 7 nn.Linear-equivalent GEMMs, autograd input
gradient with create_graph, losses, total.backward() (double back-prop), torch.optim.AdamW.
Used by bench.py for the `cpu_baseline` object and for `--impl reference` (kind "port": the Python
reference itself cannot travel to the GPU box; this port is pinned to it by tests/test_oracle.py)."""
import time

import numpy as np
import torch

from . import isdf_oracle as O


class CpuStepper:
    def __init__(self, H, W, cam, n_frames=5, n_rays=200, cfg=None, seed=1):
        self.cfg = cfg or O.default_cfg()
        self.H, self.W, self.cam = H, W, cam
        self.n_rays, self.n_frames = n_rays, n_frames
        torch.manual_seed(seed)
        np.random.seed(seed)
        v = torch.arange(H, dtype=torch.float32)[:, None]
        u = torch.arange(W, dtype=torch.float32)[None, :]
        self.depth = torch.stack([2.0 + 0.5 * torch.sin(u / 80.0 + 0.1 * k) + 0.3 * torch.cos(v / 60.0)
                                  for k in range(n_frames)])
        self.T = torch.eye(4).repeat(n_frames, 1, 1)
        for k in range(n_frames):
            self.T[k, 0, 3] = 0.05 * k
        nrm = torch.zeros(n_frames, H, W, 3)
        nrm[..., 2] = -1.0
        self.normals = nrm
        E = O.embedding_size(self.cfg["n_freqs"])
        Hd, B = self.cfg["hidden"], self.cfg["block"]
        shapes = [(Hd, E)] + [(Hd, Hd)] * B + [(Hd, Hd + E)] + [(Hd, Hd)] * B + [(1, Hd)]
        self.params = []
        for o, i in shapes:
            w = torch.empty(o, i)
            torch.nn.init.xavier_normal_(w)
            b = (torch.rand(o) * 2 - 1) / (i ** 0.5)
            self.params += [w.requires_grad_(True), b.requires_grad_(True)]
        self.opt = torch.optim.AdamW(self.params, lr=self.cfg["lr"], weight_decay=self.cfg["weight_decay"])

    def step(self):
        cfg = self.cfg
        R = self.n_rays * self.n_frames
        ih = torch.randint(0, self.H, (R,))
        iw = torch.randint(0, self.W, (R,))
        ib = torch.arange(self.n_frames).repeat_interleave(self.n_rays)
        d = self.depth[ib, ih, iw]
        keep = d != 0
        n_keep = int(keep.sum())
        u = torch.rand(n_keep, cfg["n_strat"])
        near = torch.normal(torch.zeros(n_keep, cfg["n_surf"] - 1), 0.1)
        batch = O.sample_rays(self.depth, self.T, self.normals, ib, ih, iw, u, near, self.cam, cfg["min_depth"],
                              cfg["dist_behind_surf"], cfg["n_strat"], cfg["n_surf"])
        layers = [(self.params[2 * i], self.params[2 * i + 1]) for i in range(len(self.params) // 2)]
        pc = batch["pc"].detach().requires_grad_(True)
        noise = torch.randn(pc.shape[:-1])
        sdf = O.sdf_forward(layers, pc, cfg, noise)
        (g,) = torch.autograd.grad(sdf, pc, torch.ones_like(sdf), create_graph=True)
        terms = O.loss_terms(sdf, g, batch, cfg)
        total = terms["total_mat"].mean()
        O.frame_avg(terms["total_mat"].detach(), (self.n_frames, self.H, self.W), batch["indices_b"],
                    batch["indices_h"], batch["indices_w"])
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.opt.step()
        return float(total), n_keep * (cfg["n_strat"] + cfg["n_surf"])


def time_cpu_steps(stepper, warmup, steps):
    for _ in range(warmup):
        stepper.step()
    t0 = time.perf_counter()
    pts = 0
    for _ in range(steps):
        _, n = stepper.step()
        pts += n
    dt = time.perf_counter() - t0
    return dt, pts
