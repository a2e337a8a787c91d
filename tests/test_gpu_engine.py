"""GPU parity tests: the CUDA path, called through the C ABI, against the golden vectors of the
reference and against the fp64 oracle on seeded inputs.  Tolerances are stated per precision mode."""
import os

import pytest
import torch

from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = torch.device("cuda:0")

# (sdf, g, loss scalars, weight-gradient relative Frobenius) tolerances, max-abs/max-abs-ref
# gw = weight-gradient relative Frobenius error at the 27 000-sample batch; gw_small = on the ~1000-sample
# golden cases, where a single sample flipping a kink of the loss (|.|, max, relu, sign(|g|-1)) moves the
# gradient by ~1/N and the tolerance has to absorb a few such flips.
TOL = {
    "fp32": dict(sdf=5e-6, g=1e-4, loss=5e-5, gw=1e-3, gw_small=1e-3),
    "bf16x3": dict(sdf=1e-4, g=1e-3, loss=1e-3, gw=5e-3, gw_small=2e-2),   # north-star: sdf within 1e-4 rel
    # product default: sdf / g / losses are those of bf16x3 (same products); the weight-gradient operands are single
    # bf16 (2^-9 relative rounding per operand, unbiased): the measured rel. Frobenius error of dW is ~1.5e-3 at
    # 27 000 samples -- below the step-to-step sampling noise of the gradient itself (new rays every step)
    "bf16x3g": dict(sdf=1e-4, g=1e-3, loss=1e-3, gw=5e-3, gw_small=2e-2),
    # fast mode: one bf16 pass.  With Softplus(beta=100) a 2^-9 relative error on a pre-activation of O(1)
    # is comparable to the 0.01-wide transition of the activation, so sigma -- hence d sdf/d x and the
    # gradients -- are only statistically close.  Stated separately; not the parity mode.
    "bf16": dict(sdf=5e-2, g=0.5, loss=0.15, gw=0.5, gw_small=0.6),
}
MODES = [m for m in os.environ.get("ISDFB_TEST_MODES", "fp32,bf16x3,bf16x3g,bf16").split(",") if m]


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _engine(cfg, mode, **kw):
    try:
        return P.make_engine(DEV, cfg, mode, **kw)
    except Exception as e:   # a mode that is not built must fail loudly, not silently pass
        pytest.fail("engine creation failed for %s: %s" % (mode, e))


# ---------------------------------------------------------------------------------- K1
def test_sampling_matches_reference_golden():
    from isdf_b200.engine import make_camera
    gold = load("sample.pt")
    F, H, W = 3, 32, 48
    depth = torch.stack([C.synthetic_depth(k, H, W, invalid_frac=0.15) for k in range(F)]).to(DEV)
    T = torch.stack([C.synthetic_pose(k) for k in range(F)]).to(DEV)
    nrm = torch.stack([C.synthetic_normals(H, W, 0.1, 40 + k) for k in range(F)]).to(DEV)
    cam = make_camera(40.0, 42.0, 23.5, 15.5, H, W)
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    ib, ih, iw = gold["ib"].to(DEV), gold["ih"].to(DEV), gold["iw"].to(DEV)
    d, n, valid = eng.gather_rays(depth, nrm, ib, ih, iw, cam)
    keep = valid.bool()
    assert torch.equal(ib[keep].cpu(), gold["ib2"])
    assert torch.equal(d[keep].cpu(), gold["depth"])
    assert torch.equal(n[keep].cpu(), gold["norm"])
    lin = torch.linspace(0, 1, 20).to(DEV)
    pc, z, dirs_C, T_s = eng.sample_rays(T, ib[keep], ih[keep], iw[keep], d[keep], gold["u"].to(DEV),
                                         gold["n_near"].to(DEV), lin, 19, 8, cam, 0.07, 0.1)
    assert torch.equal(z.cpu(), gold["z"])                       # bit-exact depths
    assert torch.equal(dirs_C.cpu(), gold["dirs_C"])
    assert torch.equal(T_s.cpu(), gold["T"])
    assert torch.allclose(pc.cpu(), gold["pc"], atol=1e-6, rtol=0)


def test_sampling_frame_map_and_empty():
    from isdf_b200.engine import make_camera
    F, H, W = 4, 16, 24
    depth = torch.stack([C.synthetic_depth(k, H, W) for k in range(F)]).to(DEV)
    cam = make_camera(20.0, 20.0, 11.5, 7.5, H, W)
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    g = C.gen(3)
    ib = torch.randint(0, 2, (64,), generator=g).to(DEV)
    ih = torch.randint(0, H, (64,), generator=g).to(DEV)
    iw = torch.randint(0, W, (64,), generator=g).to(DEV)
    fmap = torch.tensor([3, 1], device=DEV)
    d, _, valid = eng.gather_rays(depth, None, ib, ih, iw, cam, frame_map=fmap)
    assert torch.equal(d, depth[fmap[ib], ih, iw]) and bool(valid.all())
    e = torch.empty(0, dtype=torch.int64, device=DEV)
    d0, _, v0 = eng.gather_rays(depth, None, e, e, e, cam)
    assert d0.numel() == 0 and v0.numel() == 0


# ---------------------------------------------------------------------------------- K2/K3
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag,seed,gain,tr", [("g1", 21, 1.0, None), ("g2_rigid", 22, 2.0, 6)])
def test_forward_and_input_gradient_golden(mode, tag, seed, gain, tr):
    gold = load("sdfmap.pt")[tag]
    sd = C.golden_weights(seed, gain=gain)
    cfg = O.default_cfg(transform=C.rigid_transform(tr) if tr else None)
    eng = _engine(cfg, mode, max_points=1024)
    eng.pack_weights(P.flat_params(sd, DEV))
    x = ((torch.rand(96, 3, generator=C.gen(12)) - 0.5) * torch.tensor([12.0, 4.0, 12.0])).to(DEV)
    sdf = eng.forward(x)
    sdf2, g = eng.forward(x, want_grad=True)
    t = TOL[mode]
    assert P.rel(sdf.cpu(), gold["sdf"]) < t["sdf"]
    assert P.rel(sdf2.cpu(), gold["sdf"]) < t["sdf"]
    assert P.rel(g.cpu(), gold["grad"]) < t["g"]


@pytest.mark.parametrize("mode", MODES)
def test_forward_ragged_sizes_and_noise(mode):
    sd = C.golden_weights(5)
    cfg = O.default_cfg()
    eng = _engine(cfg, mode, max_points=256)        # forces internal chunking + ragged tail
    eng.pack_weights(P.flat_params(sd, DEV))
    layers = [(w.double(), b.double()) for w, b in O.layers_from_state_dict(sd, 2)]
    for n in (1, 127, 129, 700):
        x = (torch.rand(n, 3, generator=C.gen(n)) - 0.5) * 6
        nz = torch.randn(n, generator=C.gen(n + 1))
        ref = O.sdf_forward(layers, x.double(), dict(cfg, noise_std=0.3), nz.double())
        out = eng.forward(x.to(DEV), noise=nz.to(DEV), noise_std=0.3)
        assert out.shape == (n,)
        assert P.rel(out.cpu(), ref) < TOL[mode]["sdf"]
    assert eng.forward(torch.empty(0, 3, device=DEV)).numel() == 0


def test_fused_sampler_matches_the_separate_kernels_in_distribution():
    """K1 fused (fast mode): same geometry as gather_rays + sample_rays on the pixels it drew, the reference's
    distributions for every random quantity, fresh numbers on every call and on every CUDA-graph replay."""
    from isdf_b200.engine import make_camera
    F, H, W, n_rays, n_strat, n_surf = 4, 96, 128, 512, 19, 8
    S = n_strat + n_surf
    depth = torch.stack([C.synthetic_depth(k, H, W, invalid_frac=0.1) for k in range(F + 2)]).to(DEV)
    nrm = torch.stack([C.synthetic_normals(H, W, 0.05, 70 + k) for k in range(F + 2)]).to(DEV)
    T = torch.stack([C.synthetic_pose(k) for k in range(F + 2)]).to(DEV)
    fmap = torch.tensor([5, 0, 3, 2], device=DEV)
    cam = make_camera(100.0, 101.0, 63.5, 47.5, H, W)
    lin = torch.linspace(0, 1, n_strat + 1).to(DEV)
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    a = eng.sample_fused(depth, nrm, T, fmap, F, n_rays, n_strat, n_surf, cam, 0.07, 0.1, lin, seed=1234)
    R = F * n_rays
    ib, ih, iw = a["indices_b"], a["indices_h"], a["indices_w"]
    assert torch.equal(ib, torch.arange(F, device=DEV).repeat_interleave(n_rays))
    assert int(ih.min()) >= 0 and int(ih.max()) < H and int(iw.min()) >= 0 and int(iw.max()) < W
    assert abs(float(ih.float().mean()) - (H - 1) / 2) < 0.05 * H and abs(float(iw.float().mean()) - (W - 1) / 2) < 0.05 * W
    # gather + geometry: the separate K1 kernels on the same pixels / numbers give the same values
    d2, n2, v2 = eng.gather_rays(depth, nrm, ib, ih, iw, cam, frame_map=fmap)
    assert torch.equal(d2, a["depth_sample"]) and torch.equal(v2, a["ray_valid"])
    assert torch.equal(torch.nan_to_num(n2, nan=7.0), torch.nan_to_num(a["norm_sample"], nan=7.0))
    z, d = a["z_vals"], a["depth_sample"]
    far = d + 0.1
    u = ((z[:, n_surf:] - 0.07) / (far - 0.07)[:, None] * n_strat - torch.arange(n_strat, device=DEV)[None, :])
    ok = a["ray_valid"].bool()
    assert torch.equal(z[:, 0], d)
    assert float(u[ok].min()) > -1e-4 and float(u[ok].max()) < 1 + 1e-4 and abs(float(u[ok].mean()) - 0.5) < 0.02
    near = z[:, 1:n_surf]
    assert bool((near[ok] >= 0.07).all()) and bool((near[ok] <= far[ok][:, None] + 1e-6).all())
    inner = (near > 0.07 + 1e-6) & (near < far[:, None] - 1e-6) & ok[:, None]
    off = (near - d[:, None])[inner]                      # un-clamped offsets ~ N(0, 0.1^2) truncated at +0.1
    assert abs(float(off[off < 0].std()) / 0.1 - 0.6028) < 0.05          # std of a half-normal = 0.6028 sigma
    nz = a["noise"]
    assert abs(float(nz.mean())) < 0.02 and abs(float(nz.std()) - 1.0) < 0.02
    dirs = a["dirs_C_sample"]
    assert torch.allclose(dirs[:, 0], (iw.float() - 63.5) / 100.0) and torch.allclose(dirs[:, 1], (ih.float() - 47.5) / 101.0)
    Ts = T[fmap[ib]]
    assert torch.equal(a["T_WC_sample"], Ts)
    pc_ref = Ts[:, None, :3, 3] + (Ts[:, :3, :3] @ dirs[:, :, None])[:, None, :, 0] * z[:, :, None]
    assert torch.allclose(a["pc"], pc_ref, atol=1e-5)
    assert abs(float(a["inv_count_dev"]) - 1.0 / (int(ok.sum()) * S)) < 1e-9
    # fresh numbers per call, and per graph replay (the step counter lives on the device)
    b = eng.sample_fused(depth, nrm, T, fmap, F, n_rays, n_strat, n_surf, cam, 0.07, 0.1, lin, seed=1234)
    assert not torch.equal(a["indices_h"], b["indices_h"]) and not torch.equal(a["noise"], b["noise"])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        c = eng.sample_fused(depth, nrm, T, fmap, F, n_rays, n_strat, n_surf, cam, 0.07, 0.1, lin, seed=1234)
    g.replay(); torch.cuda.synchronize(); h1 = c["indices_h"].clone()
    g.replay(); torch.cuda.synchronize()
    assert not torch.equal(h1, c["indices_h"])


# ---------------------------------------------------------------------------------- K4
CASES = [("c1", 31, 1.0, None, 48, 0.25, "L1"),
         ("c2_rigid_gain2", 32, 2.0, 9, 40, 0.04, "L1"),
         ("c3_L2", 33, 1.5, None, 24, 0.0, "L2")]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_train_step_matches_reference_golden(mode, case):
    tag, seed, gain, tr, R, nstd, lt = case
    gold = load("step.pt")[tag]
    sd = C.golden_weights(seed, gain=gain)
    cfg = O.default_cfg(noise_std=nstd, loss_type=lt, transform=C.rigid_transform(tr) if tr else None)
    batch, noise = C.loss_batch(seed + 100, R)
    eng = _engine(cfg, mode, max_points=1024)       # 1296 / 1080 / 648 points: >1 chunk for c1, c2
    out = P.run_train(eng, sd, batch, noise, cfg, DEV)
    t = TOL[mode]
    assert P.rel(out["sdf"], gold["sdf"]) < t["sdf"]
    assert P.rel(out["g"], gold["grad"]) < t["g"]
    assert P.rel(out["loss_mat"], gold["total_mat"]) < max(t["loss"], 10 * t["g"] * 0.02)
    n = out["sdf"].numel()
    for k, idx in (("sdf_loss", 0), ("grad_loss", 1), ("eikonal_loss", 2), ("total_loss", 3)):
        ref = gold["losses"][k]
        assert abs(float(out["sums"][idx]) / n - ref) <= t["loss"] * max(abs(ref), 1e-3), k
    for name, gr in zip(sd.keys(), out["grads"]):
        sub = C.subsample(gr) if gr.numel() > 4096 else gr
        assert P.rel_fro(sub, gold["grad_sub"][name]) < t["gw_small"], name


@pytest.mark.parametrize("mode", MODES)
def test_train_step_vs_fp64_oracle_default_size(mode):
    """27 000 samples (1000 rays x 27): BASELINE configs[1] batch shape."""
    cfg = O.default_cfg(noise_std=0.08)
    sd = C.golden_weights(91, gain=1.3)
    batch, noise = C.loss_batch(92, 1000)
    eng = _engine(cfg, mode, max_points=8192)
    out = P.run_train(eng, sd, batch, noise, cfg, DEV)
    ref = P.oracle_train(sd, batch, noise, cfg)
    e = P.compare_train(out, ref)
    t = TOL[mode]
    assert e["sdf"] < t["sdf"] and e["g"] < t["g"], e
    assert e["total_loss"] < t["loss"] and e["sdf_loss"] < t["loss"], e
    assert e["grad_max_rel_fro"] < t["gw"], e


@pytest.mark.parametrize("mode", MODES)
def test_train_properties_full_size(mode):
    """Size-independent properties: chunk invariance, additivity over ray subsets, masked rays,
    mean(loss_mat) == loss_sums / N."""
    cfg = O.default_cfg(noise_std=0.05)
    sd = C.golden_weights(93)
    R = 1000
    batch, noise = C.loss_batch(94, R)
    t = TOL[mode]
    a = P.run_train(_engine(cfg, mode, max_points=32768), sd, batch, noise, cfg, DEV)
    b = P.run_train(_engine(cfg, mode, max_points=4096), sd, batch, noise, cfg, DEV)
    # tensor-core modes: a tile's K order is rotated per CTA (tc_chain.cu rot_kstep), so a point that lands on
    # another CTA after re-chunking sees a different fp32 summation order -> equal up to rounding, not bit-wise
    assert P.rel(a["sdf"], b["sdf"]) < {"fp32": 1e-6, "bf16x3": 5e-5, "bf16x3g": 5e-5, "bf16": 2e-2}[mode]
    assert max(P.rel_fro(x, y) for x, y in zip(a["grads"], b["grads"])) < max(1e-4, 0.1 * t["gw"])
    assert abs(float(a["loss_mat"].double().mean()) - float(a["sums"][3]) / (R * 27)) < 1e-5
    # additivity: grads(first half) + grads(second half) == grads(all) at fixed inv_count
    eng = _engine(cfg, mode, max_points=32768)
    eng.pack_weights(P.flat_params(sd, DEV))
    eng.zero_grad()
    lc = P.loss_cfg_from(cfg, R * 27)
    bd = {k: v.to(DEV) for k, v in batch.items()}
    nz = noise.to(DEV)
    for sl in (slice(0, 400), slice(400, R)):
        eng.train_fwd_bwd(bd["pc"][sl], bd["z_vals"][sl], bd["depth_sample"][sl], bd["dirs_C_sample"][sl],
                          bd["T_WC_sample"][sl], bd["norm_sample"][sl], nz[sl], lc)
    both = P.unflatten(eng.export_grads().cpu(), sd)
    assert max(P.rel_fro(x, y) for x, y in zip(both, a["grads"])) < max(1e-4, 0.1 * t["gw"])
    # masked rays contribute nothing
    eng.zero_grad()
    valid = torch.ones(R, dtype=torch.uint8, device=DEV)
    valid[400:] = 0
    _, _, lm, sums = eng.train_fwd_bwd(bd["pc"], bd["z_vals"], bd["depth_sample"], bd["dirs_C_sample"],
                                       bd["T_WC_sample"], bd["norm_sample"], nz, lc, ray_valid=valid)
    assert float(lm[400:].abs().max()) == 0.0
    eng2 = _engine(cfg, mode, max_points=32768)
    eng2.pack_weights(P.flat_params(sd, DEV))
    eng2.zero_grad()
    sl = slice(0, 400)
    eng2.train_fwd_bwd(bd["pc"][sl], bd["z_vals"][sl], bd["depth_sample"][sl], bd["dirs_C_sample"][sl],
                       bd["T_WC_sample"][sl], bd["norm_sample"][sl], nz[sl], lc)
    g1 = eng.export_grads().cpu()
    g2 = eng2.export_grads().cpu()
    assert P.rel_fro(g1, g2) < max(1e-4, 0.1 * t["gw"])


# ---------------------------------------------------------------------------------- N2 ('pc' bound)
PC_CASES = [("p1", 61, 1.0, None, 48, 0.25, "L1"), ("p2_rigid_L2", 62, 1.5, 9, 32, 0.0, "L2")]


@pytest.mark.parametrize("case", PC_CASES, ids=[c[0] for c in PC_CASES])
def test_bounds_pc_matches_reference_golden(case):
    """isdfb_bounds_pc against loss.bounds_pc of the unmodified reference (loss.py:56-89), incl. the NaN row."""
    tag, seed, gain, tr, R, nstd, lt = case
    gold = load("step_pc.pt")[tag]
    batch, _ = C.loss_batch_pc(seed + 100, R)
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    b = {k: v.to(DEV) for k, v in batch.items()}
    bnd, vec = eng.bounds_pc(b["pc"], b["z_vals"], b["depth_sample"])
    bnd, vec = bnd.cpu(), vec.cpu()[:, 1:]
    assert torch.allclose(bnd, gold["bounds"], atol=1e-6, rtol=1e-6)
    nan_gold = gold["grad_vec"][..., 0].isnan()
    assert torch.equal(vec[..., 0].isnan(), nan_gold)
    assert torch.allclose(vec[~nan_gold], gold["grad_vec"][~nan_gold], atol=2e-6)


def test_bounds_pc_full_size_properties():
    """C4-sized batch (20 480 rays x 64 samples vs 20 480 surface points): every bound is attained by a
    surface point (brute-force check on a slice), |bound| of a surface sample is 0, masked rays are ignored."""
    R, S = 20480, 64
    batch, _ = C.loss_batch(301, R, S=S, n_surf=8)
    b = {k: v.to(DEV) for k, v in batch.items()}
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    valid = torch.ones(R, dtype=torch.uint8, device=DEV)
    valid[::7] = 0
    bnd, vec = eng.bounds_pc(b["pc"], b["z_vals"], b["depth_sample"], ray_valid=valid)
    keep = valid.bool()
    assert float(bnd[keep][:, 0].abs().max()) == 0.0            # a surface sample is its own closest point
    assert float(bnd[~keep].abs().max()) == 0.0
    surf = b["pc"][keep][:, 0]
    rows = torch.arange(0, R, 97, device=DEV)
    rows = rows[keep[rows]]
    d = (b["pc"][rows][:, :, None, :] - surf[None, None]).norm(dim=-1).min(dim=-1).values
    assert torch.allclose(bnd[rows].abs(), d, atol=1e-5, rtol=1e-5)
    n = vec[keep][:, 1:].norm(dim=-1)
    assert float((n[~n.isnan()] - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", PC_CASES, ids=[c[0] for c in PC_CASES])
def test_train_step_pc_bound_matches_reference_golden(mode, case):
    tag, seed, gain, tr, R, nstd, lt = case
    gold = load("step_pc.pt")[tag]
    sd = C.golden_weights(seed, gain=gain)
    cfg = O.default_cfg(noise_std=nstd, loss_type=lt, transform=C.rigid_transform(tr) if tr else None,
                        bounds_method="pc")
    batch, noise = C.loss_batch_pc(seed + 100, R)
    out = P.run_train(_engine(cfg, mode, max_points=1024), sd, batch, noise, cfg, DEV)
    t = TOL[mode]
    assert P.rel(out["sdf"], gold["sdf"]) < t["sdf"]
    assert P.rel(out["g"], gold["grad"]) < t["g"]
    assert P.rel(out["loss_mat"], gold["total_mat"]) < max(t["loss"], 10 * t["g"] * 0.02)
    n = out["sdf"].numel()
    for k, idx in (("sdf_loss", 0), ("grad_loss", 1), ("eikonal_loss", 2), ("total_loss", 3)):
        ref = gold["losses"][k]
        assert abs(float(out["sums"][idx]) / n - ref) <= t["loss"] * max(abs(ref), 1e-3), k
    for name, gr in zip(sd.keys(), out["grads"]):
        sub = C.subsample(gr) if gr.numel() > 4096 else gr
        assert P.rel_fro(sub, gold["grad_sub"][name]) < t["gw_small"], name


# ---------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_frame_bins_matches_reference_golden(case):
    gold = load("step.pt")[case[0]]
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    la, fa = eng.frame_bins(gold["total_mat"].to(DEV), gold["frame_ib"].to(DEV), gold["frame_ih"].to(DEV),
                            gold["frame_iw"].to(DEV), 4, 16, 24, 8)
    assert torch.allclose(la.cpu(), gold["loss_approx"], atol=1e-5)
    assert torch.allclose(fa.cpu(), gold["frame_avg"], atol=1e-6)


# ---------------------------------------------------------------------------------- K6
def test_adamw_matches_torch_and_repacks():
    cfg = O.default_cfg()
    sd = C.golden_weights(61)
    eng = _engine(cfg, "fp32", max_points=1024)
    p = P.flat_params(sd, DEV)
    p_ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([p_ref], lr=0.0013, weight_decay=0.012)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    batch, noise = C.loss_batch(62, 32)
    bd = {k: v_.to(DEV) for k, v_ in batch.items()}
    lc = P.loss_cfg_from(dict(cfg, noise_std=0.0), 32 * 27)
    eng.pack_weights(p)
    for it in range(3):
        eng.zero_grad()
        eng.train_fwd_bwd(bd["pc"], bd["z_vals"], bd["depth_sample"], bd["dirs_C_sample"], bd["T_WC_sample"],
                          bd["norm_sample"], None, lc)
        p_ref.grad = eng.export_grads().clone()
        opt.step()
        eng.adamw(p, m, v, it + 1, 0.0013, weight_decay=0.012)
        assert torch.allclose(p, p_ref.detach(), atol=1e-7, rtol=1e-6)
    # the packed weights follow the update: forward with the engine == oracle with updated params
    x = ((torch.rand(64, 3, generator=C.gen(9)) - 0.5) * 6).to(DEV)
    new_sd = {k: t for k, t in zip(sd.keys(), P.unflatten(p.cpu(), sd))}
    layers = [(w.double(), b.double()) for w, b in O.layers_from_state_dict(new_sd, 2)]
    ref = O.sdf_forward(layers, x.cpu().double(), cfg)
    assert P.rel(eng.forward(x).cpu(), ref) < 5e-6


def test_errors_are_loud():
    from isdf_b200 import _lib
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    with pytest.raises(_lib.IsdfbError):
        eng.forward(torch.zeros(4, 3, device=DEV))         # weights not packed
    with pytest.raises(ValueError):
        eng.pack_weights(torch.zeros(10, device=DEV))
    with pytest.raises(RuntimeError):
        P.make_engine(torch.device("cpu"), O.default_cfg())
    # 'pc' bound: the two precomputed arrays go together (host check and C-ABI check)
    with pytest.raises(ValueError):
        P.loss_cfg_from(O.default_cfg(), 10, bounds=torch.zeros(2, 5, device=DEV))
    # gradient exchange: refused for the CUDA-core mode, and buffer selection needs an installed exchange
    buf = torch.zeros(4, device=DEV)
    with pytest.raises(_lib.IsdfbError):
        eng.set_grad_exchange(buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4)
    with pytest.raises(_lib.IsdfbError):
        eng.select_grad_buffer(0)
    tc = _engine(O.default_cfg(), "bf16x3", max_points=1024)
    with pytest.raises(_lib.IsdfbError):                         # buffers smaller than the packed gradient
        tc.set_grad_exchange(buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4)
    with pytest.raises(_lib.IsdfbError):
        tc.zero_grad_buffer(1)


# ---------------------------------------------------------------------------------- N3
def test_ingest_normals_matches_torch_restatement():
    from isdf_b200.engine import make_camera
    from isdf_b200.geometry import transform
    H, W = 120, 160
    depth = C.synthetic_depth(3, H, W, invalid_frac=0.05).to(DEV)
    depth = 2.0 + 0.5 * torch.sin(torch.arange(W, device=DEV)[None, :] / 20.0) + 0.3 * torch.cos(torch.arange(H, device=DEV)[:, None] / 15.0)
    depth[5:9, 7:30] = 0.0
    cam = make_camera(100.0, 100.0, 79.5, 59.5, H, W)
    eng = _engine(O.default_cfg(), "fp32", max_points=1024)
    n = eng.ingest_normals(depth, cam)
    ref = transform.estimate_pointcloud_normals(transform.pointcloud_from_depth_torch(depth, 100.0, 100.0, 79.5, 59.5))
    nan_a, nan_b = torch.isnan(n[..., 0]), torch.isnan(ref[..., 0])
    assert torch.equal(nan_a, nan_b)
    ok = ~nan_a
    close = (n[ok] - ref[ok]).abs().max(dim=-1).values < 1e-4
    assert float(close.float().mean()) > 0.999       # near-ties of the neighbour-pair cost may pick another valid pair
