"""Where does the end-to-end (host frames in, loss out) step time go?  Dev tool, not part of the product.
   python tools/e2e_profile.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
import bench
import __graft_entry__ as ge

ge.build()
from isdf.modules import trainer as trainer_mod

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["default"]
cfg = bench.make_config(wl, "bf16x3", "fast")
np.random.seed(1); torch.manual_seed(1)
tr = trainer_mod.Trainer(dev, cfg, incremental=True)
for i in range(wl["keyframes"]):
    tr.last_is_keyframe = True
    tr.add_data(tr.get_data([i]))
for _ in range(30):
    tr.step(sync=False)
torch.cuda.synchronize()
tr.scene_dataset.cache_frames = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for i in range(0, n, 10):
    _ = tr.scene_dataset[100 + i // 10]
T = dict(get=0.0, add=0.0, step=0.0, read=0.0, ev=0.0)
# variants of the bare step loop (no ingest)
for name, fn in (("replay only, one sync at the end", lambda: tr.step(sync=False)),
                 ("step(sync=False) + cuda.synchronize()", lambda: (tr.step(sync=False), torch.cuda.synchronize())),
                 ("step() [events + sync]", lambda: tr.step())):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    print("%-45s %.3f ms/step" % (name, 1e3 * (time.perf_counter() - t0) / 200))
nf = 100
t_all = time.perf_counter()
for i in range(n):
    if i % 10 == 0:
        t0 = time.perf_counter(); fd = tr.get_data([nf]); nf += 1
        t1 = time.perf_counter(); tr.last_is_keyframe = False; tr.add_data(fd)
        t2 = time.perf_counter(); T["get"] += t1 - t0; T["add"] += t2 - t1
    t0 = time.perf_counter(); losses, ev_ms = tr.step(); T["ev"] += ev_ms / 1e3
    t1 = time.perf_counter(); v = float(losses["total_loss"])
    t2 = time.perf_counter(); T["step"] += t1 - t0; T["read"] += t2 - t1
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
print("e2e %.3f ms/step; per-step ms:" % (1e3 * tot / n), {k: round(1e3 * v / n, 4) for k, v in T.items()})
print("per-ingest ms: get %.3f add %.3f" % (1e3 * T["get"] / (n / 10), 1e3 * T["add"] / (n / 10)))
# finer: inside get_data
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(5):
    fd = tr.get_data([nf]); tr.last_is_keyframe = False; tr.add_data(fd)
    for _ in range(10):
        losses, _ = tr.step(); float(losses["total_loss"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

# ---- inside get_data (fast mode), line by line ------------------------------------------------------------
import numpy as np
from isdf_b200.datasets.data_util import FrameData
acc = {}
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
for k in range(10):
    tr.scene_dataset[200 + k]
for k in range(10):
    for _ in range(3):
        tr.step()
    t = time.perf_counter()
    s = tr.scene_dataset[200 + k]; t = tick("dataset[]", t)
    im_np, depth_np, T_np = s["image"][None, ...], s["depth"][None, ...], s["T"][None, ...]
    tr._stage_evt.synchronize(); t = tick("stage_evt.sync", t)
    np.copyto(tr._stage[0].numpy(), depth_np, casting="same_kind"); np.copyto(tr._stage[1].numpy(), T_np, casting="same_kind"); t = tick("copyto pinned", t)
    depth = tr._stage[0].to(tr.device, non_blocking=True); T = tr._stage[1].to(tr.device, non_blocking=True); tr._stage_evt.record(); t = tick("h2d", t)
    data = FrameData(frame_id=np.array([200 + k]), im_batch=None, im_batch_np=im_np, depth_batch=depth, depth_batch_np=depth_np, T_WC_batch=T, T_WC_batch_np=T_np); t = tick("FrameData()", t)
    data.normal_batch = tr.sdf_map.engine().ingest_normals(depth[0], tr.cam)[None, :]; t = tick("normals", t)
    out = FrameData(); out.add_frame_data(data, replace=False); t = tick("out.add", t)
    tr.last_is_keyframe = False; tr.add_data(out); t = tick("tr.add_data", t)
print("get_data pieces, ms per ingest (each followed by a device sync):", {k: round(1e2 * v, 3) for k, v in acc.items()})
