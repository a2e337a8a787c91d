"""Per-step device time of the first graph replays after a synchronize (dev tool): where does a 20-step timed region
lose time against a 200-step one?   python tools/step_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import __graft_entry__ as ge
ge.build()
from isdf.modules import trainer as T
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["default"]
np.random.seed(1); torch.manual_seed(1)
tr = T.Trainer(dev, bench.make_config(wl, "bf16x3g", "fast"), incremental=True)
for i in range(wl["keyframes"]):
    tr.last_is_keyframe = True; tr.add_data(tr.get_data([i]))
for _ in range(40):
    tr.step(sync=False)
torch.cuda.synchronize()
for trial in range(3):
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    cpu = []
    for i in range(n):
        tr.step(sync=False)
        ev[i + 1].record()
        cpu.append(time.perf_counter())
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print("trial", trial, "first 8 steps ms:", " ".join("%.3f" % m for m in ms[:8]), "| mean 20: %.4f  mean last 20: %.4f" % (sum(ms[:20]) / 20, sum(ms[20:]) / 20),
          "| cpu per launch us: %.1f" % (1e6 * (cpu[-1] - t0) / n))
    time.sleep(0.2)
