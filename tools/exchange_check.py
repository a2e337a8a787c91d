"""N-GPU check of the fused gradient exchange (C1 fused into K4) against the NCCL all-reduce.
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/exchange_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import __graft_entry__ as ge
if rank == 0:
    ge.build()
dist.barrier()
from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P
from isdf_b200 import parallel
import bench

def say(*a):
    if rank == 0:
        print(*a, flush=True)

# ---- engine level: multimem flush == NCCL sum ---------------------------------------------------------------
cfg = O.default_cfg(noise_std=0.05)
sd = C.golden_weights(17, gain=1.5)
batch, noise = C.loss_batch(100 + rank, 300)            # every rank its own rays
eng = P.make_engine(dev, cfg, "bf16x3", max_points=32768)
eng.pack_weights(P.flat_params(sd, dev))
b = {k: v.to(dev) for k, v in batch.items()}
lc = P.loss_cfg_from(cfg, 300 * 27)
nz = noise.to(dev)
def k4():
    eng.train_fwd_bwd(b["pc"], b["z_vals"], b["depth_sample"], b["dirs_C_sample"], b["T_WC_sample"], b["norm_sample"], nz, lc)
eng.zero_grad(); k4()
ref = eng.grad_buffer().clone()
dist.all_reduce(ref)
ex, err = parallel.try_grad_exchange(eng, dev)
say("exchange:", "multicast OK" if ex else "UNAVAILABLE %r" % (err,))
if ex is not None:
    for par in (0, 1, 0):
        eng.select_grad_buffer(par)
        k4()
        eng.zero_grad_buffer(1 - par)
        ex.barrier()
        got = eng.grad_buffer().clone()
        e = float((got - ref).abs().max() / ref.abs().max())
        say("parity %d: max |fused - nccl| / max|nccl| = %.3g" % (par, e))
        assert e < 1e-5, e
        torch.cuda.synchronize(); dist.barrier()
    ex.close()
del eng

# ---- trainer level: same trajectory (eager: identical RNG consumption), then timing with graphs -------------------
from isdf.modules import trainer as trainer_mod
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
for _ in range(10):
    h.barrier(channel=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    h.barrier(channel=0)
torch.cuda.synchronize()
say("symmetric-memory barrier: %.1f us each (200 back to back)" % (1e6 * (time.perf_counter() - t0) / 200))

def run(mode, graph, steps):
    np.random.seed(1 + rank); torch.manual_seed(1 + rank)
    wl = bench.WORKLOADS["default"]
    c = bench.make_config(wl, "bf16x3", "fast")
    c["b200"]["grad_exchange"] = mode
    c["b200"]["cuda_graph"] = graph
    tr = trainer_mod.Trainer(dev, c, incremental=True)
    dist.broadcast(tr.sdf_map.flat_parameters(), 0)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(wl["keyframes"]):
            tr.last_is_keyframe = True
            tr.add_data(tr.get_data([rank + i * world]))
    for _ in range(10):
        tr.step(sync=False)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(sync=False)
    torch.cuda.synchronize(); dist.barrier()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    flat = tr.sdf_map.flat_parameters().clone()
    chk = flat.clone(); dist.broadcast(chk, 0)
    kind = "multicast" if tr._xchg is not None else "nccl"
    say("mode %-9s graph=%d -> exchange=%s nccl_in_graph=%s: %.3f ms/step; replicas differ by %.3g" %
        (mode, graph, kind, tr._nccl_in_graph, ms, float((flat - chk).abs().max())))
    if tr._xchg is not None:
        tr._xchg.close()
    return flat

if os.environ.get("QUICK"):            # N > 2 smoke: fused exchange only, graph mode
    run("multicast", 1, 100)
    dist.destroy_process_group()
    sys.exit(0)
a = run("nccl", 0, 40)
b = run("multicast", 0, 40)
say("eager, same RNG: params after 50 steps |multicast - nccl| max = %.3g (of max %.3g)" %
    (float((a - b).abs().max()), float(a.abs().max())))
run("nccl", 1, 300)
run("multicast", 1, 300)
dist.destroy_process_group()
