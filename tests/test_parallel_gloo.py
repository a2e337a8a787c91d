"""CPU test of the N>1 path (world_size 2, gloo): keyframe sharding + ONE all-reduce of the flat
gradient reproduces the single-process gradient.  The per-rank compute is done by the oracle here
(no GPU in this container); the plumbing under test is isdf_b200/parallel.py, the code Trainer uses."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import isdf_oracle as O
from tests.golden import common as C
from isdf_b200 import parallel


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = O.default_cfg(noise_std=0.0, hidden=256)
    sd = C.golden_weights(5)
    layers = O.layers_from_state_dict(sd, 2)
    R = 16
    batch, _ = C.loss_batch(77, R)                       # the global batch; rank r owns rays r::world
    sel = torch.arange(rank, R, world)
    shard = {k: v[sel] for k, v in batch.items()}
    out = O.step_autograd(layers, shard, cfg, None)       # mean over the rank's own samples
    flat = torch.cat([g.reshape(-1) for g in out["grads"]])
    # parameters: replicas must start identical
    p = torch.cat([t.reshape(-1) for t in sd.values()]).clone()
    if rank != 0:
        p += 1.0
    parallel.broadcast_parameters_(p, 0)
    parallel.average_gradients_(flat)
    w, r = parallel.world()
    torch.save(dict(flat=flat, p=p, world=w, rank=r), os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r), weights_only=False) for r in range(world)]
    assert torch.equal(res[0]["flat"], res[1]["flat"])           # every replica applies the same gradient
    assert torch.equal(res[0]["p"], res[1]["p"])                 # ... to the same parameters
    assert res[0]["world"] == 2 and {res[0]["rank"], res[1]["rank"]} == {0, 1}
    cfg = O.default_cfg(noise_std=0.0, hidden=256)
    sd = C.golden_weights(5)
    batch, _ = C.loss_batch(77, 16)
    full = O.step_autograd(O.layers_from_state_dict(sd, 2), batch, cfg, None)
    ref = torch.cat([g.reshape(-1) for g in full["grads"]])
    err = float((res[0]["flat"] - ref).norm() / ref.norm())
    assert err < 1e-5, err                                       # equal shard sizes: mean of means == global mean


class _FakeEngine:
    """Just enough of Engine for GradExchange.__init__ to reach the symmetric-memory allocation."""

    def grad_buffer(self):
        return torch.zeros(1000)

    def set_grad_exchange(self, *a):
        raise AssertionError("must not be reached on a CPU device")


def _fallback_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # no NVLink multicast here: construction fails on every rank, the collective vote says "use the all-reduce"
    ex, err = parallel.try_grad_exchange(_FakeEngine(), torch.device("cpu"))
    g = torch.full((8,), float(rank + 1))
    parallel.allreduce_sum_(g)
    torch.save(dict(ex=ex is None, err=repr(err), g=g), os.path.join(out_dir, "f%d.pt" % rank))
    dist.destroy_process_group()


def test_exchange_setup_falls_back_collectively(tmp_path):
    """try_grad_exchange is a COLLECTIVE decision: when the multicast exchange cannot be built, every rank learns it
    (no rank waits in a rendezvous) and the NCCL / gloo all-reduce path carries on."""
    world = 2
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_fallback_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "f%d.pt" % r), weights_only=False) for r in range(world)]
    assert all(r["ex"] for r in res) and all(r["err"] != "None" for r in res)
    assert torch.equal(res[0]["g"], torch.full((8,), 3.0)) and torch.equal(res[1]["g"], res[0]["g"])


def test_keyframe_sharding_is_a_partition():
    frames = list(range(23))
    parts = [parallel.shard_keyframes(frames, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == frames
    assert all(parallel.keyframe_owner(f, 4) == r for r in range(4) for f in parts[r])
    assert parallel.world() == (1, 0)
    t = torch.ones(4)
    assert torch.equal(parallel.average_gradients_(t.clone()), t)   # single process: no-op
