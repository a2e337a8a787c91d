"""Data-parallel plumbing of the hot path: keyframes are sharded over the ranks of one
box, every rank samples rays from its own shard, and ONE all-reduce of the flat gradient per step is
the only collective (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests).

Replicas stay bit-identical because every rank applies the same averaged gradient with the same
AdamW state; rank 0's initial parameters are broadcast once."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def keyframe_owner(frame_index, world_size):
    """Round-robin by insertion order: keyframe k lives on rank k mod world."""
    return int(frame_index) % int(world_size)


def shard_keyframes(frame_indices, rank, world_size):
    return [f for f in frame_indices if keyframe_owner(f, world_size) == rank]


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of a flat buffer (the packed gradient).  No-op for a single rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def average_gradients_(flat, group=None):
    """sum over ranks / world: each rank's gradient is the mean over ITS samples, so the result is the
    mean over all samples when the per-rank valid-sample counts agree (they do for all-valid depth;
    otherwise it is the mean of per-rank means, the usual data-parallel convention)."""
    n = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    allreduce_sum_(flat, group)
    if n > 1:
        flat.mul_(1.0 / n)
    return flat


def broadcast_parameters_(flat, src=0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src, group=group)
    return flat


class GradExchange:
    """C1 fused into K4: the packed gradient lives in SYMMETRIC memory (two buffers, alternating per step) and the
    weight-gradient kernel flushes its tiles with `multimem.red` through the buffers' NVLink-multicast alias, so
    the sum over ranks materialises in every rank's copy while the kernel drains -- no all-reduce kernel, one
    barrier per step (include/isdf_b200.h, isdfb_set_grad_exchange).  Needs NVSwitch multicast (NVLS); the caller
    falls back to the NCCL all-reduce when construction fails on any rank."""

    def __init__(self, engine, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        n = engine.grad_buffer().numel()
        self.n_pad = (n + 1023) // 1024 * 1024                 # keep the second buffer 4 KB aligned
        self.buf = symm_mem.empty(2 * self.n_pad, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        mc = int(self.hdl.multicast_ptr)
        if mc == 0:
            raise RuntimeError("this system exposes no NVLink multicast address for symmetric memory")
        off = int(getattr(self.hdl, "offset", 0) or 0)
        self.buf.zero_()
        base = self.buf.data_ptr()
        stride = 4 * self.n_pad
        engine.set_grad_exchange(base, base + stride, mc + off, mc + off + stride, self.n_pad)
        self.engine = engine
        self.parity = 0
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)                            # nobody adds before everybody cleared

    def barrier(self):
        """All ranks' multimem reductions of this step are complete and visible after this returns (stream order)."""
        self.hdl.barrier(channel=0)

    def close(self):
        try:
            self.engine.set_grad_exchange(0, 0, 0, 0, 0)
        except Exception:
            pass


def try_grad_exchange(engine, device, group=None):
    """Collective: every rank tries to build the exchange; all use it only if all succeeded."""
    ok, ex, err = 1, None, None
    try:
        ex = GradExchange(engine, device, group)
    except Exception as e:     # noqa: BLE001 -- any failure (no NVLS, no symmetric-memory support) means "use NCCL"
        ok, err = 0, e
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return ex, None
    if ex is not None:
        ex.close()
    return None, err
