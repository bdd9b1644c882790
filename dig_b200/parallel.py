"""Graph-sharded data parallelism for the 3D-graph path.

Every op of the path is confined to one molecule, so a global batch is split into contiguous
molecule ranges, one per rank (one process per GPU), with NO data-path collective in the forward
pass.  The only exchanges are bookkeeping: the max-over-ranks step time (bench.py) and -- once the
backward kernels exist -- one all-reduce of the flat fp32 gradient buffer per step.
torch.distributed is the plumbing (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous [lo, hi) of rank `rank`; sizes differ by at most one, earlier ranks get the extras."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_molecules(molecules, rank=None, world=None):
    """This rank's contiguous slice of a global list of molecules."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(molecules), rank, world)
    return molecules[lo:hi]


def shard_molecules_equal(molecules, rank=None, world=None):
    """This rank's contiguous slice with the SAME length on every rank (the trailing len % world molecules are
    dropped): data-parallel training needs identical step counts per epoch, or the ranks issue different numbers
    of gradient all-reduces and the collectives pair up gradients of different steps (then hang)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    per = len(molecules) // world
    return molecules[rank * per:(rank + 1) * per]


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (bench timing rule: the slowest rank defines the step)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_per_molecule(values, device=None):
    """Concatenate per-molecule results (e.g. energies [n_local, C]) of all ranks in rank order."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return values
    world = dist.get_world_size()
    n_local = torch.tensor([values.size(0)], dtype=torch.long, device=values.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    m = int(max(int(s) for s in sizes))
    pad = torch.zeros(m, *values.shape[1:], dtype=values.dtype, device=values.device)
    pad[:values.size(0)] = values
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:int(s)] for o, s in zip(out, sizes)], dim=0)


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_gradients(parameters):
    """Data-parallel gradient exchange (BASELINE configs[4]: 'NCCL grad all-reduce'): every rank ran backward on its
    own molecule shard; the gradients are averaged with ONE all-reduce of a flat fp32 buffer per step (the model has
    ~1.9 M parameters = 7.6 MB, far below the size where bucketing/overlap would pay on NVLink 5).

    NCCL averages in the collective (ReduceOp.AVG); gloo (CPU tests only) has no AVG, so SUM + scale there.
    Parameters without a gradient on some rank would desynchronise the flat layout, so they are zero-filled."""
    params = [p for p in parameters if p.requires_grad]
    world = world_size()
    if world == 1 or not params:
        return 0
    grads = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)              # one multi-tensor copy instead of ~160 small ones
    return flat.numel() * flat.element_size()
