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


# ------------------------------------------------------------------------------------------------------------
# Flat parameter / gradient storage, overlapped gradient all-reduce, fused Adam
class FlatParameters:
    """All trainable parameters of a model as views of ONE flat fp32 buffer, their gradients as views of a second
    one (allocated once): the gradient exchange is an all-reduce of slices of `grad` with no gather / copy-back,
    zero_grad is one fill, and the optimizer is one kernel over the flat buffers.

    Slices start at multiples of 4 elements (16 B) so that every view is vector-load aligned."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("FlatParameters: the model has no trainable parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params) or dt != torch.float32:
            raise ValueError("FlatParameters: parameters must be fp32 tensors on one device")
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.data = torch.zeros(total, dtype=dt, device=dev)
        self.grad = torch.zeros(total, dtype=dt, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.data[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off:off + p.numel()].view_as(p)
        invalidate = getattr(model, "invalidate_packed", None)
        if callable(invalidate):
            invalidate()

    def zero_grad(self):
        """One fill; re-attaches the views if something (optimizer.zero_grad(set_to_none=True)) dropped them."""
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view_as(p)

    def adopt_stray_gradients(self):
        """Autograd accumulates in place into an existing .grad, so gradients normally land in the flat buffer; a
        gradient that was replaced instead (create_graph=True accumulates out of place) is copied back in."""
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            if g is not None and g.data_ptr() != self.grad.data_ptr() + 4 * off:
                view = self.grad[off:off + p.numel()].view_as(p)
                view.copy_(g.detach())
                p.grad = view


class GradientReducer:
    """Data-parallel gradient averaging over the flat gradient buffer, launched DURING backward.

    The buffer is cut into `n_buckets` contiguous ranges (parameter order); a post-accumulate hook on every
    parameter counts arrivals, and the moment the last gradient of a bucket has been written the bucket's
    all-reduce is enqueued asynchronously (NCCL runs it on its own stream, behind an event of the compute stream),
    i.e. under the tail of the backward pass.  `finish()` reduces whatever did not fire (parameters without a
    gradient this step) and makes the compute stream wait for the collectives.  With world_size 1 it does nothing."""

    def __init__(self, flat, n_buckets=2):
        self.flat = flat
        self.world = world_size()
        self.bytes_per_step = flat.numel * 4
        self.buckets = []          # (start, stop, [param indices])
        if self.world == 1:
            return
        target = (flat.numel + n_buckets - 1) // n_buckets
        start, members = 0, []
        for idx, (p, off) in enumerate(zip(flat.params, flat.offsets)):
            members.append(idx)
            stop = off + (p.numel() + 3) // 4 * 4
            if stop - start >= target or idx == len(flat.params) - 1:
                self.buckets.append((start, stop, members))
                start, members = stop, []
        self.bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for idx in mem:
                self.bucket_of[idx] = b
        self.pending = [0] * len(self.buckets)
        self.works = [None] * len(self.buckets)
        self.armed = False
        for idx, p in enumerate(flat.params):
            p.register_post_accumulate_grad_hook(self._make_hook(idx))

    def _make_hook(self, idx):
        def hook(_param):
            if not self.armed:
                return
            b = self.bucket_of[idx]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        start, stop, _ = self.buckets[b]
        view = self.flat.grad[start:stop]
        if dist.get_backend() == "nccl":
            self.works[b] = dist.all_reduce(view, op=dist.ReduceOp.AVG, async_op=True)
        else:
            self.works[b] = (dist.all_reduce(view, op=dist.ReduceOp.SUM, async_op=True), view)

    def arm(self):
        """Call once per step before backward."""
        if self.world == 1:
            return
        self.pending = [len(mem) for _, _, mem in self.buckets]
        self.works = [None] * len(self.buckets)
        self.armed = True

    def finish(self):
        """Call after backward: every bucket reduced, compute stream ordered behind the collectives."""
        if self.world == 1:
            return
        self.armed = False
        self.flat.adopt_stray_gradients()
        for b in range(len(self.buckets)):
            if self.works[b] is None:
                self._launch(b)
        for b, w in enumerate(self.works):
            if isinstance(w, tuple):            # gloo: SUM, then scale
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self.works = [None] * len(self.buckets)


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (amsgrad off) as ONE kernel over FlatParameters (dig3d_adam_step).

    An Optimizer subclass so that lr schedulers (run.py:51 StepLR) and state_dict / load_state_dict keep working:
    the per-parameter state entries ('step', 'exp_avg', 'exp_avg_sq') are views of the flat moment buffers, i.e. the
    checkpoint has the keys of the reference's Adam checkpoint (run.py:92)."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not isinstance(flat, FlatParameters):
            raise TypeError("FlatAdam takes a FlatParameters object")
        self.flat = flat
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.steps = 0
        for p, off in zip(flat.params, flat.offsets):
            self.state[p] = {"step": torch.tensor(0.0),
                             "exp_avg": self.exp_avg[off:off + p.numel()].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + p.numel()].view_as(p)}

    def zero_grad(self, set_to_none=True):
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        from . import ops
        loss = closure() if closure is not None else None
        self.flat.adopt_stray_gradients()
        grp = self.param_groups[0]
        self.steps += 1
        ops.adam_step(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, float(grp["lr"]),
                      float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]),
                      self.steps)
        for st in self.state.values():
            st["step"].fill_(float(self.steps))
        ops.invalidate_packed()            # the weights changed behind tensor._version
        ops.repack_h16_all()               # ... and the training linears get fresh packed copies in a few launches
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # re-home the loaded moments in the flat buffers
        steps = 0
        for p, off in zip(self.flat.params, self.flat.offsets):
            st = self.state[p]
            for key, buf in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                view = buf[off:off + p.numel()].view_as(p)
                if st[key].data_ptr() != view.data_ptr():
                    view.copy_(st[key])
                    st[key] = view
            steps = max(steps, int(float(st["step"])))
        self.steps = steps


class PacedGC:
    """Training-loop garbage-collection pacing for data-parallel runs.

    A step allocates a few thousand short-lived Python objects (tensor wrappers, autograd nodes, ctypes arguments), so
    CPython's generational collector fires at arbitrary steps, at different steps on different ranks; with a gradient
    all-reduce every step, each rank's pause becomes everybody's pause (the step time is the slowest rank's).  Inside
    this context automatic collection is off, the long-lived objects (model, optimizer state, CUDA library handles) are
    frozen out of the young generations, and `tick()` collects every `every` steps -- the SAME steps on every rank.
    Reference counting still frees tensors immediately; only cycle collection is paced."""

    def __init__(self, every=64):
        self.every, self.count, self._was_enabled = int(every), 0, True

    def __enter__(self):
        import gc
        self._was_enabled = gc.isenabled()
        gc.collect()
        gc.freeze()
        gc.disable()
        self.count = 0
        return self

    def tick(self):
        self.count += 1
        if self.every > 0 and self.count % self.every == 0:
            import gc
            gc.collect()

    def __exit__(self, *exc):
        import gc
        gc.unfreeze()
        if self._was_enabled:
            gc.enable()
        return False
