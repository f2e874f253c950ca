"""Data-parallel helpers of the hot path (ref slowfast/utils/distributed.py:15-110, :305-320).

The only cross-sample coupling of the training step is the InfoNCE similarity matrix: every rank
needs every rank's (B_local, 256) video and audio embeddings.  `all_gather_with_grad` performs ONE
all-gather of the concatenated [v | a] rows (the reference issues one per tensor) and is
differentiable: backward returns the local rank's slice of the incoming gradient.

Deliberate fix: the reference stores ``ctx.rank = 0`` (distributed.py:23) so every
rank back-propagates rank 0's slice.  Here the true rank is used, and the slice is multiplied by
the world size so that — after DDP's mean all-reduce of parameter gradients — the result equals
the gradient of the single-process global-batch loss (each rank evaluates the full-batch NCE loss
but can only differentiate through its own rows).
"""
import os

import torch
import torch.distributed as dist


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def is_master_proc(num_gpus=8):
    return get_rank() % num_gpus == 0 if get_world_size() > 1 else True


class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor):
        world = get_world_size()
        ctx.rows = tensor.shape[0]
        ctx.rank = get_rank()
        ctx.world = world
        out = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(out, tensor.contiguous())
        return out

    @staticmethod
    def backward(ctx, grad_output):
        lo = ctx.rows * ctx.rank
        return grad_output[lo: lo + ctx.rows] * float(ctx.world)


def all_gather_with_grad(tensors):
    """Differentiable all-gather (concatenation along dim 0) of a list of equally-shaped 2-D tensors."""
    if get_world_size() == 1:
        return list(tensors)
    widths = [t.shape[1] for t in tensors]
    packed = _AllGatherRows.apply(torch.cat(list(tensors), dim=1))
    return list(torch.split(packed, widths, dim=1))


def all_reduce(tensors, average=True):
    """In-place sum (or mean) across ranks; returns the list (distributed.py:75-91)."""
    world = get_world_size()
    if world == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, async_op=False)
    if average:
        for t in tensors:
            t.mul_(1.0 / world)
    return tensors


def all_gather(tensors):
    """Concatenate each tensor across ranks along dim 0 (distributed.py:52-72)."""
    world = get_world_size()
    if world == 1:
        return list(tensors)
    out = []
    for t in tensors:
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous())
        out.append(buf)
    return out


def allreduce_gradients(params, bucket_bytes=256 << 20):
    """Average the gradients of `params` across ranks with a few large NCCL all-reduces over NVLink /
    NVSwitch (NVLS when available): gradients are packed into contiguous f32 buckets (reverse
    registration order ~ the order backward produces them), reduced with ReduceOp.AVG, and `p.grad`
    is re-pointed at its slice of the bucket (no copy back).  Unlike the DDP reducer this is plain
    stream-ordered work, so the whole training step including the exchange can live in one CUDA graph."""
    world = get_world_size()
    if world == 1:
        return
    todo = [p for p in params if p.grad is not None][::-1]
    i = 0
    while i < len(todo):
        bucket, size = [], 0
        while i < len(todo) and (not bucket or size + todo[i].numel() * 4 <= bucket_bytes):
            bucket.append(todo[i])
            size += todo[i].numel() * 4
            i += 1
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        off = 0
        for p in bucket:
            n = p.numel()
            p.grad = flat[off: off + n].view_as(p)
            off += n


class OverlappedGradSync:
    """Gradient averaging overlapped with backward, as plain stream-ordered work (graph-capturable).

    Every parameter gradient lives in the replica's GradArena (host/grad_arena.py), laid out in reverse
    registration order — the order backward produces them — so a bucket is a contiguous slice of the arena and is
    all-reduced IN PLACE: no packing copy, no re-pointing of ``p.grad``, and no buffer whose lifetime spans two
    streams (the arena is persistent).  A post-accumulate-grad hook counts down each bucket; when it is complete
    its slice is all-reduced (NCCL, AVG) on a side stream while backward keeps running on the main stream.
    ``finish()`` joins the side stream before the clip / optimizer step.  Buckets are ~32 MB, so only the last small
    bucket of early-encoder gradients is exposed after backward ends.

    Factored exchange.  Three parameters — the (768, 768, 1, 8, 8) frame-pool kernels — hold 453 MB of the 753 MB
    gradient, yet each of those gradients is a product of two thin factors: dW = dY^T . X with dY (B*T, 768) and
    X (B*T, 49152), B*T = 32 rows per rank.  Averaging dW over ranks equals ONE product over the concatenated rows,
    mean_r dY_r^T X_r = (1/R) [dY_1; ..; dY_R]^T [X_1; ..; X_R], so for these parameters (``model.factored_grad_params()``)
    the ranks all-gather the 16-bit factors (3.2 MB per rank and kernel) and every rank forms the averaged gradient
    itself, straight into its arena slot (``factored_wgrad``): 60 % of the all-reduce volume never crosses NVLink, and
    the sum over ranks is accumulated in f32 inside one tensor-core product instead of being rounded per rank.
    """

    def __init__(self, model, bucket_bytes=32 << 20, tail_bytes=40 << 20, tail_bucket_bytes=8 << 20):
        self.model = model
        if hasattr(model, "_wc"):
            model._wc.sync_aware = True           # this exchange waits for the model's audio and weight-gradient streams
        self.bucket_bytes, self.tail_bytes, self.tail_bucket_bytes = bucket_bytes, tail_bytes, tail_bucket_bytes
        self.params = [p for p in model.parameters() if p.requires_grad][::-1]
        self.arena = None
        self.groups, self.ranges, self.group_of, self.pending = [], [], {}, []
        self.stream = None
        self.enabled = False
        factored = model.factored_grad_params() if hasattr(model, "factored_grad_params") and \
            os.environ.get("CSTS_FACTORED_WGRAD", "1") == "1" else []
        self.factored = {id(p) for p in factored}
        self._factor_bufs = {}
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _bind(self, arena):
        self.arena = arena
        self.groups, cur, size = [], [], 0
        left = sum(p.numel() * 4 for p in self.params if id(p) not in self.factored)
        for p in self.params:
            if id(p) in self.factored:           # its own group, never all-reduced (factored_wgrad fills the slot)
                if cur:
                    self.groups.append(cur)
                self.groups.append([p])
                cur, size = [], 0
                continue
            cur.append(p)
            size += p.numel() * 4
            left -= p.numel() * 4
            # the last buckets (the first encoder blocks, whose gradients arrive when backward ends) are small: what is
            # exposed after the last kernel of backward is one short all-reduce
            if size >= (self.bucket_bytes if left > self.tail_bytes else self.tail_bucket_bytes):
                self.groups.append(cur)
                cur, size = [], 0
        if cur:
            self.groups.append(cur)
        self.external = [len(g) == 1 and id(g[0]) in self.factored for g in self.groups]
        self.ranges = [arena.range_of(g) for g in self.groups]
        self.group_of = {id(p): g for g, ps in enumerate(self.groups) for p in ps}

    def start(self):
        """Call right before loss.backward() (after the forward pass, which creates the arena)."""
        if get_world_size() == 1:
            return
        arena = getattr(getattr(self.model, "_wc", None), "arena", None)
        assert arena is not None, "OverlappedGradSync needs the model's gradient arena (CSTS_GRAD_ARENA=1, training forward first)"
        if arena is not self.arena:
            self._bind(arena)
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        self.pending = [len(g) for g in self.groups]
        self.enabled = True
        self._touched = False            # whether this backward has put work on the exchange stream
        self.model._wc.grad_sync = self

    def _hook(self, param):
        if not self.enabled:
            return
        self.arena.adopt(param)          # no-op for gradients written in place; a small copy for the others
        g = self.group_of[id(param)]
        self.pending[g] -= 1
        if self.pending[g] == 0 and not self.external[g]:
            lo, hi = self.ranges[g]
            self.stream.wait_stream(torch.cuda.current_stream())
            wc = getattr(self.model, "_wc", None)
            if wc is not None:
                # a bucket mixes gradients of the video encoder (this stream), of the audio encoder (its own stream) and
                # weight gradients produced on the backward pass's second stream.  (Waiting per bucket for exactly the
                # points at which its gradients were produced — one event per stream — measured no faster: 21.98 vs
                # 21.93 ms/step on 2 GPUs.)
                for s in wc.branch_streams():
                    self.stream.wait_stream(s)
            self._touched = True
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.arena.flat[lo:hi], op=dist.ReduceOp.AVG)

    def finish(self):
        """Call after loss.backward(): joins the exchange before gradients are consumed."""
        if not self.enabled:
            return
        self.enabled = False
        if hasattr(self.model, "_wc"):
            self.model._wc.grad_sync = None
        assert all(c == 0 for c in self.pending), "a parameter received no gradient"
        if self._touched:
            torch.cuda.current_stream().wait_stream(self.stream)

    def factored_wgrad(self, weight, dy16, x16, out):
        """Rank-averaged weight gradient of a factored parameter: all-gather the rows of [dY | X] (16-bit, this rank's
        (rows, O) and (rows, Kd) factors), then out (O, Kd) f32 = (1/R) [dY_1;..;dY_R]^T [X_1;..;X_R] on the exchange
        stream.  `out` is the parameter's arena slot.  The staging buffers are persistent (one pair per parameter), so no
        allocation's lifetime spans two streams; they are free again once ``finish()`` has joined the exchange stream."""
        from .. import kernels as K
        world = get_world_size()
        rows, O = dy16.shape
        Kd = x16.shape[1]
        bufs = self._factor_bufs.get(id(weight))
        if bufs is None or bufs[0].dtype != dy16.dtype or bufs[0].shape != (rows, O + Kd) or bufs[1].shape[0] != world * rows:
            bufs = (torch.empty((rows, O + Kd), dtype=dy16.dtype, device=dy16.device),
                    torch.empty((world * rows, O + Kd), dtype=dy16.dtype, device=dy16.device))
            self._factor_bufs[id(weight)] = bufs
        packed, gathered = bufs
        packed[:, :O].copy_(dy16)
        packed[:, O:].copy_(x16)
        self.stream.wait_stream(torch.cuda.current_stream())
        self._touched = True
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(gathered, packed)
            K.gemm(gathered, gathered, b_off=O, M=O, N=Kd, K=world * rows, a_kmajor=False, b_kmajor=False, lda=O + Kd, ldb=O + Kd,
                   out=out, out_dtype=torch.float32, alpha=1.0 / world)
        return out


def broadcast_parameters(model, src=0):
    """Make every rank start from rank `src`'s parameters and buffers (what DDP's constructor does)."""
    if get_world_size() == 1:
        return
    tensors = [p.data for p in model.parameters()] + [b.data for b in model.buffers()]
    for t in tensors:
        dist.broadcast(t, src=src)


def init_distributed_training(cfg):
    """distributed.py:305-320 creates one process group per machine for SyncBN; CSTS has no
    BatchNorm, so nothing is needed beyond the default group."""
    return None
