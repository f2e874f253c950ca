"""One flat f32 buffer that holds every parameter gradient of the replica.

Why: the gradient exchange of the data-parallel step (ref slowfast/models/build.py:44-46 — DDP's bucketed
all-reduce) wants large contiguous messages.  Instead of packing 524 separately allocated gradients into a
fresh bucket every step (an extra 753 MB read + 753 MB write, and buffers whose lifetime has to be tracked
across streams), the weight-gradient kernels write straight into their slot of this arena and ``p.grad`` is a
view of it: a bucket is then simply a slice of the arena, reduced in place.

Slots are laid out in REVERSE registration order — the order in which backward produces gradients — so the
parameters of one transformer block are contiguous (one memset zeroes every accumulate-into gradient of the
block) and so are the buckets of the overlapped all-reduce (host/distributed.py::OverlappedGradSync).
"""
import torch


class GradArena:
    ALIGN = 4            # elements: every slot starts on a 16-byte boundary (vectorised f32 atomics / stores)

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.slot = {}
        off = 0
        for p in reversed(self.params):
            assert p.dtype == torch.float32 and p.device == dev
            self.slot[id(p)] = (off, p.numel())
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self._base = self.flat.data_ptr()
        self._key = tuple(p.data_ptr() for p in self.params)

    def matches(self, params):
        """Still the arena of these parameter tensors (same objects, same device)?"""
        ps = [p for p in params if p.requires_grad]
        return len(ps) == len(self.params) and all(a is b for a, b in zip(ps, self.params)) and ps[0].device == self.flat.device

    def has(self, p):
        return id(p) in self.slot

    def view(self, p):
        """A fresh view of p's slot, shaped like p (fresh on purpose: autograd adopts a gradient tensor without a
        copy only if nobody else holds a reference to that tensor object)."""
        lo, n = self.slot[id(p)]
        return self.flat[lo: lo + n].view(p.shape)

    def owns(self, t, p):
        """Does tensor t (e.g. p.grad) already live in p's slot?"""
        return t is not None and t.data_ptr() == self._base + 4 * self.slot[id(p)][0]

    def span(self, params):
        """The contiguous slice covering the slots of `params` (they must be neighbours in registration order)."""
        lo = min(self.slot[id(p)][0] for p in params)
        hi = max(self.slot[id(p)][0] + self.slot[id(p)][1] for p in params)
        return self.flat[lo:hi]

    def range_of(self, params):
        lo = min(self.slot[id(p)][0] for p in params)
        hi = max(self.slot[id(p)][0] + (self.slot[id(p)][1] + self.ALIGN - 1) // self.ALIGN * self.ALIGN for p in params)
        return lo, hi

    def adopt(self, p):
        """Make p.grad live in the arena (a copy only for the few small gradients whose producer is not arena-aware)."""
        if p.grad is None or self.owns(p.grad, p):
            return
        v = self.view(p)
        v.copy_(p.grad)
        p.grad = v


def grad_slot(wc, param):
    """p's slot of the model's arena as a fresh, un-initialised view — or None when there is no arena, the tensor is not
    one of its parameters, or the parameter already holds a gradient (autograd then has to accumulate)."""
    arena = getattr(wc, "arena", None)
    if arena is None or not arena.has(param) or param.grad is not None:
        return None
    return arena.view(param)
