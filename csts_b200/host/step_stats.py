"""Per-iteration bookkeeping of the training loop without per-iteration host synchronisation.

The reference loop (tools/train_avgaze_net.py:95,112-128) reads the loss on the host every iteration
(``check_nan_losses``, three ``.item()`` calls), issues three scalar all-reduces and two heat-map all-gathers, and
evaluates ``adaptive_f1`` through ~100 small launches — all inside the timed loop.  ``AsyncStepStats`` keeps the same
quantities on the device: every iteration adds the three loss terms to a device accumulator (one small launch); every
``period`` iterations the fused metric kernel runs on the current batch, ONE all-reduce averages the packed record across
ranks, and the record is copied to pinned host memory asynchronously.  ``poll()`` hands it out once the copy has landed,
so the loop never waits for the GPU.  A NaN loss shows up in the record (``nan`` flag) one period late instead of stalling
every step.
"""
import torch
import torch.distributed as dist

from . import distributed as du
from .metrics import adaptive_f1_async, thresholds_for


class AsyncStepStats:
    FIELDS = ("loss", "kldiv_loss", "egonce_loss", "f1", "recall", "precision")

    def __init__(self, dataset, period=10, device=None):
        self.dataset, self.period = dataset, max(1, int(period))
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._acc = torch.zeros(3, dtype=torch.float32, device=self.device)
        self._n = 0
        self._it = 0
        self._pending = []          # (iteration, pinned host record, event, steps averaged, threshold grid)
        self.last = None

    def update(self, loss, kldiv_loss=None, egonce_loss=None, preds=None, labels_hm=None, labels=None):
        """Call once per iteration with the device loss tensors (and, for the metric, the frame-softmaxed predictions and
        labels of the iteration).  Never synchronises."""
        zero = loss.detach().new_zeros(())
        terms = torch.stack([loss.detach().float().reshape(()), (kldiv_loss if kldiv_loss is not None else zero).detach().float().reshape(()),
                             (egonce_loss if egonce_loss is not None else zero).detach().float().reshape(())])
        self._acc += terms
        self._n += 1
        self._it += 1
        if self._it % self.period:
            return
        rec = torch.zeros(8, dtype=torch.float32, device=self.device)
        rec[:3] = self._acc / float(self._n)
        if preds is not None:
            m = adaptive_f1_async(preds, labels_hm, labels, self.dataset, rescale=True)
            rec[3:6] = m[:3]
            rec[6] = m[4]                                      # threshold index (rank 0's is reported)
        rec[7] = torch.isnan(self._acc).any().float()
        if du.get_world_size() > 1:
            avg = rec[:6].clone()
            dist.all_reduce(avg)                               # one collective for the whole record
            rec[:6] = avg / du.get_world_size()
        host = torch.empty(8, dtype=torch.float32).pin_memory()
        host.copy_(rec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((self._it, host, ev, self._n))
        self._acc.zero_()
        self._n = 0

    def poll(self, wait=False):
        """The most recent record whose copy has landed (dict) or None.  wait=True blocks for the newest one."""
        out = None
        while self._pending and (wait or self._pending[0][2].query()):
            it, host, ev, n = self._pending.pop(0)
            if wait:
                ev.synchronize()
            vals = host.tolist()
            out = dict(zip(self.FIELDS, vals[:6]))
            out.update(iteration=it, steps=n, threshold=float(thresholds_for(self.dataset)[int(vals[6])]), nan=bool(vals[7]))
        if out is not None:
            self.last = out
        return out
