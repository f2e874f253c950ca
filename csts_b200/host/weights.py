"""Precision mode and the 16-bit operand copies of the fp32 master parameters.

Parameters stay fp32 ``nn.Parameter``s under the reference's names (checkpoints, optimizer and DDP
see exactly the reference's state).  The tensor-core kernels consume 16-bit operands, so each GEMM
weight gets a cached 16-bit copy.  The same (N, K) copy serves forward (K-major B operand of
Y = X . W^T) and backward (MN-major B operand of dX = dY . W) — the tcgen05 kernel takes either layout
through its shared-memory descriptors, so no transposed copy is kept.  A copy is refreshed whenever
the parameter's version counter or storage changes, i.e. once per optimizer step.

Precision modes (one 16-bit type per mode: a tcgen05 kind::f16 MMA cannot mix f16 and bf16 operands):

* default — bf16 activations and bf16 activation gradients.  fp32 exponent range, no loss scaling;
  7 mantissa bits put the whole-model gradient ~3e-2 (relative L2) from the fp32 reference.
* ``TRAIN.MIXED_PRECISION: True`` — the reference's own mixed-precision contract (fp16 autocast +
  ``GradScaler``, tools/train_avgaze_net.py:70,99-109): fp16 activations and fp16 activation gradients.
  10 mantissa bits bring the gradient within ~1e-2 of fp32; the caller scales the loss
  (``scaler.scale(loss).backward()``) exactly as the reference loop does.
"""
import os
from collections import namedtuple

import torch

from .. import kernels as K

Precision = namedtuple("Precision", ["act", "grad"])
BF16 = Precision(torch.bfloat16, torch.bfloat16)
FP16 = Precision(torch.float16, torch.float16)


def precision_of(cfg):
    return FP16 if cfg.TRAIN.MIXED_PRECISION else BF16


class WeightCache:
    """Per-model context handed to every autograd node of the path: the precision mode, the 16-bit operand copies
    of the weights, the gradient arena (host/grad_arena.py) and the second stream of the backward pass."""

    def __init__(self, precision=BF16):
        self._store = {}
        self.act, self.grad = precision
        self._gen = 0            # training-step generation: copies made in an earlier step are not trusted
        self._fresh = False      # set by the fused optimizer step, which rewrites the copies itself
        self._dirty = False      # a training step ran since the copies were last rebuilt for a non-training forward
        self.arena = None        # GradArena, created by the model at its first training forward
        self._fork = None
        self._audio = None
        self._main = None
        self.sync_aware = False         # set by OverlappedGradSync: the gradient exchange waits for every stream of the model
        self.grad_sync = None           # the active OverlappedGradSync between its start() and finish() (factored weight gradients)
        self._handoff = {}
        self.act_checkpoint = False     # MODEL.ACT_CHECKPOINT: encoder blocks keep their input only and recompute in backward
        self.defer_join = False
        self.parallel_audio = os.environ.get("CSTS_PARALLEL_AUDIO", "1") == "1"
        self.fork_backward = os.environ.get("CSTS_FORK_WGRAD", "1") == "1"
        self.audio_wgrad_inline = os.environ.get("CSTS_AUDIO_WGRAD_INLINE", "1") == "1"

    def backward_fork(self):
        """The weight-gradient branch of the backward pass (block.py::_Fork): one second stream per model.  With
        `defer_join` (set by train_step for the duration of a backward it controls) the branch is joined once, by
        join_backward(), instead of at the end of every block, so it keeps running under the next blocks' dX chain."""
        from .block import _Fork
        if not self.fork_backward:
            return _Fork()
        # The audio encoder's backward already runs on its own stream, next to the four times longer video chain: its
        # weight gradients stay on that stream.  On the shared branch they would queue behind every video weight gradient
        # (autograd enqueues the whole video backward first) and finish last, after both dX chains.
        if self.audio_wgrad_inline and self._audio is not None and torch.cuda.current_stream() == self._audio:
            return _Fork()
        if self._fork is None or self._fork.side.device.index != torch.cuda.current_device():
            self._fork = _Fork(torch.cuda.Stream())
        return self._fork

    def audio_stream(self):
        """Second stream of the forward pass (the audio encoder, csts.py), or None when disabled (CSTS_PARALLEL_AUDIO=0)."""
        if not self.parallel_audio:
            return None
        # Under data parallelism a gradient reducer must know about the second stream (OverlappedGradSync does and says so);
        # DistributedDataParallel's reducer orders its buckets after the hook's stream only, so under DDP everything stays
        # on one stream.
        if not self.sync_aware and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            return None
        if self._audio is None or self._audio.device.index != torch.cuda.current_device():
            self._audio = torch.cuda.Stream()
        return self._audio

    def note_main_stream(self):
        """Called at the top of the forward pass: the stream the video encoder, fusion and decoder (and their backward) run on."""
        self._main = torch.cuda.current_stream()

    def branch_streams(self):
        """Every stream on which gradients of this model may still be in flight: the forward's own stream, the audio
        encoder's and the weight-gradient branch's.  A gradient bucket mixes tensors produced on all three, and the hook
        that completes it runs on only one of them."""
        out = [s for s in (self._main, self._audio) if s is not None]
        if self._fork is not None and self._fork.side is not None:
            out.append(self._fork.side)
        return out

    def join_backward(self, waiter=None):
        if self._fork is not None:
            self._fork.join(waiter)

    # ---- gradient hand-over between consecutive blocks (block.py::block_backward) ---------------------------------
    def give_handoff(self, dx, dx16):
        """dx: the f32 gradient a block returns for its input; dx16: its 16-bit DropPath-scaled copy.  Keeping dx in the entry
        keeps its address unique until the producer block picks the pair up."""
        self._handoff[dx.data_ptr()] = (dx, dx16)

    def take_handoff(self, dy):
        hit = self._handoff.pop(dy.data_ptr(), None)
        if hit is None or hit[0].shape.numel() != dy.numel() or hit[1].shape != dy.shape:
            return None
        return hit[1]

    def begin_training_step(self):
        self._handoff.clear()
        """Called at the top of every training forward.  A parameter's version counter is not a reliable
        "changed" signal (torch's *fused* optimizers update parameters without bumping it), so in training the
        copies are rebuilt once per step — unless the fused clip+AdamW step (host/optimizer.py) has just
        rewritten them, which it announces through after_fused_step()."""
        self._dirty = True
        if self._fresh:
            self._fresh = False
        else:
            self._gen += 1

    def begin_inference(self):
        """Called at the top of every non-training forward: after a training step the parameters have moved behind
        the version counter's back, so the copies are rebuilt once before they are used for evaluation."""
        if self._dirty:
            self._dirty = False
            self._fresh = False
            self._gen += 1

    def invalidate(self):
        """The parameters changed outside the cache's knowledge (a CUDA-graph replay of a step that does not refresh the
        copies itself, a checkpoint load): rebuild every copy at its next use."""
        self._gen += 1
        self._fresh = False

    def _get(self, param, kind, make):
        key = (id(param), kind)
        ver = (param.data_ptr(), param._version, self._gen)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            # an existing copy is refreshed IN PLACE: its address is baked into captured CUDA graphs and into the
            # fused optimizer's pointer table, both of which keep rewriting / reading it
            val = make(param.detach(), hit[1] if hit is not None and hit[0][0] == ver[0] else None)
        self._store[key] = (ver, val)
        return val

    def w(self, param):
        """Linear weight (N, K) f32 -> 16-bit (N, K)."""
        return self._get(param, "w", lambda p, out: K.cast16(p.reshape(p.shape[0], -1), self.act, out=out))

    def w_padded(self, param, kp):
        """Conv weight (N, ...) f32 -> 16-bit (N, kp), zero padded columns (patch embed)."""
        return self._get(param, ("pad", kp), lambda p, out: K.cast16(p.reshape(p.shape[0], -1), self.act, ld_out=kp, out=out))

    # ---- fused optimizer step (host/optimizer.py): the AdamW kernel writes the 16-bit copy itself -----------
    def bound_copy(self, param):
        """The live plain (N, K) copy of `param` (same element order as the parameter), or None."""
        hit = self._store.get((id(param), "w"))
        return hit[1] if hit is not None else None

    def after_fused_step(self):
        """The fused step updated the parameters in place behind autograd's back (no version bump): the plain
        copies it refreshed stay valid, the re-laid-out ones (padded / permuted) must be rebuilt."""
        for key in [k for k in self._store if k[1] != "w"]:
            del self._store[key]
        self._fresh = True

    def clear(self):
        self._store.clear()
