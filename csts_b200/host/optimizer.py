"""AdamW with the gradient-norm clip, the GradScaler unscale and the 16-bit weight-copy refresh fused into
one pass over the parameters (reference: tools/train_avgaze_net.py:101-109 and
slowfast/models/optimizer.py:98-104).

``FusedClipAdamW`` is a ``torch.optim.AdamW`` (same param_groups / state_dict layout: ``step``, ``exp_avg``,
``exp_avg_sq`` per parameter, so checkpoints interchange with the reference's optimizer) whose
``clip_and_step(max_norm, scaler)`` replaces the sequence

    scaler.unscale_(optimizer); clip_grad_norm_(params, max_norm); scaler.step(optimizer); scaler.update()

by two launches of libcsts_b200 (csts_grad_sqnorm, csts_clip_adamw_step).  Plain ``step()`` still works (it is
torch's).  Everything the kernels read lives on the device (learning rate, step count, loss scale), so the
step is CUDA-graph capturable; gradients are left as produced (scaled, un-clipped) — nothing reads them after
the step.
"""
import ctypes as C
import struct

import torch

from .. import _lib


class FusedClipAdamW(torch.optim.AdamW):
    def __init__(self, params, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8, weight_cache=None):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=True, capturable=True)
        assert len(self.param_groups) <= 2, "the CSTS recipe has two parameter groups (decay / no decay)"
        self._wc = weight_cache
        self._table = None
        self._table_key = None
        dev = self.param_groups[0]["params"][0].device
        self._total_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._step = torch.zeros(1, dtype=torch.float32, device=dev)          # shared step count (all parameters step together)
        self._lr = [torch.zeros(1, dtype=torch.float32, device=dev) for _ in range(2)]
        self._keepalive = []          # host/device tables referenced by captured graph nodes

    # ---- state -------------------------------------------------------------------------------------------
    def _init_state(self):
        loaded_step = None
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = self._step                                   # one shared device scalar
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif st["step"] is not self._step:                            # state loaded from a checkpoint
                    if loaded_step is None:
                        loaded_step = float(st["step"])                       # all parameters step together: read it once
                    st["step"] = self._step
        if loaded_step is not None:
            self._step.fill_(loaded_step)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._table_key = None        # the moment buffers were replaced: the pointer table must be rebuilt

    def _build_table(self):
        """Device array of csts_mt_tensor (64 bytes each) + the chunk list.  Rebuilt when any pointer moved."""
        chunk = _lib.load().csts_mt_chunk_elems()
        rows, chunks, key = [], [], []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                assert p.dtype == torch.float32 and p.grad.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                st = self.state[p]
                w16 = self._wc.bound_copy(p) if self._wc is not None else None
                w16_ptr = w16.data_ptr() if w16 is not None else 0
                w16_dt = _lib.dt(w16) if w16 is not None else 0
                wd_bits = struct.unpack("<I", struct.pack("<f", float(group["weight_decay"])))[0]
                ti = len(rows)
                rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), w16_ptr,
                             p.numel(), wd_bits | (gi << 32), w16_dt])
                key.append((p.data_ptr(), p.grad.data_ptr(), w16_ptr, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
                chunks.extend([ti, c] for c in range((p.numel() + chunk - 1) // chunk))
        return rows, chunks, tuple(key)

    def _ensure_table(self):
        rows, chunks, key = self._build_table()
        if key == self._table_key:
            return
        dev = self._total_sq.device
        host_t = torch.tensor(rows, dtype=torch.int64).pin_memory()
        host_c = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).pin_memory()
        self._table = (host_t.to(dev, non_blocking=True), host_c.to(dev, non_blocking=True), len(chunks), host_t, host_c)
        self._table_key = key
        if torch.cuda.is_current_stream_capturing():
            self._keepalive.append(self._table)      # the H2D copies are graph nodes: their buffers must outlive the graph

    # ---- the fused tail of the training step ---------------------------------------------------------------
    @torch.no_grad()
    def clip_and_step(self, max_norm=0.0, scaler=None):
        """unscale + clip_grad_norm_(max_norm) + AdamW + 16-bit weight refresh.  `scaler`: an enabled
        torch.amp.GradScaler whose scale() produced the loss that was back-propagated, or None."""
        self._init_state()
        self._ensure_table()
        tensors, chunks, n_chunks = self._table[:3]
        for gi, group in enumerate(self.param_groups):
            lr = group["lr"]
            if torch.is_tensor(lr):
                self._lr[gi].copy_(lr.reshape(1))
            else:
                self._lr[gi].fill_(float(lr))
        if len(self.param_groups) == 1:
            self._lr[1].copy_(self._lr[0])
        scaling = scaler is not None and scaler.is_enabled()
        scale = scaler._get_scale_async() if scaling else None
        beta1, beta2 = self.param_groups[0]["betas"]
        _lib.call("csts_grad_sqnorm", C.c_void_p(tensors.data_ptr()), C.c_void_p(chunks.data_ptr()), n_chunks, _lib.ptr(self._total_sq))
        _lib.call("csts_clip_adamw_step", C.c_void_p(tensors.data_ptr()), C.c_void_p(chunks.data_ptr()), n_chunks, _lib.ptr(self._total_sq),
                  _lib.ptr(scale), _lib.ptr(self._found_inf), _lib.ptr(self._step), _lib.ptr(self._lr[0]), _lib.ptr(self._lr[1]),
                  float(beta1), float(beta2), float(self.param_groups[0]["eps"]), float(max_norm or 0.0))
        self._step.add_(1.0 - self._found_inf)
        if scaling:
            # scaler.update() without the host round trip: same op GradScaler uses
            torch._amp_update_scale_(scaler._scale, scaler._growth_tracker, self._found_inf, scaler.get_growth_factor(),
                                     scaler.get_backoff_factor(), scaler.get_growth_interval())
        if self._wc is not None:
            self._wc.after_fused_step()

    def grad_norm(self, scaler=None):
        """Total gradient L2 norm of the last clip_and_step (device tensor, unscaled)."""
        n = self._total_sq.sqrt().float()
        if scaler is not None and scaler.is_enabled():
            n = n / scaler._get_scale_async()
        return n
