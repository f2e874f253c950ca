"""KLDiv and EgoNCE (ref slowfast/models/losses.py:51-82, :152-170, :187-207) on the CUDA kernels.

The reference composes ``kldiv+egonce`` in its train loop (tools/train_avgaze_net.py:76-88):

    preds = frame_softmax(preds, temperature=2)
    similarity = sim_matrix(v_embed, a_embed)
    loss = KLDiv()(preds, labels_hm) + LOSS_ALPHA * EgoNCE()(similarity)

The same three calls work here unchanged.  ``frame_softmax`` returns the soft-maxed heat-maps and
keeps a handle to the logits, so ``KLDiv`` can run the fused softmax+KL kernel whose backward is
already known (d loss / d logits) — the eight element-wise passes of the reference collapse into
one launch.
"""
import torch
import torch.nn as nn

from .. import kernels as K

def _kldiv_prob(logits, temperature):
    assert logits.dim() == 5 and logits.shape[1] == 1, "expected logits of shape (B, 1, T, H, W)"
    prob = _FrameSoftmaxFn.apply(logits, temperature)
    prob._csts_source = (logits, temperature)      # lets KLDiv run the fused softmax+KL kernel
    return prob


class _FrameSoftmaxFn(torch.autograd.Function):
    """Stand-alone frame softmax (used when the heat-maps feed something other than KLDiv)."""

    @staticmethod
    def forward(ctx, logits, temperature):
        lg = logits.contiguous().float()
        B, _, T, H, W = lg.shape
        # the fused kernel also yields probabilities; a uniform target keeps it well defined
        tgt = torch.full((B, T, H, W), 1.0 / (H * W), dtype=torch.float32, device=lg.device)
        _, prob, _ = K.kldiv_frame_softmax(lg, tgt, temperature, T, want_grad=False)
        ctx.save_for_backward(prob)
        ctx.temperature = temperature
        return prob

    @staticmethod
    def backward(ctx, dprob):
        (prob,) = ctx.saved_tensors
        # softmax Jacobian; tiny (B*T rows of 4096) and only reached if the heat-maps are consumed
        # by a loss other than KLDiv
        inner = (dprob * prob).sum(dim=(-1, -2), keepdim=True)
        return prob * (dprob - inner) / ctx.temperature, None


class _KLDivFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, temperature):
        lg = logits.contiguous().float()
        T = lg.shape[2]
        loss, _, dlogits = K.kldiv_frame_softmax(lg, target.contiguous().float(), temperature, T, want_grad=True)
        ctx.save_for_backward(dlogits)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return K.scale_f32(dlogits, dloss.reshape(1).contiguous().float()), None, None


class KLDiv(nn.Module):
    """KL divergence between predicted and target heat-maps, normalised by T*log(H*W), mean over the
    batch (losses.py:59-82).  `pred` is the output of frame_softmax."""

    def __init__(self):
        super().__init__()
        self.register_buffer("norm_scalar", torch.tensor(1, dtype=torch.float32))

    def forward(self, pred, target=None):
        if target is None:
            # uniform prior (losses.py:67-71): sum p log p - log(1/HW), i.e. the same divergence against q = 1/HW
            B, _, T, H, W = pred.shape
            target = torch.full((B, T, H, W), 1.0 / (H * W), dtype=torch.float32, device=pred.device)
        src = getattr(pred, "_csts_source", None)
        if src is not None:
            logits, temperature = src
        else:
            # a plain probability map: softmax(log p) == p, so the same kernel applies with T = 1
            logits, temperature = torch.log(pred.float() + 1e-30), 1.0
        return _KLDivFusedFn.apply(logits, target, temperature)


class _EgoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, temperature):
        loss, dsim = K.egonce(sim.contiguous().float(), temperature, want_grad=True)
        ctx.save_for_backward(dsim)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (dsim,) = ctx.saved_tensors
        return K.scale_f32(dsim, dloss.reshape(1).contiguous().float()), None


class EgoNCE(nn.Module):
    """Symmetric InfoNCE over a cosine-similarity matrix (losses.py:152-170).  The reference builds its
    diagonal mask with a hard-coded `.cuda()`; here the diagonal is taken on the input's device."""

    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, x):
        assert x.dim() == 2 and x.shape[0] == x.shape[1], "EgoNCE expects a square similarity matrix"
        return _EgoNCEFn.apply(x, self.temperature)


_LOSSES = {"kldiv": KLDiv, "egonce": EgoNCE}


def get_loss_func(loss_name):
    """losses.py:198-207 — returns the loss *class*."""
    if loss_name not in _LOSSES:
        raise NotImplementedError("Loss {} is not supported".format(loss_name))
    return _LOSSES[loss_name]
