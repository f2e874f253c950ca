"""frame_softmax and sim_matrix (ref slowfast/utils/utils.py:5-24) on the CUDA kernels."""
import torch

from .. import kernels as K


class _SimMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, eps):
        a, b = a.contiguous().float(), b.contiguous().float()
        sim, na, nb = K.sim_matrix_fwd(a, b, eps)
        ctx.save_for_backward(a, b, sim, na, nb)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        a, b, sim, na, nb = ctx.saved_tensors
        da, db = K.sim_matrix_bwd(a, b, sim, dsim.contiguous(), na, nb)
        return da, db, None


def sim_matrix(a, b, eps=1e-8):
    """Cosine-similarity matrix of two (n, D) embedding sets."""
    return _SimMatrixFn.apply(a, b, eps)


def frame_softmax(logits, temperature):
    """softmax(logits / temperature) over H*W per (b, t) frame; logits (B, 1, T, H, W).
    The result remembers its logits so that KLDiv()(frame_softmax(preds, 2), labels) runs as one
    fused softmax + KL + gradient kernel (see losses.py)."""
    from .losses import _kldiv_prob
    return _kldiv_prob(logits, float(temperature))
