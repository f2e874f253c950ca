"""bf16 operand copies of the fp32 master parameters.

Parameters stay fp32 ``nn.Parameter``s under the reference's names (checkpoints, optimizer and DDP
see exactly the reference's state).  The tensor-core kernels consume bf16 operands, so each GEMM
weight gets a cached bf16 copy.  The same (N, K) copy serves forward (K-major B operand of Y = X . W^T)
and backward (MN-major B operand of dX = dY . W) — the tcgen05 kernel takes either layout through its
shared-memory descriptors, so no transposed copy is kept.  A copy is refreshed whenever the parameter's
version counter or storage changes, i.e. once per optimizer step.
"""
import torch

from .. import kernels as K


class WeightCache:
    def __init__(self):
        self._store = {}

    def _get(self, param, kind, make):
        key = (id(param), kind)
        ver = (param.data_ptr(), param._version)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            val = make(param.detach())
        self._store[key] = (ver, val)
        return val

    def w(self, param):
        """Linear weight (N, K) f32 -> bf16 (N, K)."""
        return self._get(param, "w", lambda p: K.cast_bf16(p.reshape(p.shape[0], -1)))

    def w_padded(self, param, kp):
        """Conv weight (N, ...) f32 -> bf16 (N, kp), zero padded columns (patch embed)."""
        return self._get(param, ("pad", kp), lambda p: K.cast_bf16(p.reshape(p.shape[0], -1), ld_out=kp))

    def frame_pool_w(self, param):
        """Conv3d(C, C, (1,8,8)) weight (O, C, 1, 8, 8) -> bf16 (O, 64*C) with columns ordered
        (hw, c) to match the token-major activation layout."""
        o, c = param.shape[0], param.shape[1]
        hw = param.shape[3] * param.shape[4]
        return self._get(param, "fpw", lambda p: K.permute_021(p.reshape(o, c, hw), o, c, hw, torch.bfloat16).reshape(o, hw * c))

    def clear(self):
        self._store.clear()
