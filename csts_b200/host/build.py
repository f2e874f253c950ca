"""Model construction — the drop-in boundary.  Mirrors ``slowfast/models/build.py:18-47``:
``build_model(cfg, gpu_id=None)`` looks the class up in ``MODEL_REGISTRY`` by
``cfg.MODEL.MODEL_NAME``, constructs it with ``cfg``, moves it to the current CUDA device and wraps
it in ``DistributedDataParallel`` when ``cfg.NUM_GPUS > 1``."""
import torch

from .registry import Registry

MODEL_REGISTRY = Registry("MODEL")
MODEL_REGISTRY.__doc__ = """Registry for video models: the registered object is called as obj(cfg)
and returns a torch.nn.Module."""


def build_model(cfg, gpu_id=None, ddp=True):
    """`ddp=False` (extension of the reference signature) returns the bare replica with rank 0's
    parameters broadcast, for the CUDA-graph data-parallel step whose gradient exchange is a captured
    NCCL all-reduce (train_step.GraphedTrainStep) instead of DDP's host-driven reducer."""
    from . import csts  # noqa: F401  (registers CSTS)
    if torch.cuda.is_available():
        assert cfg.NUM_GPUS <= torch.cuda.device_count(), "Cannot use more GPU devices than available"
    else:
        assert cfg.NUM_GPUS == 0, "Cuda is not available. Please set `NUM_GPUS: 0 for running on CPUs."
    model = MODEL_REGISTRY.get(cfg.MODEL.MODEL_NAME)(cfg)
    if cfg.NUM_GPUS:
        cur_device = torch.cuda.current_device() if gpu_id is None else gpu_id
        model = model.cuda(device=cur_device)
    if cfg.NUM_GPUS > 1 and not ddp:
        from . import distributed as du
        du.broadcast_parameters(model)
        return model
    if cfg.NUM_GPUS > 1:
        # Gradient all-reduce over NVLink/NVSwitch, bucketed and overlapped with backward.  Every
        # parameter receives a gradient when return_embed=True, so no unused-
        # parameter search; 128 MB buckets suit the three 151 MB frame-pool gradients and NVLS.
        model = torch.nn.parallel.DistributedDataParallel(
            module=model, device_ids=[cur_device], output_device=cur_device,
            bucket_cap_mb=128, gradient_as_bucket_view=True)
    return model
