"""Host-side mirror of the reference's interface for the CSTS hot path.

  config.py       get_cfg() / CfgNode: the cfg keys the path reads (ref slowfast/config/defaults.py, custom_config.py)
  build.py        MODEL_REGISTRY / build_model(cfg, gpu_id)            (ref slowfast/models/build.py:18-47)
  csts.py         class CSTS(nn.Module) with the reference's parameter names (ref custom_multimodal_builder.py)
  block.py        transformer block forward/backward on the CUDA kernels (ref attention.py, av_attention.py)
  losses.py       KLDiv, EgoNCE, get_loss_func                           (ref slowfast/models/losses.py)
  utils.py        frame_softmax, sim_matrix                              (ref slowfast/utils/utils.py)
  distributed.py  all_gather_with_grad, all_reduce, all_gather            (ref slowfast/utils/distributed.py)
"""
