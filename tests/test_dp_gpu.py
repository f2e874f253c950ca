"""Data-parallel step on 2 GPUs (NCCL): gradients after the NCE all-gather + overlapped gradient
all-reduce equal the single-process global-batch gradients (up to bf16 rounding).  Skipped on boxes
with fewer than 2 GPUs; the host-side logic is covered on CPU (gloo) in tests/test_host_cpu.py."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csts_oracle as O
    from csts_b200.host import distributed as du
    from csts_b200.host.build import build_model
    from csts_b200.host.config import get_cfg
    from csts_b200.host.train_step import compute_loss
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "Ego4D", "CSTS_Ego4D_Gaze_Forecast.yaml"))
    cfg.merge_from_list(["NUM_GPUS", world, "MODEL.LOSS_FUNC", "kldiv+egonce", "MVIT.DROPPATH_RATE", 0.0,
                         "TRAIN.BATCH_SIZE", 2 * world, "TEST.BATCH_SIZE", 2 * world])
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "param_shapes.json")))
    sd = O.synthetic_state(shapes, seed=0)
    model = build_model(cfg, ddp=False)
    model.load_state_dict(sd)
    model.train()
    video, audio, hm = (t.cuda() for t in O.synthetic_batch(2 * world, seed=1))
    lo = 2 * rank
    sync = du.OverlappedGradSync(model)
    loss, _, _, _ = compute_loss(cfg, model, [video[lo:lo + 2]], audio[lo:lo + 2], hm[lo:lo + 2])
    sync.start()
    loss.backward()
    sync.finish()
    torch.cuda.synchronize()
    if rank == 0:
        dp = {n: p.grad.clone() for n, p in model.named_parameters()}
        in_arena = all(model._wc.arena.owns(p.grad, p) for p in model.parameters())      # every bucket was reduced in place
        # single-process global batch on the same weights (world size is still 2 for the process group, so
        # evaluate the loss by hand without the gather)
        for p in model.parameters():
            p.grad = None
        from csts_b200.host import losses
        from csts_b200.host.utils import frame_softmax, sim_matrix
        logits, v, a = model([video], audio, return_embed=True)
        kld = losses.get_loss_func("kldiv")()(frame_softmax(logits, 2), hm)
        nce = losses.get_loss_func("egonce")()(sim_matrix(v, a))
        (kld + cfg.MODEL.LOSS_ALPHA * nce).backward()
        num = sum((dp[n] - p.grad).pow(2).sum().item() for n, p in model.named_parameters())
        den = sum(p.grad.pow(2).sum().item() for p in model.parameters())
        # yardstick: the single-process step evaluated a second time (summation-order noise of 16-bit gradient storage)
        first = {n: p.grad.clone() for n, p in model.named_parameters()}
        for p in model.parameters():
            p.grad = None
        logits, v, a = model([video], audio, return_embed=True)
        kld = losses.get_loss_func("kldiv")()(frame_softmax(logits, 2), hm)
        nce = losses.get_loss_func("egonce")()(sim_matrix(v, a))
        (kld + cfg.MODEL.LOSS_ALPHA * nce).backward()
        noise = (sum((first[n] - p.grad).pow(2).sum().item() for n, p in model.named_parameters()) / den) ** 0.5
        worst = max(((dp[n] - p.grad).norm() / p.grad.norm()).item() for n, p in model.named_parameters() if p.grad.norm() > 1e-5)
        # the three frame-pool kernels are averaged from their all-gathered factors, not all-reduced
        pools = {n: ((dp[n] - p.grad).norm() / p.grad.norm()).item() for n, p in model.named_parameters()
                 if any(p is f for f in model.factored_grad_params())}
        assert len(pools) == 3 and len(sync._factor_bufs) == 3 and sum(sync.external) == 3
        q.put(((num / den) ** 0.5, worst, in_arena, noise, pools))
    dist.barrier()
    os._exit(0)


def test_two_gpu_gradients_match_global_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    rel, worst, in_arena, noise, pools = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    print("global-batch vs data-parallel gradient, relative L2:", rel, "worst tensor:", worst, "run-to-run noise of the global step:", noise,
          "frame-pool kernels (factored exchange):", pools)
    # Samples are independent through the network, so both sides evaluate the same per-sample arithmetic; what differs is
    # what differs between any two runs of ONE computation — f32 summation order (split-K reduce-adds, the cross-rank
    # reduction) and the 16-bit rounding flips it causes (tools/grad_noise.py: ~1e-2 in bf16 storage) — so the yardstick is
    # the single-process step evaluated twice.  A corrupted bucket — a gradient buffer recycled while the side-stream
    # all-reduce still reads it — would be an O(1) error in a large tensor.
    assert in_arena, "a gradient was not reduced inside the arena"
    assert rel < max(2.5 * noise, 5e-3), (rel, noise)       # measured: rel 8e-3 beside a run-to-run floor of 1e-2 (bf16 storage)
    assert worst < 0.3, worst
    assert max(pools.values()) < max(5 * noise, 2e-2), pools
