"""Host-side logic on CPU: config surface, registry / build_model, the state_dict contract, init
parity with the reference, optimizer grouping, and the data-parallel helpers (gloo, world_size 2)."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = ["Ego4D/CSTS_Ego4D_Gaze_Forecast.yaml", "Ego4D/CSTS_Ego4D_Gaze_Estimation.yaml",
           "Aria/CSTS_Aria_Gaze_Forecast.yaml", "Aria/CSTS_Aria_Gaze_Estimation.yaml"]


def load_cfg(path, overrides=()):
    from csts_b200.host.config import get_cfg
    cfg = get_cfg()
    cfg.merge_from_file(path)
    cfg.merge_from_list(["NUM_GPUS", 0] + list(overrides))
    return cfg


@pytest.mark.parametrize("name", CONFIGS)
def test_repo_configs_load_and_build_the_same_plan(name):
    from csts_b200.host.plan import build_plan
    import csts_oracle as O
    cfg = load_cfg(os.path.join(ROOT, "configs", name))
    assert cfg.MODEL.MODEL_NAME == "CSTS" and cfg.MVIT.PATCH_KERNEL == [3, 7, 7] and cfg.SOLVER.COSINE_END_LR == 1e-6
    for s in build_plan(cfg):
        r = O.ARCH[s.name]
        assert (r[1], r[2], r[3], r[4], r[5], r[6]) == (s.kind, s.dim, s.dim_out, s.heads, s.stride_q, s.stride_kv)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ref_shim.REFERENCE_ROOT, "configs")), reason="reference tree (with configs/) not mounted")
@pytest.mark.parametrize("name", CONFIGS)
def test_reference_yaml_loads_unchanged_and_equals_repo_yaml(name):
    import yaml
    ref_path = os.path.join(ref_shim.REFERENCE_ROOT, "configs", name)
    a = load_cfg(ref_path)
    b = load_cfg(os.path.join(ROOT, "configs", name))
    assert a == b
    assert yaml.safe_load(open(ref_path)) == yaml.safe_load(open(os.path.join(ROOT, "configs", name)))


def test_cfg_node_surface():
    from csts_b200.host.config import assert_and_infer_cfg, get_cfg
    cfg = get_cfg()
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MVIT.NOT_A_KEY", 1])
    cfg.merge_from_list(["MVIT.DROPPATH_RATE", "0.3", "NUM_GPUS", 2, "TRAIN.BATCH_SIZE", 16, "TEST.BATCH_SIZE", 16])
    assert cfg.MVIT.DROPPATH_RATE == 0.3
    c2 = cfg.clone()
    c2.MVIT.DEPTH = 3
    assert cfg.MVIT.DEPTH == 16
    assert "MVIT" in cfg.dump()
    assert_and_infer_cfg(cfg)
    cfg.TRAIN.BATCH_SIZE = 15
    with pytest.raises(AssertionError):
        assert_and_infer_cfg(cfg)


@pytest.fixture(scope="module")
def cpu_model():
    from csts_b200.host.build import MODEL_REGISTRY, build_model
    cfg = load_cfg(os.path.join(ROOT, "configs", CONFIGS[0]), ["MODEL.LOSS_FUNC", "kldiv+egonce"])
    torch.manual_seed(0)
    model = build_model(cfg)
    assert "CSTS" in MODEL_REGISTRY and type(model) is MODEL_REGISTRY.get("CSTS")
    return model, cfg


def test_registry_errors():
    from csts_b200.host.build import MODEL_REGISTRY
    with pytest.raises(KeyError):
        MODEL_REGISTRY.get("NoSuchModel")


def test_state_dict_contract(cpu_model, golden_dir):
    model, _ = cpu_model
    shapes = json.load(open(os.path.join(golden_dir, "param_shapes.json")))
    sd = model.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                     # names AND order of the reference
    assert all(list(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(p.numel() for p in model.parameters()) == 188182401


def test_without_nce_there_are_no_projection_heads():
    from csts_b200.host.build import build_model
    cfg = load_cfg(os.path.join(ROOT, "configs", CONFIGS[0]))
    assert cfg.MODEL.LOSS_FUNC == "kldiv"
    model = build_model(cfg)
    assert not hasattr(model, "vision_proj") and len(model.state_dict()) == 520


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_init_is_bitwise_the_reference_init(cpu_model):
    model, _ = cpu_model
    ref, _ = ref_shim.reference_model(seed=0, overrides=["MVIT.DROPPATH_RATE", 0.2])
    a, b = model.state_dict(), ref.state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_optimizer_grouping(cpu_model):
    from csts_b200.host.train_step import construct_optimizer
    model, cfg = cpu_model
    opt = construct_optimizer(model, cfg)
    decay, no_decay = opt.param_groups
    assert decay["weight_decay"] == 0.05 and no_decay["weight_decay"] == 0.0
    assert len(decay["params"]) + len(no_decay["params"]) == 524
    assert all(p.dim() > 1 for p in decay["params"])
    names = {id(p): n for n, p in model.named_parameters()}
    # position embeddings are 3-D and DO get weight decay (ZERO_DECAY_POS_CLS False), SURVEY.md App. C
    assert "pos_embed_spatial" in {names[id(p)] for p in decay["params"]}
    assert all(p.dim() == 1 for p in no_decay["params"])


def test_drop_path_masks_are_drawn_once_per_step(cpu_model):
    """DropPath (ref common.py:46-59): scale is 0 or 1/keep per sample, E[scale] = 1, rate = the block's
    linearly increasing drop probability (ref custom_multimodal_builder.py:90)."""
    model, _ = cpu_model                                              # built with DROPPATH_RATE 0.2
    torch.manual_seed(0)
    model._draw_drop_path(20000, torch.device("cpu"))
    sites = model._dp_site
    blocks = [b for b in model.blocks if b.spec.drop_path > 0]
    assert len(blocks) == 15 and all(id(b) in sites for b in blocks)
    for b in (blocks[0], blocks[7], blocks[-1]):
        sc = model._dp_scales[sites[id(b)]]
        assert sc.shape == (2, 20000)                                 # attention branch, MLP branch (attention.py:242, :247)
        keep = 1.0 - b.spec.drop_path
        vals = torch.unique(sc)
        assert all(min(abs(v - 0.0), abs(v - 1.0 / keep)) < 1e-6 for v in vals.tolist())
        assert abs((sc > 0).float().mean().item() - keep) < 0.012
        # the two branches of a block are dropped independently: P(both kept) = keep^2, not keep
        both = ((sc[0] > 0) & (sc[1] > 0)).float().mean().item()
        assert abs(both - keep * keep) < 0.012, (both, keep)
    assert abs(blocks[-1].spec.drop_path - 0.2) < 1e-6


def test_fused_optimizer_state_interchanges_with_torch_adamw():
    """FusedClipAdamW keeps torch.optim.AdamW's state layout (step / exp_avg / exp_avg_sq per parameter, same
    param_groups), so an optimizer checkpoint written by the reference loop loads into it and vice versa."""
    from csts_b200.host.optimizer import FusedClipAdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    groups = lambda q: [{"params": [q[0]], "weight_decay": 0.05}, {"params": [q[1]], "weight_decay": 0.0}]
    ref = torch.optim.AdamW(groups(ps), lr=1e-3, eps=1e-8, weight_decay=0.05)
    for _ in range(2):
        for p in ps:
            p.grad = torch.randn_like(p)
        ref.step()
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ours = FusedClipAdamW(groups(qs), lr=1e-3, eps=1e-8, weight_decay=0.05)
    ours.load_state_dict(ref.state_dict())
    ours._init_state()
    assert ours._step.item() == 2.0
    for p, q in zip(ps, qs):
        assert torch.equal(ours.state[q]["exp_avg"], ref.state[p]["exp_avg"])
        assert torch.equal(ours.state[q]["exp_avg_sq"], ref.state[p]["exp_avg_sq"])
    back = torch.optim.AdamW(groups([torch.nn.Parameter(p.detach().clone()) for p in ps]), lr=1e-3, eps=1e-8, weight_decay=0.05)
    back.load_state_dict(ours.state_dict())
    assert [g["weight_decay"] for g in back.param_groups] == [0.05, 0.0]
    assert float(back.state[back.param_groups[0]["params"][0]]["step"]) == 2.0


def test_weight_cache_does_not_trust_version_counters():
    """torch's fused optimizers update parameters without bumping `_version`; in training the 16-bit copies are
    therefore re-made once per step (begin_training_step) unless the fused clip+AdamW step has just rewritten them."""
    from csts_b200.host.weights import WeightCache
    wc = WeightCache()
    calls = []
    p = torch.nn.Parameter(torch.zeros(2, 2))

    def make(q, out):
        calls.append(out)
        if out is None:
            return q.clone()
        out.copy_(q)
        return out

    first = wc._get(p, "w", make)
    wc._get(p, "w", make)
    assert len(calls) == 1                       # same step, same version: cached
    wc.begin_training_step()
    again = wc._get(p, "w", make)
    assert len(calls) == 2                       # new step: rebuilt even though p._version is unchanged
    assert calls[1] is first and again is first  # ... IN PLACE: captured graphs / the optimizer's pointer table hold its address
    wc._get(p, ("pad", 8), make)
    wc.after_fused_step()                        # the fused optimizer refreshed the plain copies, dropped the re-laid-out ones
    wc.begin_training_step()
    wc._get(p, "w", make)
    assert len(calls) == 3
    wc._get(p, ("pad", 8), make)
    assert len(calls) == 4
    # evaluation after training: a non-training forward rebuilds the copies once, whatever the optimizer was
    wc.begin_training_step()
    wc._get(p, "w", make)
    n = len(calls)
    with torch.no_grad():
        p.add_(1.0)                              # what torch's fused AdamW does: no version bump visible to the cache key
    wc.begin_inference()
    assert torch.equal(wc._get(p, "w", make), p.detach())
    assert len(calls) == n + 1
    wc.begin_inference()
    wc._get(p, "w", make)
    assert len(calls) == n + 1                   # a second evaluation forward hits the cache


def test_cpu_forward_fails_loudly(cpu_model):
    model, _ = cpu_model
    with pytest.raises(RuntimeError):
        model([torch.zeros(1, 3, 8, 256, 256)], torch.zeros(1, 1, 8, 256, 256))


# ------------------------------------------------------------------------------------------- gloo, world 2
def _nce_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csts_oracle as O
    from csts_b200.host import distributed as du
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(world * 3, 2, 16, generator=g)         # global batch of (video, audio) features
    w = torch.randn(16, 8, generator=g).requires_grad_(True)    # the shared "model"
    local = feats[rank * 3:(rank + 1) * 3]
    v, a = local[:, 0] @ w, local[:, 1] @ w
    assert du.get_world_size() == world and du.get_rank() == rank
    vg, ag = du.all_gather_with_grad([v, a])
    assert vg.shape == (world * 3, 8)
    loss = O.egonce(O.sim_matrix(vg, ag))
    loss.backward()
    grad = w.grad.clone()
    dist.all_reduce(grad)
    grad /= world                                              # what DDP's mean all-reduce produces
    # helpers used by the loop
    t = torch.tensor([float(rank + 1)])
    du.all_reduce([t])
    gathered = du.all_gather([torch.full((2, 1), float(rank))])[0]
    if rank == 0:
        out.put((loss.item(), grad, t.item(), gathered))
    dist.destroy_process_group()


def test_nce_all_gather_gradient_equals_single_process_global_batch():
    import csts_oracle as O
    world, port = 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    loss, grad, reduced, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(world * 3, 2, 16, generator=g)
    w = torch.randn(16, 8, generator=g).requires_grad_(True)
    ref = O.egonce(O.sim_matrix(feats[:, 0] @ w, feats[:, 1] @ w))
    ref.backward()
    assert abs(loss - ref.item()) < 1e-6
    torch.testing.assert_close(grad, w.grad, rtol=1e-5, atol=1e-6)     # the reference's ctx.rank = 0 bug would fail this
    assert reduced == 1.5
    assert gathered.flatten().tolist() == [0.0, 0.0, 1.0, 1.0]


def test_grad_arena_layout_and_adoption():
    """host/grad_arena.py: slots in reverse registration order, 16-byte aligned, a block's parameters contiguous;
    adopt() moves a foreign gradient into its slot; views are fresh objects (autograd steals them without a copy)."""
    from csts_b200.host.grad_arena import GradArena, grad_slot
    ps = [torch.nn.Parameter(torch.randn(*s)) for s in [(3, 5), (7,), (2, 2, 2), (9,)]]
    ar = GradArena(ps)
    offs = [ar.slot[id(p)][0] for p in ps]
    assert offs == sorted(offs, reverse=True) and all(o % 4 == 0 for o in offs)
    assert ar.numel == 12 + 8 + 8 + 16 and ar.matches(ps) and not ar.matches(ps[:3])
    assert ar.view(ps[0]).shape == (3, 5) and ar.view(ps[0]) is not ar.view(ps[0])
    lo, hi = ar.range_of(ps[1:3])
    assert (lo, hi) == (offs[2], offs[1] + 8)
    assert ar.span(ps[1:3]).numel() == offs[1] + 7 - offs[2]
    ps[1].grad = torch.arange(7.0)
    assert not ar.owns(ps[1].grad, ps[1])
    ar.adopt(ps[1])
    assert ar.owns(ps[1].grad, ps[1]) and torch.equal(ar.flat[offs[1]: offs[1] + 7], torch.arange(7.0))

    class WC:
        arena = ar
    assert grad_slot(WC, ps[1]) is None                      # already holds a gradient: autograd must accumulate
    assert grad_slot(WC, ps[0]).data_ptr() == ar.flat.data_ptr() + 4 * offs[0]
    assert grad_slot(WC, torch.nn.Parameter(torch.zeros(1))) is None
    # a gradient written into the slot and handed to autograd is adopted as p.grad without a copy
    x = ps[0]

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            ctx.w = w
            return w.sum()

        @staticmethod
        def backward(ctx, g):
            out = grad_slot(WC, ctx.w)
            out.fill_(2.0)
            return out
    Fn.apply(x).backward()
    assert ar.owns(x.grad, x) and float(x.grad.sum()) == 30.0


def test_overlapped_grad_sync_buckets_keep_factored_parameters_apart():
    """host/distributed.py::OverlappedGradSync: buckets are contiguous arena slices in backward order; a parameter whose
    gradient is exchanged through its factors (the frame-pool kernels) forms a group of its own that is never all-reduced,
    and the groups on either side of it stay contiguous."""
    from csts_b200.host import distributed as du
    from csts_b200.host.grad_arena import GradArena

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.f = torch.nn.Parameter(torch.zeros(5))
            self.e = torch.nn.Parameter(torch.zeros(5))
            self.d = torch.nn.Parameter(torch.zeros(5))
            self.a = torch.nn.Parameter(torch.zeros(10))
            self.pool = torch.nn.Parameter(torch.zeros(64, 8))
            self.b = torch.nn.Parameter(torch.zeros(6))
            self.c = torch.nn.Parameter(torch.zeros(30))

        def factored_grad_params(self):
            return [self.pool]
    m = Model()
    # 100-byte buckets in backward order; the last 90 bytes of gradient (what arrives when backward ends) go in 40-byte ones
    sync = du.OverlappedGradSync(m, bucket_bytes=100, tail_bytes=90, tail_bucket_bytes=40)
    arena = GradArena(list(m.parameters()))
    sync._bind(arena)
    names = {id(p): n for n, p in m.named_parameters()}
    assert [[names[id(p)] for p in g] for g in sync.groups] == [["c"], ["b"], ["pool"], ["a"], ["d", "e"], ["f"]]
    assert sync.external == [False, False, True, False, False, False]
    for g, (lo, hi) in zip(sync.groups, sync.ranges):
        assert (lo, hi) == arena.range_of(g) and hi - lo >= sum(p.numel() for p in g)
    # the identity the factored exchange rests on: the rank-mean of thin products is one product over concatenated rows
    gen = torch.Generator().manual_seed(3)
    dys, xs = [torch.randn(4, 6, generator=gen) for _ in range(3)], [torch.randn(4, 9, generator=gen) for _ in range(3)]
    mean = sum(dy.t() @ x for dy, x in zip(dys, xs)) / 3
    torch.testing.assert_close(torch.cat(dys).t() @ torch.cat(xs) / 3, mean, rtol=1e-5, atol=1e-6)


def _k400_like_checkpoint(model_state, tmp_path):
    """A pre-trained-MViT-like checkpoint: 224-pixel / 16-frame position embeddings, a DDP-style "module." prefix, one
    tensor of a different shape (the Kinetics classification head) and one unknown name."""
    g = torch.Generator().manual_seed(11)
    pre = {"module." + k: torch.randn(v.shape, generator=g) for k, v in list(model_state.items())[:40]}
    pre["module.pos_embed_spatial"] = torch.randn(1, 56 * 56, 96, generator=g)
    pre["module.pos_embed_temporal"] = torch.randn(1, 8, 96, generator=g)
    pre["module.pos_embed_spatial_audio"] = torch.randn(1, 56 * 56, 96, generator=g)   # not in the reference's interpolate list
    pre["module.blocks.0.attn.qkv.weight"] = torch.randn(288, 96, generator=g)
    pre["module.blocks.0.attn.proj.bias"] = torch.randn(7, generator=g)          # wrong shape: must be skipped
    pre["module.head.projection.weight"] = torch.randn(400, 768, generator=g)    # no such tensor in CSTS
    path = os.path.join(tmp_path, "k400_like.pyth")
    torch.save({"epoch": 199, "model_state": pre, "optimizer_state": {}, "cfg": "dummy"}, path)
    return path, pre


def test_checkpoint_finetune_load_matches_by_name_and_shape_and_interpolates_pos_embed(cpu_model, tmp_path):
    """host/checkpoint.py vs slowfast/utils/checkpoint.py:290-354 (run live when the reference is importable)."""
    from csts_b200.host import checkpoint as cu
    from csts_b200.host.build import build_model
    _, cfg = cpu_model
    torch.manual_seed(1)
    model = build_model(cfg)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    path, pre = _k400_like_checkpoint(before, str(tmp_path))
    epoch = cu.load_checkpoint(path, model, data_parallel=False, epoch_reset=True, clear_name_pattern=("module.",))
    assert epoch == -1
    after = model.state_dict()
    assert torch.equal(after["blocks.0.attn.qkv.weight"], pre["module.blocks.0.attn.qkv.weight"])
    assert torch.equal(after["blocks.0.attn.proj.bias"], before["blocks.0.attn.proj.bias"])          # shape mismatch: untouched
    assert "blocks.0.attn.proj.bias" in cu.load_checkpoint.last_not_loaded
    assert "pos_embed_spatial" not in cu.load_checkpoint.last_not_loaded
    want = torch.nn.functional.interpolate(pre["module.pos_embed_spatial"].unsqueeze(0), (4096, 96), mode="bilinear").squeeze(0)
    assert torch.equal(after["pos_embed_spatial"], want) and after["pos_embed_temporal"].shape == (1, 4, 96)
    assert not torch.equal(after["pos_embed_temporal"], before["pos_embed_temporal"])
    assert torch.equal(after["pos_embed_spatial_audio"], before["pos_embed_spatial_audio"])           # not in the interpolate list
    if ref_shim.reference_available():
        ref_shim.install()
        try:
            import slowfast.utils.checkpoint as rcu
        except Exception as e:          # the module pulls optional packages the image may lack
            pytest.skip(f"reference checkpoint module not importable here: {e}")
        torch.manual_seed(1)
        twin = build_model(cfg)
        rcu.load_checkpoint(path, twin, data_parallel=False, epoch_reset=True, clear_name_pattern=("module.",))
        ts = twin.state_dict()
        assert all(torch.equal(after[k], ts[k]) for k in after)


def test_checkpoint_round_trip_resumes_epoch_and_optimizer(cpu_model, tmp_path):
    from csts_b200.host import checkpoint as cu
    model, cfg = cpu_model
    opt = torch.optim.AdamW([p for p in model.parameters()][:3], lr=1e-3)
    for p in opt.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    opt.step()
    path = cu.save_checkpoint(os.path.join(str(tmp_path), "checkpoints", "checkpoint_epoch_00003.pyth"), model, opt, 2, cfg,
                              data_parallel=False)
    rec = torch.load(path, weights_only=False)
    assert set(rec) == {"epoch", "model_state", "optimizer_state", "cfg"} and list(rec["model_state"]) == list(model.state_dict())
    opt2 = torch.optim.AdamW([p for p in model.parameters()][:3], lr=1e-3)
    assert cu.load_checkpoint(path, model, data_parallel=False, optimizer=opt2) == 2
    assert float(opt2.state[opt2.param_groups[0]["params"][0]]["step"]) == 1.0
