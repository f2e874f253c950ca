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


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
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
        keep = 1.0 - b.spec.drop_path
        vals = torch.unique(sc)
        assert all(min(abs(v - 0.0), abs(v - 1.0 / keep)) < 1e-6 for v in vals.tolist())
        assert abs((sc > 0).float().mean().item() - keep) < 0.012
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
    make = lambda q: calls.append(1) or q.clone()
    wc._get(p, "w", make)
    wc._get(p, "w", make)
    assert len(calls) == 1                       # same step, same version: cached
    wc.begin_training_step()
    wc._get(p, "w", make)
    assert len(calls) == 2                       # new step: rebuilt even though p._version is unchanged
    wc._get(p, ("pad", 8), make)
    wc.after_fused_step()                        # the fused optimizer refreshed the plain copies, dropped the re-laid-out ones
    wc.begin_training_step()
    wc._get(p, "w", make)
    assert len(calls) == 3
    wc._get(p, ("pad", 8), make)
    assert len(calls) == 4


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
