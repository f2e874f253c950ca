"""Pin oracle/csts_oracle.py against outputs of the unmodified reference (tests/golden/*, made by
oracle/make_golden.py) and, when /root/reference is present, against the live reference."""
import json
import os

import pytest
import torch

import csts_oracle as O
import ref_shim

torch.set_num_threads(os.cpu_count() or 1)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("case", ["enc_poolq", "enc_plain", "enc_kv1", "dec_hw", "dec_t", "spatial", "temporal"])
def test_unit_blocks(golden_dir, case):
    rec = _load(golden_dir, "unit_blocks.pt")[case]
    sd = {"blk." + k: v for k, v in rec["state"].items()}
    x = rec["x"].clone().requires_grad_(True)
    y, thw = O.block(sd, "blk", x, tuple(rec["thw"]), spec=rec["spec"])
    assert tuple(thw) == tuple(rec["thw_out"])
    torch.testing.assert_close(y, rec["y"], rtol=1e-5, atol=1e-5)
    (gx,) = torch.autograd.grad((y * rec["probe"]).sum(), x)
    torch.testing.assert_close(gx, rec["gx"], rtol=1e-4, atol=1e-5)


def test_losses(golden_dir):
    rec = _load(golden_dir, "losses.pt")
    p = O.frame_softmax(rec["logits"], 2.0)
    torch.testing.assert_close(p, rec["p"], rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(O.kldiv(p, rec["hm"]), rec["kld"], rtol=1e-6, atol=1e-7)
    sim = O.sim_matrix(rec["v"], rec["a"])
    torch.testing.assert_close(sim, rec["sim"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(O.egonce(sim), rec["nce"], rtol=1e-5, atol=1e-6)


def test_arch_table_matches_param_shapes(golden_dir):
    shapes = json.load(open(os.path.join(golden_dir, "param_shapes.json")))
    assert len(shapes) == 524
    assert sum(torch.Size(s).numel() for s in shapes.values()) == 188182401
    for name, (_, kind, dim, dim_out, heads, sq, skv) in O.ARCH.items():
        assert shapes[f"{name}.attn.qkv.weight"] == [3 * dim, dim]
        hidden = 4 * (dim_out if kind == "dec" else dim)
        assert shapes[f"{name}.mlp.fc1.weight"] == [hidden, dim]
        assert shapes[f"{name}.mlp.fc2.weight"] == [dim_out, hidden]
        assert (f"{name}.proj.weight" in shapes) == (dim != dim_out)
        d = dim // heads
        qname = "upsample_q" if kind == "dec" else "pool_q"
        assert (f"{name}.attn.{qname}.weight" in shapes) == (sq is not None)
        assert (f"{name}.attn.pool_k.weight" in shapes) == (skv is not None)
        if skv is not None:
            assert shapes[f"{name}.attn.pool_k.weight"] == [d, 1, 3, 3, 3]


@pytest.mark.parametrize("fixture", ["full_b2.pt", "full_b2_gain4.pt"])
def test_full_model_against_reference_golden(golden_dir, fixture):
    rec = _load(golden_dir, fixture)
    shapes = json.load(open(os.path.join(golden_dir, "param_shapes.json")))
    sd = O.synthetic_state(shapes, seed=rec["seed"], gain=rec["gain"])
    video, audio, hm = O.synthetic_batch(rec["B"], seed=rec["seed"] + 1)
    loss, kld, nce, logits, grads = O.loss_and_grads(sd, video, audio, hm, alpha=rec["alpha"])
    torch.testing.assert_close(logits, rec["logits"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(loss, rec["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(kld, rec["kld"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(nce, rec["nce"], rtol=1e-5, atol=1e-6)
    assert set(grads) == set(rec["grad_norms"])
    for n, ref_norm in rec["grad_norms"].items():
        assert abs(grads[n].norm().item() - ref_norm) <= 1e-3 * ref_norm + 1e-7, n   # a few biases have analytically zero gradient (softmax shift invariance)
    for n, g in rec.get("grads", {}).items():
        if g.norm() < 1e-6:          # analytically zero (softmax shift invariance): only noise
            assert grads[n].norm() < 1e-6, n
            continue
        err = (grads[n] - g).norm() / g.norm()
        assert err < 1e-3, (n, err.item())


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_oracle_against_live_reference_forward():
    """Different seed / eval mode than the committed fixture, forward only, B=1."""
    model, _ = ref_shim.reference_model(seed=3)
    model.eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    video, audio, _ = O.synthetic_batch(1, seed=11)
    with torch.no_grad():
        ref = model([video], audio, return_embed=True)
        got = O.csts_forward(sd, video, audio, return_embed=True)
    for r, g in zip(ref, got):
        torch.testing.assert_close(g, r, rtol=1e-4, atol=2e-5)


def test_optional_paths_against_reference_golden(golden_dir):
    """MVIT.SPATIAL_AUDIO_ATTN=True and the return_spatial_attn / return_temporal_attn outputs
    (custom_multimodal_builder.py:425-440,448-451,483-491; av_attention.py:356-370), B=1."""
    rec = _load(golden_dir, "optional_b1.pt")
    shapes = json.load(open(os.path.join(golden_dir, "param_shapes.json")))
    sd = O.synthetic_state(shapes, seed=rec["seed"], gain=2.0)
    video, audio, hm = O.synthetic_batch(rec["B"], seed=rec["seed"] + 1)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits, v, a = O.csts_forward(leaves, video, audio, return_embed=True, spatial_audio_attn=True)
    assert (logits - rec["saa_logits"]).abs().max() < 2e-4
    assert (v - rec["saa_v"]).abs().max() < 2e-4
    kld = O.kldiv(O.frame_softmax(logits, 2.0), hm)
    assert abs(kld.item() - rec["saa_kld"].item()) < 1e-5 * abs(rec["saa_kld"].item())
    kld.backward()
    for n, g in rec["saa_grads"].items():
        assert ((leaves[n].grad - g).norm() / g.norm()).item() < 2e-3, n
    worst = max(abs(leaves[n].grad.norm().item() - gn) / gn for n, gn in rec["saa_grad_norms"].items() if gn > 1e-6)
    assert worst < 5e-3, worst
    # the flag changes the result (the fixture is not vacuous)
    assert (O.csts_forward(sd, video, audio) - rec["saa_logits"]).abs().max() > 1e-3
    with torch.no_grad():
        out = O.csts_forward(sd, video, audio, return_spatial_attn=True, return_temporal_attn=True)
    assert len(out) == 3 and out[1].shape == (1, 8, 260, 260) and out[2].shape == (1, 8, 8, 8)
    assert (out[0] - rec["attn_logits"]).abs().max() < 2e-4
    assert (out[2] - rec["temporal_attn"]).abs().max() < 1e-5
    assert (out[1][:, :, ::13, :] - rec["spatial_attn_rows"]).abs().max() < 1e-5
    assert (out[1].sum(-1) - rec["spatial_attn_rowsum"]).abs().max() < 1e-5
    assert len(O.csts_forward(sd, video, audio, return_temporal_attn=True)) == 2


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_oracle_drop_path_against_live_reference_block():
    """DropPath (common.py:46-59, applied at attention.py:242 and :247 with independent draws): the reference block in
    train mode with torch.rand replaced by a recorded sequence equals the oracle block given the same two scale vectors."""
    ref_shim.install()
    from functools import partial
    import slowfast.models.common as common
    from slowfast.models.attention import MultiScaleBlock
    torch.manual_seed(0)
    blk = MultiScaleBlock(dim=96, dim_out=192, num_heads=1, qkv_bias=True, drop_path=0.4, kernel_q=(), kernel_kv=(3, 3, 3),
                          stride_q=(), stride_kv=(1, 2, 2), norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), mode="conv",
                          has_cls_embed=False)
    blk.train()
    x = torch.randn(4, 256, 96)
    draws = [torch.tensor([0.9, 0.1, 0.7, 0.3]).reshape(4, 1, 1), torch.tensor([0.2, 0.95, 0.61, 0.05]).reshape(4, 1, 1)]
    seq = list(draws)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: seq.pop(0)
    try:
        y_ref, _ = blk(x, (4, 8, 8))
    finally:
        torch.rand = real_rand
    assert not seq
    keep = 0.6
    scales = tuple(torch.floor(keep + d.reshape(-1)) / keep for d in draws)
    assert len(set(scales[0].tolist())) == 2 and float(scales[0].min()) == 0.0
    sd = {"b." + k: v.detach() for k, v in blk.state_dict().items()}
    y, _ = O.block(sd, "b", x, (4, 8, 8), spec=("enc", 96, 192, 1, None, (1, 2, 2)), drop=scales)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    y0, _ = O.block(sd, "b", x, (4, 8, 8), spec=("enc", 96, 192, 1, None, (1, 2, 2)))
    assert (y0 - y_ref).abs().max() > 1e-2


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("dataset", ["ego4d_av_gaze_forecast", "aria_av_gaze_forecast", "ego4d_av_gaze"])
def test_oracle_adaptive_f1_against_live_reference(dataset):
    """oracle adaptive_f1 / minmax_rescale vs slowfast/utils/metrics.py:9-74 and tools/train_avgaze_net.py:125-127."""
    ref_shim.install()
    from slowfast.utils import metrics as rmetrics
    g = torch.Generator().manual_seed(2)
    _, _, hm = O.synthetic_batch(3, seed=5)
    logits = torch.randn(3, 1, 8, 64, 64, generator=g) + 6.0 * hm.unsqueeze(1) / hm.amax(dim=(-1, -2), keepdim=True).unsqueeze(1)
    preds = O.minmax_rescale(O.frame_softmax(logits, 2.0))
    if "forecast" in dataset:
        preds = preds * 0.08                                # put the maps inside the dataset's threshold grid
    labels = torch.rand(3, 8, 3, generator=g)
    labels[:, :, 2] = (torch.rand(3, 8, generator=g) > 0.3).float() * 0      # tracked frames: type 0 ...
    labels[0, 1, 2] = 1.0                                                    # ... and two untracked ones
    labels[2, 5, 2] = 2.0
    want = rmetrics.adaptive_f1(preds, hm, labels, dataset)
    got = O.adaptive_f1(preds, hm, labels, dataset)
    assert all(abs(a - b) <= 1e-6 for a, b in zip(got, want)), (got, want)
    assert 0.0 < want[0] < 1.0
