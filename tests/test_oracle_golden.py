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
