"""Generate tests/golden/* by running the UNMODIFIED reference (imported read-only from
/root/reference through oracle/ref_shim.py) on seeded inputs.  TEST INFRASTRUCTURE.

Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_golden.py [--skip-full]

Fixtures written:
  param_shapes.json   the 524-entry state_dict (name -> shape) contract      
  unit_blocks.pt      reference block modules at reduced width: inputs, weights, outputs, input-grads
  losses.pt           frame_softmax / KLDiv / sim_matrix / EgoNCE known answers
  full_b2.pt          full CSTS fwd + kldiv+egonce + backward at B=2 on synthetic_state(seed 0):
                      logits, embeddings, loss terms, all 524 per-tensor gradient norms and a few
                      small gradients in full.
  optional_b1.pt      the config-reachable optional paths at B=1 (python oracle/make_golden.py --only-optional):
                      MVIT.SPATIAL_AUDIO_ATTN=True (logits, loss, gradient norms) and the
                      return_spatial_attn / return_temporal_attn attention maps of the default model.
"""
import argparse
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import csts_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def unit_blocks():
    ref_shim.install()
    from functools import partial
    import torch.nn as nn
    from slowfast.models.attention import MultiScaleBlock, MultiScaleDecoderBlock
    from slowfast.models.av_attention import SpatialBlock, TemporalBlock
    norm = partial(nn.LayerNorm, eps=1e-6)
    cases = {
        # name: (ctor, kind, dim, dim_out, heads, stride_q, stride_kv, thw, extra tokens)
        "enc_poolq": (MultiScaleBlock, "enc", 32, 64, 2, (1, 2, 2), (1, 2, 2), (2, 8, 8), 0),
        "enc_plain": (MultiScaleBlock, "enc", 32, 32, 2, None, (1, 4, 4), (2, 8, 8), 0),
        "enc_kv1": (MultiScaleBlock, "enc", 48, 48, 3, (1, 2, 2), (1, 1, 1), (2, 4, 4), 0),
        "dec_hw": (MultiScaleDecoderBlock, "dec", 64, 32, 2, (1, 2, 2), (1, 2, 2), (2, 4, 4), 0),
        "dec_t": (MultiScaleDecoderBlock, "dec", 32, 16, 1, (2, 1, 1), (1, 4, 4), (2, 8, 8), 0),
        "spatial": (SpatialBlock, "spatial", 32, 32, 2, None, None, (3, 2, 2), 3),
        "temporal": (TemporalBlock, "temporal", 32, 32, 2, None, None, (2, 2, 2), 0),
    }
    out = {}
    for name, (ctor, kind, dim, dim_out, heads, sq, skv, thw, extra) in cases.items():
        torch.manual_seed(hash(name) % 1000)
        kw = dict(dim=dim, dim_out=dim_out, num_heads=heads, mlp_ratio=4.0, qkv_bias=True,
                  norm_layer=norm, mode="conv", has_cls_embed=False, pool_first=False,
                  kernel_q=[3, 3, 3] if sq else [1, 1, 1], kernel_kv=[3, 3, 3] if skv else [1, 1, 1],
                  stride_q=list(sq) if sq else [1, 1, 1], stride_kv=list(skv) if skv else [1, 1, 1])
        if kind == "enc" and sq is None:
            kw["kernel_q"], kw["stride_q"] = [], []
        m = ctor(**kw)
        with torch.no_grad():
            for p in m.parameters():           # non-trivial biases / affine parameters
                if p.ndim == 1:
                    p.add_(0.1 * torch.randn_like(p))
        B = 2
        n = thw[0] * thw[1] * thw[2] + extra
        x = torch.randn(B, n, dim, requires_grad=True)
        y, thw_out = m(x, list(thw))[:2]
        probe = torch.randn_like(y)
        (gx,) = torch.autograd.grad((y * probe).sum(), x)
        out[name] = dict(spec=(kind, dim, dim_out, heads, sq, skv), thw=thw, thw_out=tuple(thw_out),
                         state={k: v.detach().clone() for k, v in m.state_dict().items()},
                         x=x.detach(), y=y.detach(), probe=probe, gx=gx)
        print("unit", name, tuple(y.shape), tuple(thw_out))
    return out


def losses():
    ref_shim.install()
    from slowfast.models import losses as L
    from slowfast.utils.utils import frame_softmax, sim_matrix
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(3, 1, 8, 64, 64, generator=g) * 3
    _, _, hm = O.synthetic_batch(3, seed=3)
    p = frame_softmax(logits, temperature=2)
    kld = L.get_loss_func("kldiv")()(p, hm)
    v = torch.randn(5, 256, generator=g)
    a = torch.randn(5, 256, generator=g) + 0.5 * v
    sim = sim_matrix(v, a)
    _, _, nce = ref_shim.reference_loss(torch.zeros(5, 1, 8, 64, 64), v, a, torch.full((5, 8, 64, 64), 1 / 4096.))
    return dict(logits=logits, hm=hm, p=p, kld=kld, v=v, a=a, sim=sim, nce=nce)


def full(B=2, seed=0, gain=1.0, tag="full_b2"):
    model, cfg = ref_shim.reference_model(seed=0)
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = O.synthetic_state(shapes, seed=seed, gain=gain)
    model.load_state_dict(sd, strict=True)
    model.train()
    video, audio, hm = O.synthetic_batch(B, seed=seed + 1)
    t0 = time.time()
    logits, v, a = model([video], audio, return_embed=True)
    loss, kld, nce = ref_shim.reference_loss(logits, v, a, hm, alpha=cfg.MODEL.LOSS_ALPHA)
    loss.backward()
    print(f"{tag}: reference fwd+bwd {time.time() - t0:.1f}s loss {loss.item():.6f} kld {kld.item():.6f} nce {nce.item():.6f}")
    grads = {n: p.grad for n, p in model.named_parameters()}
    assert all(g is not None for g in grads.values())
    keep = ["classifier.weight", "classifier.bias", "pos_embed_temporal", "pos_embed_temporal_audio",
            "blocks.0.attn.pool_k.weight", "blocks.1.attn.pool_q.weight", "decode_block1.attn.upsample_q.weight",
            "decode_block4.attn.upsample_q.weight", "blocks.0.norm1.weight", "blocks.15.norm2.bias",
            "spatial_fusion.attn.qkv.bias", "temporal_fusion.mlp.fc2.bias", "vision_proj.weight",
            "audio_proj.bias", "patch_embed.proj.weight", "patch_embed_audio.proj.weight",
            "decode_block4.mlp.fc2.weight", "blocks.0.attn.qkv.weight", "decode_block2.attn.norm_k.weight"]
    return dict(B=B, seed=seed, gain=gain, alpha=cfg.MODEL.LOSS_ALPHA,
                logits=logits.detach(), v=v.detach(), a=a.detach(),
                loss=loss.detach(), kld=kld.detach(), nce=nce.detach(),
                grad_norms={n: g.norm().item() for n, g in grads.items()},
                grads={n: grads[n].clone() for n in keep}), shapes


def optional_paths(B=1, seed=7):
    """custom_multimodal_builder.py:425-440,448-451,483-491 and av_attention.py:356-370 on the unmodified reference."""
    rec = dict(B=B, seed=seed)
    video, audio, hm = O.synthetic_batch(B, seed=seed + 1)
    # (1) MVIT.SPATIAL_AUDIO_ATTN = True: audio attention re-weights the temporal-fusion input
    model, cfg = ref_shim.reference_model(seed=0, overrides=["MVIT.SPATIAL_AUDIO_ATTN", True])
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = O.synthetic_state(shapes, seed=seed, gain=2.0)
    model.load_state_dict(sd, strict=True)
    model.train()
    logits, v, a = model([video], audio, return_embed=True)
    p = torch.softmax(logits.reshape(B, 1, 8, -1) / 2, dim=-1).reshape(logits.shape)
    from slowfast.models import losses as ref_losses
    kld = ref_losses.get_loss_func("kldiv")()(p, hm)
    kld.backward()
    rec["saa_logits"], rec["saa_v"], rec["saa_kld"] = logits.detach(), v.detach(), kld.detach()
    rec["saa_grad_norms"] = {n: q.grad.norm().item() for n, q in model.named_parameters() if q.grad is not None}
    rec["saa_grads"] = {n: model.get_parameter(n).grad.clone() for n in ("spatial_fusion.attn.qkv.bias", "spatial_fusion.norm1.weight",
                                                                        "vision_pool.bias", "blocks.15.norm2.bias")}
    # (2) attention maps of the default model (visualisation outputs)
    model, cfg = ref_shim.reference_model(seed=0)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        out = model([video], audio, return_spatial_attn=True, return_temporal_attn=True)
    rec["attn_logits"], rec["temporal_attn"] = out[0], out[2]
    rec["spatial_attn_rows"] = out[1][:, :, ::13, :].clone()          # every 13th query row of (B, 8, 260, 260), all keys
    rec["spatial_attn_rowsum"] = out[1].sum(-1)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    ap.add_argument("--only-optional", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if args.only_optional:
        torch.save(optional_paths(), os.path.join(OUT, "optional_b1.pt"))
        return
    torch.save(unit_blocks(), os.path.join(OUT, "unit_blocks.pt"))
    torch.save(losses(), os.path.join(OUT, "losses.pt"))
    if not args.skip_full:
        rec, shapes = full()
        with open(os.path.join(OUT, "param_shapes.json"), "w") as f:
            json.dump(shapes, f, indent=0)
        torch.save(rec, os.path.join(OUT, "full_b2.pt"))
        rec4, _ = full(B=2, seed=5, gain=4.0, tag="full_b2_gain4")
        rec4.pop("grads")
        torch.save(rec4, os.path.join(OUT, "full_b2_gain4.pt"))
    torch.save(optional_paths(), os.path.join(OUT, "optional_b1.pt"))


if __name__ == "__main__":
    main()
