"""FusedClipAdamW (csts_grad_sqnorm + csts_clip_adamw_step) against the reference sequence it replaces:
scaler.unscale_ -> clip_grad_norm_ -> torch.optim.AdamW.step -> scaler.update (tools/train_avgaze_net.py:101-109)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"

SHAPES = [(768, 768), (96,), (3, 7, 11), (1,), (2304, 768), (96, 1, 3, 3, 3), (1, 4096, 96), (50001,), (16384 * 3 + 5,)]


def make_params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in SHAPES]


def set_grads(params, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        p.grad = (torch.randn(p.shape, generator=g) * 0.3 * scale).to(dev)


def groups_of(params):
    decay = [p for p in params if p.dim() > 1]
    no_decay = [p for p in params if p.dim() <= 1]
    return [{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}]


@pytest.mark.parametrize("max_norm", [1.0, 0.0, 1e6])
def test_clip_and_step_matches_torch(max_norm):
    from csts_b200.host.optimizer import FusedClipAdamW
    ref = make_params(0)
    ours = make_params(0)
    opt_ref = torch.optim.AdamW(groups_of(ref), lr=1e-3, eps=1e-8, weight_decay=0.05)
    opt = FusedClipAdamW(groups_of(ours), lr=1e-3, eps=1e-8, weight_decay=0.05)
    for step in range(4):
        set_grads(ref, 10 + step)
        set_grads(ours, 10 + step)
        lr = 1e-3 * (1 + step)                      # the schedule changes the rate every iteration (lr_policy.py)
        for o in (opt_ref, opt):
            for gsel in o.param_groups:
                gsel["lr"] = lr
        if max_norm:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt_ref.step()
        opt.clip_and_step(max_norm)
        if max_norm:
            assert abs(opt.grad_norm().item() - total.item()) <= 1e-5 * total.item()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (step, a.shape, (a - b).abs().max().item())
            sa, sb = opt.state[a], opt_ref.state[b]
            assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=2e-6, atol=1e-9)
            assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    assert opt._step.item() == 4.0
    sd = opt.state_dict()                           # same layout as torch's AdamW state (checkpoint interchange)
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and len(sd["param_groups"]) == 2


def test_grad_scaler_path_and_skipped_step():
    from csts_b200.host.optimizer import FusedClipAdamW
    ref, ours = make_params(1), make_params(1)
    opt_ref = torch.optim.AdamW(groups_of(ref), lr=2e-3, eps=1e-8, weight_decay=0.05)
    opt = FusedClipAdamW(groups_of(ours), lr=2e-3, eps=1e-8, weight_decay=0.05)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=2)
    scaler.scale(torch.ones((), device=dev))        # lazy init of the scale tensor
    for step in range(3):
        S = scaler.get_scale()
        set_grads(ref, 20 + step)
        set_grads(ours, 20 + step, scale=S)         # what back-propagating scaler.scale(loss) leaves in .grad
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt_ref.step()
        opt.clip_and_step(1.0, scaler)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=3e-6, atol=3e-7), (step, a.shape)
    assert scaler.get_scale() == 2048.0             # two clean steps -> growth (interval 2), third step counted from zero again
    # a non-finite gradient: the step is skipped, the scale backs off, the step count does not advance
    before = [p.detach().clone() for p in ours]
    set_grads(ours, 99, scale=scaler.get_scale())
    ours[3].grad[0] = float("inf")
    opt.clip_and_step(1.0, scaler)
    assert all(torch.equal(a, b) for a, b in zip(ours, before))
    assert scaler.get_scale() == 1024.0 and opt._step.item() == 3.0 and opt._found_inf.item() == 1.0


def test_weight_copies_are_refreshed_by_the_step():
    from csts_b200.host.optimizer import FusedClipAdamW
    from csts_b200.host.weights import FP16, WeightCache
    for prec, dt in ((None, torch.bfloat16), (FP16, torch.float16)):
        wc = WeightCache() if prec is None else WeightCache(prec)
        params = make_params(2)
        w0 = wc.w(params[0])                        # (768, 768) Linear weight: plain copy, refreshed by the kernel
        wpad = wc.w_padded(params[5], 32)           # conv weight: padded copy, must be rebuilt instead
        assert w0.dtype == dt
        opt = FusedClipAdamW(groups_of(params), lr=1e-2, eps=1e-8, weight_decay=0.05, weight_cache=wc)
        set_grads(params, 5)
        opt.clip_and_step(1.0)
        assert wc.w(params[0]) is w0 and torch.equal(w0, params[0].detach().to(dt))
        wpad2 = wc.w_padded(params[5], 32)
        assert wpad2 is not wpad and torch.equal(wpad2[:, :27], params[5].detach().reshape(96, 27).to(dt))
