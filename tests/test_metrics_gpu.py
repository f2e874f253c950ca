"""GPU metric path: the fused adaptive-F1 kernel against the oracle restatement of
slowfast/utils/metrics.py:9-74 (itself pinned to the live reference in tests/test_oracle_golden.py), the fused min-max
rescale of the loops, and the host-sync-free step statistics."""
import pytest
import torch

import csts_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _case(B, seed, scale):
    g = torch.Generator().manual_seed(seed)
    _, _, hm = O.synthetic_batch(B, seed=seed + 3)
    logits = torch.randn(B, 1, 8, 64, 64, generator=g) + 6.0 * hm.unsqueeze(1) / hm.amax(dim=(-1, -2), keepdim=True).unsqueeze(1)
    probs = O.frame_softmax(logits, 2.0)
    labels = torch.rand(B, 8, 3, generator=g)
    labels[:, :, 2] = 0.0
    labels[0, 1, 2] = 1.0
    labels[B - 1, 5, 2] = 2.0
    return probs * scale, hm, labels


@pytest.mark.parametrize("dataset,scale", [("ego4d_av_gaze_forecast", 0.08), ("aria_av_gaze_forecast", 0.02), ("ego4d_av_gaze", 0.02)])
def test_adaptive_f1_kernel_matches_the_reference_formula(dataset, scale):
    from csts_b200.host import metrics
    probs, hm, labels = _case(3, 1, 1.0)
    resc = O.minmax_rescale(probs) * scale                     # inside the dataset's threshold grid
    want = O.adaptive_f1(resc, hm, labels, dataset)
    got = metrics.adaptive_f1(resc.to(dev), hm.to(dev), labels.to(dev), dataset)
    assert all(abs(a - b) <= 2e-6 for a, b in zip(got, want)), (got, want)
    assert 0.0 < want[0] < 1.0 and got[3] == want[3]


def test_adaptive_f1_fuses_the_min_max_rescale_of_the_loops():
    from csts_b200.host import metrics
    probs, hm, labels = _case(4, 7, 1.0)
    want = O.adaptive_f1(O.minmax_rescale(probs), hm, labels, "ego4d_av_gaze")      # tools/train_avgaze_net.py:125-128
    out = metrics.adaptive_f1_async(probs.to(dev), hm.to(dev), labels.to(dev), "ego4d_av_gaze", rescale=True).cpu()
    assert abs(out[0].item() - want[0]) <= 2e-6 and abs(out[1].item() - want[1]) <= 2e-6 and abs(out[2].item() - want[2]) <= 2e-6
    assert metrics.thresholds_for("ego4d_av_gaze")[int(out[4])] == want[3]


def test_step_stats_average_without_host_sync():
    from csts_b200.host.step_stats import AsyncStepStats
    probs, hm, labels = (t.to(dev) for t in _case(2, 11, 1.0))
    st = AsyncStepStats("ego4d_av_gaze_forecast", period=3)
    vals = [(1.0, 0.5, 10.0), (2.0, 1.5, 10.0), (6.0, 4.0, 40.0)]
    for i, (a, b, c) in enumerate(vals):
        st.update(torch.tensor(a, device=dev), torch.tensor(b, device=dev), torch.tensor(c, device=dev), probs, hm, labels)
        if i < 2:
            assert st.poll() is None
    rec = st.poll(wait=True)
    assert rec["iteration"] == 3 and rec["steps"] == 3 and not rec["nan"]
    assert abs(rec["loss"] - 3.0) < 1e-6 and abs(rec["kldiv_loss"] - 2.0) < 1e-6 and abs(rec["egonce_loss"] - 20.0) < 1e-5
    want = O.adaptive_f1(O.minmax_rescale(probs.cpu()), hm.cpu(), labels.cpu(), "ego4d_av_gaze_forecast")
    assert abs(rec["f1"] - want[0]) <= 2e-6 and rec["threshold"] == want[3]
    st.update(torch.tensor(float("nan"), device=dev))
    st.update(torch.tensor(1.0, device=dev))
    st.update(torch.tensor(1.0, device=dev))
    assert st.poll(wait=True)["nan"]
