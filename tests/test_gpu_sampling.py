"""Sampling / perplexity tail: `ekv_sample_top_p` / `ekv_token_nll` through the C ABI against the
reference's own outputs (tests/golden/sampling_tail.npz), the CPU restatement, and torch.multinomial's draw.

Floating point: the kernel's block reductions sum in a different order than ATen's, so probabilities agree to fp32
rounding (rtol 2e-5 here) and the nucleus boundary may move by one token only when the exclusive cumulative mass is
within 2e-5 of top_p (`restate.top_p_margin`; ATen's CPU cumsum is a sequential fp32 sum, ~sqrt(n) ulps of noise).  Which members of a run of EQUAL probabilities survive a cut is index
order here and unspecified in the reference (unstable sort)."""
import pytest
import torch

from oracle import replay, restate

pytestmark = pytest.mark.gpu
RTOL = 2e-5


def _check_prob(got, ref, logits, temperature, top_p):
    got, ref = got.cpu(), ref.cpu()
    prob = torch.softmax(logits.cpu() / temperature, -1)
    margin = restate.top_p_margin(logits.cpu(), temperature, top_p)
    for r in range(got.shape[0]):
        differ = (got[r] > 0) != (ref[r] > 0)
        same = ~differ
        if margin[r] > 2e-5 or top_p == 0.0:                           # the boundary is not decided by rounding
            assert int((got[r] > 0).sum()) == int((ref[r] > 0).sum()), (r, int((got[r] > 0).sum()), int((ref[r] > 0).sum()))
            assert torch.allclose(got[r].sort().values, ref[r].sort().values, rtol=RTOL, atol=1e-12)
            assert prob[r][differ].unique().numel() <= 1               # only inside one run of equal probabilities
            assert torch.allclose(got[r][same], ref[r][same], rtol=RTOL, atol=1e-12)
        else:                                                          # (top_p = 1.0 always lands here: the tail's
            assert float(prob[r][differ].sum()) < 2e-4                #  cumulative mass rounds to either side of 1)
            assert torch.allclose(got[r][same], ref[r][same], rtol=5e-4, atol=1e-12)
        assert abs(float(got[r].sum()) - 1.0) < 1e-5


def test_sampling_tail_matches_reference_vectors(ekv_lib):
    from easykv_b200 import sampling
    for c in replay.load_sampling_tail():
        x = c["logits"].cuda()
        prob, raw = sampling.logits_adapter(x, c["temperature"], c["top_p"], arith=0)
        _check_prob(prob, c["final"], c["logits"], c["temperature"], c["top_p"])
        assert torch.allclose(raw.cpu(), c["raw"], rtol=RTOL, atol=1e-12)
        nll = sampling.token_nll(x, c["targets"].cuda())
        assert torch.allclose(nll.cpu(), c["nll"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("V,R,T,top_p", [(32000, 8, 0.8, 0.9), (32000, 3, 1.0, 0.95), (152064, 2, 0.7, 0.8),
                                          (1000, 64, 1.5, 0.5), (50257, 4, 1.0, 1.0), (17, 5, 0.9, 0.6)])
def test_sampling_tail_random_vs_restatement(ekv_lib, V, R, T, top_p):
    """Llama (32000), Qwen (152064: the global-memory workspace path), GPT-2 sized vocabularies; both ATen flavours."""
    from easykv_b200 import sampling
    x = torch.randn(R, V, generator=torch.Generator().manual_seed(V + R)) * 2.5
    for arith in (0, 1):
        prob, raw = sampling.logits_adapter(x.cuda(), T, top_p, arith=arith)
        ref, ref_raw = restate.logits_adapter(x.clone(), T, top_p, scale_mul=bool(arith))
        _check_prob(prob, ref, x, T, top_p)
        assert torch.allclose(raw.cpu(), ref_raw, rtol=RTOL, atol=1e-12)


def test_token_draw_is_torch_multinomial(ekv_lib):
    """Same generator state -> the token torch.multinomial(prob, 1) returns (ATen: argmax(prob / Exp(1)))."""
    from easykv_b200 import sampling
    x = (torch.randn(16, 32000, generator=torch.Generator().manual_seed(5)) * 2).cuda()
    hits = 0
    for seed in range(8):
        torch.manual_seed(seed)
        tok, prob, _ = sampling.sample_top_p(x, 0.9, 0.9, want_prob=True)
        torch.manual_seed(seed)
        ref = torch.multinomial(prob, num_samples=1)
        assert tok.shape == ref.shape and tok.dtype == ref.dtype
        hits += int((tok == ref).sum())
        assert bool((prob.gather(-1, tok) > 0).all())
        # the draw without the prob output (the decode loop's call) is the same token
        torch.manual_seed(seed)
        tok2 = sampling.sample_top_p(x, 0.9, 0.9)[0]
        assert torch.equal(tok, tok2)
    assert hits == 8 * 16
    # distribution sanity on a small vocabulary: empirical frequencies follow prob
    y = torch.tensor([[2.0, 1.0, 0.5, 0.0, -1.0, -3.0]]).cuda().repeat(4096, 1)
    torch.manual_seed(0)
    tok, prob, _ = sampling.sample_top_p(y, 1.0, 0.9, want_prob=True)
    freq = torch.bincount(tok[:, 0], minlength=6).float() / 4096
    assert torch.allclose(freq, prob[0], atol=0.03)
    assert float(freq[prob[0] == 0].sum()) == 0.0


def test_greedy_temperature_and_errors(ekv_lib):
    from easykv_b200 import sampling
    x = torch.randn(4, 32000, generator=torch.Generator().manual_seed(9)).cuda()
    tok = sampling.sample_top_p(x, 1e-9, 1.0)[0]
    assert torch.equal(tok[:, 0], x.argmax(-1))
    with pytest.raises(ValueError):
        sampling.sample_top_p(x, 0.0, 0.9)
    with pytest.raises(ValueError):
        sampling.sample_top_p(x, 1.0, -0.1)
    with pytest.raises(ValueError):
        sampling.token_nll(x, torch.zeros(3, dtype=torch.int64).cuda())
    with pytest.raises(RuntimeError):
        sampling.sample_top_p(x.cpu(), 1.0, 0.9)


def test_cut_through_equal_probabilities_keeps_lowest_indices(ekv_lib):
    """The slow path: the nucleus boundary falls inside a run of equal probabilities."""
    from easykv_b200 import sampling
    x = torch.full((2, 5000), -20.0)
    x[0, 100:110] = 3.0                        # ten equal tokens holding ~all the mass: top_p 0.55 keeps the first six
    x[1, ::500] = 3.0                          # ten equal tokens spread over the row (several 1024-blocks)
    prob, _ = sampling.logits_adapter(x.cuda(), 1.0, 0.55)
    ref, _ = restate.logits_adapter(x.clone(), 1.0, 0.55)
    assert torch.equal((prob > 0).cpu(), ref > 0)
    assert (prob[0] > 0).nonzero()[:, 0].tolist() == list(range(100, 106))
    assert (prob[1] > 0).nonzero()[:, 0].tolist() == list(range(0, 3000, 500))
    assert torch.allclose(prob.cpu(), ref, rtol=RTOL)
