"""Host tail of the on-device sampler (C++ host/sampler_tail.hpp) vs the numpy restatement of the reference's
`sample_topk_on_device` / `sample_cuda_topk_packed_row` (sampler.rs:1172-1273, 666-742), plus the filter semantics the
reference's own sampler tests pin (nucleus mass on the kept set, min-p without top-p, greedy rows).  Token choice:
exact; log-probabilities: 1e-6 relative (libm expf vs numpy)."""
import numpy as np
import pytest

from mistralrs_b200 import sampler
from oracle import sampler_np as osn


def _row(seed, vocab=4000, k=40, inv_t=1.0 / 0.7, scale=3.0):
    rng = np.random.default_rng(seed)
    return osn.pack_row((rng.standard_normal(vocab) * scale).astype(np.float32), k, inv_t)


@pytest.mark.parametrize("top_p,min_p", [(1.0, 0.0), (0.9, 0.0), (0.5, 0.0), (1.0, 0.1), (0.8, 0.05), (0.0, 0.0)])
def test_rows_vs_oracle(top_p, min_p):
    k, inv_t = 40, 1.0 / 0.7
    rng = np.random.default_rng(3)
    for seed in range(60):
        packed = _row(seed, k=k, inv_t=inv_t)
        u = float(rng.random())
        tok, lp = sampler.sample_topk_packed_row(packed, k, k, inv_t, top_p, min_p, u)
        etok, elp, _, _ = osn.sample_row(packed, k, k, inv_t, top_p, min_p, u)
        assert tok == etok
        assert abs(lp - elp) <= 1e-6 * max(1.0, abs(elp))


def test_draw_follows_the_filtered_distribution():
    k, inv_t = 8, 1.0
    packed = osn.pack_row(np.log(np.array([0.4, 0.25, 0.15, 0.1, 0.05, 0.03, 0.01, 0.01], dtype=np.float64)).astype(np.float32), k, inv_t)
    # nucleus 0.7 on the kept mass: 0.4 + 0.25 = 0.65 < 0.7 keeps the third entry too, then everything else goes
    _, _, report, w = osn.sample_row(packed, k, k, inv_t, 0.7, 0.0, 0.0)
    assert np.count_nonzero(w) == 3 and np.allclose(report, [0.4, 0.25, 0.15, 0.1, 0.05, 0.03, 0.01, 0.01], rtol=1e-5)
    us = (np.arange(20000) + 0.5) / 20000
    toks = np.array([sampler.sample_topk_packed_row(packed, k, k, inv_t, 0.7, 0.0, float(u))[0] for u in us])
    counts = np.bincount(toks, minlength=8) / us.size
    assert np.allclose(counts[:3], np.array([0.4, 0.25, 0.15]) / 0.8, atol=2e-4) and counts[3:].sum() == 0
    # the log-probability reported is of the UNFILTERED distribution
    assert abs(sampler.sample_topk_packed_row(packed, k, k, inv_t, 0.7, 0.0, 0.0)[1] - np.log(0.4)) < 1e-5
    # u just below 1 lands on the last surviving entry, never on a filtered one
    assert sampler.sample_topk_packed_row(packed, k, k, inv_t, 0.7, 0.0, 1.0 - 2 ** -53)[0] == 2


def test_min_p_applies_without_top_p():   # sampler.rs `test_min_p_applies_without_top_p`
    k = 4
    packed = osn.pack_row(np.log(np.array([0.6, 0.3, 0.07, 0.03])).astype(np.float32), k, 1.0)
    toks = {sampler.sample_topk_packed_row(packed, k, k, 1.0, 1.0, 0.2, float(u))[0] for u in np.linspace(0, 0.999999, 500)}
    assert toks == {0, 1}                                  # 0.07 and 0.03 are <= 0.2 * 0.6
    toks = {sampler.sample_topk_packed_row(packed, k, k, 1.0, 1.0, 0.5, float(u))[0] for u in np.linspace(0, 0.999999, 500)}
    assert toks == {0}                                     # the threshold test is `threshold >= p`: 0.3 == 0.5 * 0.6 goes too


def test_greedy_and_partial_rows():
    k = 16
    packed = _row(11, k=k)
    first = int(packed[k])
    assert sampler.sample_topk_packed_row(packed, k, 1, 1.0 / 0.7, 0.9, 0.1, 0.99)[0] == first     # row_k 1 = greedy
    # a row asking for fewer entries than the batch was packed with only sees its own k (sampler.rs:684-690)
    for u in np.linspace(0, 0.9999, 50):
        tok, _ = sampler.sample_topk_packed_row(packed, k, 4, 1.0 / 0.7, 1.0, 0.0, float(u))
        assert tok in set(int(v) for v in packed[k:k + 4])
    assert sampler.sample_topk_packed_row(packed, k, 99, 1.0 / 0.7, 1.0, 0.0, 0.0)[0] == first    # k is clamped to what was packed


def test_batch_matches_rows_and_reports_per_row_status():
    k = 24
    rows = np.stack([_row(100 + i, k=k, inv_t=1.0 / t) for i, t in enumerate((0.7, 1.0, 1.3, 0.5))])
    rows[2, 2 * k] = 0.0                                   # a broken normaliser in one row must not poison the others
    rk, it = [k, 8, k, 1], [1 / 0.7, 1.0, 1 / 1.3, 2.0]
    tp, mp, u = [0.9, 1.0, 0.9, 0.5], [0.0, 0.05, 0.0, 0.0], [0.3, 0.6, 0.1, 0.9]
    toks, lps, status = sampler.sample_topk_packed_batch(rows, k, rk, it, tp, mp, u)
    assert list(status) == [0, 0, -2, 0]
    for b in (0, 1, 3):
        tok, lp = sampler.sample_topk_packed_row(rows[b], k, rk[b], it[b], tp[b], mp[b], u[b])
        assert toks[b] == tok and lps[b] == np.float32(lp)


def test_error_rows():
    k = 8
    packed = _row(5, k=k)
    with pytest.raises(ValueError, match="length"):
        sampler.sample_topk_packed_row(packed[:-1], k, k, 1.0, 1.0, 0.0, 0.5)
    bad = packed.copy(); bad[2 * k] = np.inf
    with pytest.raises(ValueError, match="normalizer"):
        sampler.sample_topk_packed_row(bad, k, k, 1.0, 1.0, 0.0, 0.5)
    bad = packed.copy(); bad[2 * k + 1] = np.nan
    with pytest.raises(ValueError, match="normalizer"):
        sampler.sample_topk_packed_row(bad, k, k, 1.0, 1.0, 0.0, 0.5)
    bad = packed.copy(); bad[1] = np.nan
    with pytest.raises(ValueError, match="NaN/Inf"):
        sampler.sample_topk_packed_row(bad, k, k, 1.0, 1.0, 0.0, 0.5)
    under = packed.copy(); under[:k] = -1e30                # every probability underflows to zero
    with pytest.raises(ValueError, match="zero"):
        sampler.sample_topk_packed_row(under, k, k, 1.0, 1.0, 0.0, 0.5)


def test_top1_row():   # sampler.rs:1284-1297
    assert sampler.sample_top1_row([3.5, 1234.0]) == 1234
    for bad in ([np.nan, 1.0], [1.0, np.inf], [1.0, -1.0], [1.0, 2.5]):
        with pytest.raises(ValueError):
            sampler.sample_top1_row(bad)


def test_cuda_batch_sampling_plan():   # sampler.rs:613-660
    plan = sampler.cuda_batch_sampling_plan
    assert plan(None, 40, 0.9, 0.0) == ("greedy", 1.0)
    assert plan(0.5, 40, 0.9, 0.05) == ("topk", 40, 2.0)
    assert plan(0.5, 128, 1.0, 0.0) == ("topk", 128, 2.0) and plan(0.5, 129, 1.0, 0.0) is None
    assert plan(2.0, -1, 1.0, 0.0) == ("categorical", 0.5) and plan(2.0, 0, 0.0, 1.0) == ("categorical", 0.5)
    assert plan(2.0, -1, 0.9, 0.0) is None and plan(2.0, -1, 1.0, 0.1) is None       # nucleus / min-p need the sorted head
    for bad_t in (0.0, -1.0, float("inf"), float("nan"), 1e-45):
        assert plan(bad_t, 40, 1.0, 0.0) is None
    assert plan(0.7, 40, 1.0, 0.0, return_logprobs=True) is None
    assert plan(0.7, 40, 1.0, 0.0, frequency_penalty=0.1) is None and plan(0.7, 40, 1.0, 0.0, presence_penalty=-0.2) is None
    assert plan(0.7, 40, 1.0, 0.0, repetition_penalty=1.1) is None and plan(0.7, 40, 1.0, 0.0, repetition_penalty=1.0) is not None
    assert plan(0.7, 40, 1.0, 0.0, dry_multiplier=0.8) is None and plan(0.7, 40, 1.0, 0.0, dry_multiplier=0.0) is not None
    assert plan(0.7, 40, 1.0, 0.0, has_logits_bias=True) is None and plan(None, 1, 1.0, 0.0, has_logits_processors=True) is None
