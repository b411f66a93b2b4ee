"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's `paged_attention_v1` softmax variants —
ALiBi, logit soft-capping, attention sinks — REF mistralrs-paged-attn/src/cuda/pagedattention.cuh:270-345:

    qk = scale * <q, k>                                        (:273)
    qk = tanh(qk / softcapping) * softcapping   if softcapping != 1   (:277-279)
    qk += slope * (token_idx - context_len + 1)  if slope != 0  (:282-283; `context_len` is uint32_t there, so the
          difference WRAPS and the wrapped value is what gets converted to float — restated as is)
    m = max(qk_t, sink)                                          (:320-322, v1 only)
    p_t = exp(qk_t - m);  denom = sum p_t + exp(sink - m) + 1e-6 (:326-340)
    out = sum_t (p_t / denom) * v_t, rounded to the activation dtype

Works on token rows + a slot map (the cache layout does not matter to the arithmetic).  Pinned against outputs of the
reference kernel itself in tests/test_oracle_golden.py (tests/golden/ref_golden.npz: pa_out_v1_{alibi,softcap,sinks})."""
import numpy as np

from . import round_dtype


def paged_attention_v1(q, k_rows, v_rows, slots, block_tables, context_lens, num_kv_heads, head_size, block_size, scale,
                       dt="bf16", softcapping=1.0, alibi_slopes=None, sinks=None):
    S, H, D = q.shape
    group = H // num_kv_heads
    row_of = {int(s): i for i, s in enumerate(slots)}
    out = np.zeros((S, H, D), dtype=np.float32)
    k_rows = k_rows.reshape(len(slots), num_kv_heads, D).astype(np.float32)
    v_rows = v_rows.reshape(len(slots), num_kv_heads, D).astype(np.float32)
    for s in range(S):
        ctx = int(context_lens[s])
        rows = [row_of[int(block_tables[s][t // block_size]) * block_size + t % block_size] for t in range(ctx)]
        for h in range(H):
            kvh = h // group
            kk, vv = k_rows[rows, kvh], v_rows[rows, kvh]
            qk = np.float32(scale) * (kk @ q[s, h].astype(np.float32))
            if softcapping != 1.0:
                qk = np.tanh(qk / np.float32(softcapping)) * np.float32(softcapping)
            if alibi_slopes is not None and alibi_slopes[h] != 0:
                t = np.arange(ctx, dtype=np.int64)
                wrapped = ((t - ctx + 1) % (1 << 32)).astype(np.float64)      # uint32 arithmetic, then -> float
                qk = qk + np.float32(alibi_slopes[h]) * wrapped.astype(np.float32)
            m = qk.max()
            if sinks is not None:
                m = max(m, np.float32(sinks[h]))
            p = np.exp(qk - m)
            denom = p.sum() + (np.exp(np.float32(sinks[h]) - m) if sinks is not None else 0.0) + 1e-6
            out[s, h] = (p / denom) @ vv
    return round_dtype(out, dt)
