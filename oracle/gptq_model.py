"""ORACLE (test infrastructure): CPU decode forward of the GPTQ int4 layer stack, composed from the
oracle pieces with every intermediate tensor rounded through the activation dtype where the
reference materialises one.  Linears: y = dtype(x . dtype((q - 8) * s)) (Marlin numerics,
oracle/gptq.py; REF gptq_cuda.rs:357-398, marlin_kernel.cuh:93-140); block structure REF
mistralrs-core/src/models/mistral.rs (same as llama.rs:243-260): residual adds through
add_rms_norm (REF core/src/cuda/sort.cu:701-727)."""
import numpy as np

import oracle
from oracle import gptq as og


class OracleGptq:
    def __init__(self, cfg, host, rope_cos, rope_sin, dt="f16"):
        self.cfg, self.hw, self.dt = cfg, host, dt
        self.cos, self.sin = oracle.round_dtype(rope_cos, dt), oracle.round_dtype(rope_sin, dt)
        self.k = [[] for _ in range(cfg.n_layers)]
        self.v = [[] for _ in range(cfg.n_layers)]
        self._deq = {}

    def _lin(self, l, name, x):
        key = (l, name)
        if key not in self._deq:
            qw, sc = self.hw[key]
            self._deq[key] = og.dequant_gptq(qw, sc, None, self.cfg.group_size)
        return oracle.round_dtype(og.gemm(x, self._deq[key]).astype(np.float32), self.dt)

    def step(self, tokens, pos):
        c, dt = self.cfg, self.dt
        B, D, H, KVH = len(tokens), c.head_dim, c.n_heads, c.n_kv_heads
        x = oracle.round_dtype(self.hw[(0, "tok_embd")][tokens], dt)
        h = oracle.rms_norm(x, self.hw[(0, "attn_norm")], c.rms_eps, dt)
        for l in range(c.n_layers):
            q, k, v = (self._lin(l, n, h) for n in ("q_proj", "k_proj", "v_proj"))
            q, k = oracle.rotary(q, k, self.cos, self.sin, np.full(B, pos, dtype=np.uint32), bool(c.rope_neox), D, D // 2, H, KVH, dt)
            self.k[l].append(k.copy()); self.v[l].append(v.copy())
            kk = np.stack(self.k[l], axis=1).reshape(B, -1, KVH, D).astype(np.float64)
            vv = np.stack(self.v[l], axis=1).reshape(B, -1, KVH, D).astype(np.float64)
            qq = q.reshape(B, H, D).astype(np.float64)
            out = np.empty((B, H, D), dtype=np.float64)
            g = H // KVH
            for hh in range(H):
                s = np.einsum("btd,bd->bt", kk[:, :, hh // g], qq[:, hh]) * (1.0 / np.sqrt(D))
                s = np.exp(s - s.max(axis=1, keepdims=True))
                out[:, hh] = np.einsum("bt,btd->bd", s / s.sum(axis=1, keepdims=True), vv[:, :, hh // g])
            attn = oracle.round_dtype(out.reshape(B, H * D).astype(np.float32), dt)
            o = self._lin(l, "o_proj", attn)
            x2, h2 = oracle.add_rms_norm(o, x, self.hw[(l, "ffn_norm")], c.rms_eps, dt)
            act = oracle.fused_glu(self._lin(l, "gate_proj", h2), self._lin(l, "up_proj", h2), 0, dt)
            d = self._lin(l, "down_proj", act)
            nxt = self.hw[(l + 1, "attn_norm")] if l + 1 < c.n_layers else self.hw[(0, "final_norm")]
            x, h = oracle.add_rms_norm(d, x2, nxt, c.rms_eps, dt)
        return oracle.round_dtype((h.astype(np.float64) @ self.hw[(0, "lm_head")].astype(np.float64).T).astype(np.float32), dt)
