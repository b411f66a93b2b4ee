"""ORACLE (test infrastructure): CPU decode forward of the Llama-family layer stack with the
reference's GPU decode numerics, composed from the C oracle functions (every intermediate tensor
is rounded through the activation dtype exactly where the reference materialises a tensor):

  x = embed[token]                                   (dequantised row -> dtype)
  per layer (REF mistralrs-core/src/models/llama.rs:243-260, :68-135):
     h   = rms_norm(x)                               -> dtype
     q,k,v = W{q,k,v} . q8_1(h)                       (mmvq arithmetic) -> dtype
     rope(q, k) in dtype; cache write; paged attention (f64 softmax) -> dtype
     x2  = dtype(dtype(Wo . q8_1(attn)) + x)
     g   = dtype(act(dtype(Wg . q8_1(norm(x2)))) * dtype(Wu . ...))
     x   = dtype(dtype(Wd . q8_1(g)) + x2)
  logits = dtype(Wout . q8_1(rms_norm(x)))

`cpu_path=True` swaps the GEMVs for the restated candle CPU path (Q8_K / Q8_0 activations,
mrs_qmatmul_cpu) — that variant is the timed CPU baseline, not the parity oracle.
"""
import numpy as np

import oracle


class OracleLlama:
    def __init__(self, cfg, host_weights, tensor_type, rope_cos, rope_sin, dt="bf16", cpu_path=False, threads=1,
                 exact_gemm=False):
        """exact_gemm: the linears are the exact product of the dequantised weights with the (dtype-
        rounded) activations — the known-answer form the prefill dequant-GEMM is tested against —
        instead of the decode path's Q8_1 integer dots.  Stepping token by token then restates a
        prefill forward (causal attention == attention over the cache so far)."""
        self.cfg, self.hw, self.tt, self.dt = cfg, host_weights, tensor_type, dt
        self.exact_gemm = exact_gemm
        self.cos, self.sin = oracle.round_dtype(rope_cos, dt), oracle.round_dtype(rope_sin, dt)
        self.cpu_path, self.threads = cpu_path, threads
        self.k = [[] for _ in range(cfg.n_layers)]  # per layer list of [kvh*D] rows (dense cache)
        self.v = [[] for _ in range(cfg.n_layers)]

    def _gemv(self, layer, name, x, rows, cols):
        ty = self.tt(self.cfg, name, layer)
        w = self.hw[(layer, name)]
        if self.exact_gemm:
            y = oracle.matmul_exact(ty, w, x, cols, rows)
        elif self.cpu_path:
            y = oracle.qmatmul_cpu(ty, w, x, cols, rows, self.threads)
        else:
            xq, stride = oracle.quantize_q8_1(x)
            if self.threads > 1 and rows >= 4 * self.threads:
                # row slices on a thread pool (ctypes drops the GIL): same arithmetic, row by row
                from concurrent.futures import ThreadPoolExecutor
                rb = oracle.BLOCK_BYTES[ty] * cols // oracle.BLOCK_ELEMS[ty]
                wv = np.asarray(w).reshape(-1)
                cuts = [rows * i // self.threads for i in range(self.threads + 1)]
                def part(i):
                    r0, r1 = cuts[i], cuts[i + 1]
                    return oracle.mmvq_q8_1(ty, wv[r0 * rb:r1 * rb], xq, cols, r1 - r0, stride, x.shape[0])
                with ThreadPoolExecutor(self.threads) as ex:
                    y = np.concatenate(list(ex.map(part, range(self.threads))), axis=1).astype(np.float32)
            else:
                y = oracle.mmvq_q8_1(ty, w, xq, cols, rows, stride, x.shape[0]).astype(np.float32)
        return oracle.round_dtype(y.astype(np.float32), self.dt)

    def embed(self, tokens):
        c = self.cfg
        ty = self.tt(c, "token_embd", 0)
        be, bb = oracle.BLOCK_ELEMS[ty], oracle.BLOCK_BYTES[ty]
        w = self.hw[(0, "token_embd")].reshape(c.vocab, c.hidden // be * bb)
        rows = np.stack([oracle.dequantize(ty, w[t]) for t in tokens])
        return oracle.round_dtype(rows, self.dt)

    def step(self, tokens, pos):
        """tokens: list[int] (batch of independent sequences all at position `pos`); returns logits [B, vocab]."""
        c, dt = self.cfg, self.dt
        B = len(tokens)
        D, H, KVH = c.head_dim, c.n_heads, c.n_kv_heads
        x = self.embed(tokens)
        for l in range(c.n_layers):
            h = oracle.rms_norm(x, self.hw[(l, "attn_norm")], c.rms_eps, dt)
            q = self._gemv(l, "attn_q", h, H * D, c.hidden)
            k = self._gemv(l, "attn_k", h, KVH * D, c.hidden)
            v = self._gemv(l, "attn_v", h, KVH * D, c.hidden)
            q, k = oracle.rotary(q, k, self.cos, self.sin, np.full(B, pos, dtype=np.uint32), bool(getattr(c, "rope_neox", True)), D, D // 2, H, KVH, dt)
            self.k[l].append(k.copy()); self.v[l].append(v.copy())
            kk = np.stack(self.k[l], axis=1).reshape(B, -1, KVH, D).astype(np.float64)   # [B, T, KVH, D]
            vv = np.stack(self.v[l], axis=1).reshape(B, -1, KVH, D).astype(np.float64)
            qq = q.reshape(B, H, D).astype(np.float64)
            out = np.empty((B, H, D), dtype=np.float64)
            g = H // KVH
            for hh in range(H):
                s = np.einsum("btd,bd->bt", kk[:, :, hh // g], qq[:, hh]) * (1.0 / np.sqrt(D))
                s = np.exp(s - s.max(axis=1, keepdims=True))
                p = s / s.sum(axis=1, keepdims=True)
                out[:, hh] = np.einsum("bt,btd->bd", p, vv[:, :, hh // g])
            attn = oracle.round_dtype(out.reshape(B, H * D).astype(np.float32), dt)
            o = self._gemv(l, "attn_output", attn, c.hidden, H * D)
            x2 = oracle.round_dtype(o + x, dt)
            h2 = oracle.rms_norm(x2, self.hw[(l, "ffn_norm")], c.rms_eps, dt)
            gate = self._gemv(l, "ffn_gate", h2, c.inter, c.hidden)
            up = self._gemv(l, "ffn_up", h2, c.inter, c.hidden)
            act = oracle.fused_glu(gate, up, 0, dt)
            d = self._gemv(l, "ffn_down", act, c.hidden, c.inter)
            x = oracle.round_dtype(d + x2, dt)
        hf = oracle.rms_norm(x, self.hw[(0, "output_norm")], c.rms_eps, dt)
        return self._gemv(0, "output", hf, c.vocab, c.hidden)
