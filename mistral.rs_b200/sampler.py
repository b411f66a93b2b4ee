"""Host tail of the on-device sampler (C++ host/sampler_tail.hpp): packed rows of the top-k / top-1 kernels -> token and
log-probability.  REF mistralrs-core/src/sampler.rs:1172-1273 (single row), :666-742 (batched rows), :1284-1297.
Pure host code; the random variate is the caller's."""
import ctypes

import numpy as np

from .kv_index import host_lib

_ERR = {-1: "invalid CUDA top-k packed output length", -2: "invalid CUDA top-k softmax normalizer",
        -3: "All sampling probabilities are zero after CUDA top-k filtering.",
        -4: "Invalid sampling probability. The model likely produced NaN/Inf logits.", -5: "invalid CUDA top-1 output"}


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sample_topk_packed_row(packed, packed_k, row_k, inv_temperature, top_p, min_p, u):
    """One row [values k | indices k | denom | max] -> (token, logprob)."""
    p = _f32(packed)
    tok, lp = ctypes.c_uint32(0), ctypes.c_float(0.0)
    rc = host_lib().mrs_sample_topk_packed_row(ctypes.c_void_p(p.ctypes.data), ctypes.c_int64(p.size), ctypes.c_int64(packed_k),
                                               ctypes.c_int64(row_k), ctypes.c_float(inv_temperature), ctypes.c_float(top_p),
                                               ctypes.c_float(min_p), ctypes.c_double(u), ctypes.byref(tok), ctypes.byref(lp))
    if rc != 0:
        raise ValueError(_ERR[rc])
    return tok.value, lp.value


def sample_topk_packed_batch(packed, packed_k, row_k, inv_temperature, top_p, min_p, u):
    """Rows [batch, 2k+2] with per-row parameters -> (tokens u32 [batch], logprobs f32 [batch], status i32 [batch])."""
    p = _f32(packed)
    batch = p.shape[0]
    assert p.shape == (batch, 2 * packed_k + 2)
    rk = np.ascontiguousarray(np.broadcast_to(row_k, (batch,)), dtype=np.int64)
    it, tp, mp = (_f32(np.broadcast_to(v, (batch,))) for v in (inv_temperature, top_p, min_p))
    uu = np.ascontiguousarray(np.broadcast_to(u, (batch,)), dtype=np.float64)
    tokens, logprobs, status = np.zeros(batch, np.uint32), np.zeros(batch, np.float32), np.zeros(batch, np.int32)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    host_lib().mrs_sample_topk_packed_batch(P(p), ctypes.c_int64(batch), ctypes.c_int64(packed_k), P(rk), P(it), P(tp), P(mp), P(uu),
                                            P(tokens), P(logprobs), P(status))
    return tokens, logprobs, status


def sample_top1_row(packed):
    p = _f32(packed)
    tok = ctypes.c_uint32(0)
    if p.size != 2 or host_lib().mrs_sample_top1_row(ctypes.c_void_p(p.ctypes.data), ctypes.byref(tok)) != 0:
        raise ValueError(_ERR[-5])
    return tok.value


CUDA_TOPK_MAX_K = 128      # REF mistralrs-core/src/ops.rs:18


def cuda_batch_sampling_plan(temperature, top_k, top_p, min_p, *, return_logprobs=False, frequency_penalty=None, presence_penalty=None,
                             repetition_penalty=None, dry_multiplier=None, has_logits_bias=False, has_logits_processors=False):
    """Which device kernel a request's row goes through, or None when the row must take the host path
    (REF sampler.rs:613-660 `cuda_batch_sampling_plan`): ("greedy", 1.0) — no temperature;
    ("topk", k, inv_temperature) — top_k in 1..=128; ("categorical", inv_temperature) — no top-k and neither nucleus nor
    min-p filtering.  Anything that edits the logits on the host (penalties, DRY, bias, processors) or asks for logprobs
    rules the batched device path out."""
    import math
    has_penalties = ((frequency_penalty or 0.0) != 0.0 or (presence_penalty or 0.0) != 0.0
                     or (1.0 if repetition_penalty is None else repetition_penalty) != 1.0)
    if return_logprobs or has_penalties or (dry_multiplier or 0.0) != 0.0 or has_logits_bias or has_logits_processors:
        return None
    if temperature is None:
        return ("greedy", 1.0)
    if not (math.isfinite(temperature) and temperature > 0.0):
        return None
    with np.errstate(over="ignore"):
        inv_t = float(np.float32(1.0 / temperature))      # the reference keeps it as f32: a tiny temperature overflows to inf -> no plan
    if not (math.isfinite(inv_t) and inv_t > 0.0):
        return None
    if top_k > 0:
        return ("topk", int(top_k), inv_t) if top_k <= CUDA_TOPK_MAX_K else None
    if not (0.0 < top_p < 1.0) and not (0.0 < min_p < 1.0):
        return ("categorical", inv_t)
    return None
