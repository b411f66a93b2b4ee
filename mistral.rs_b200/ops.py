"""Host-side mirror of the small fused ops the reference calls between the GEMMs:
`fused_glu` (mistralrs-quant/src/utils/ops.rs, C: utils/ffi.rs:274-305), `apply_rotary_qk`
(mistralrs-quant/src/rotary/mod.rs:851 -> rotary/ffi.rs), `RmsNorm::forward` /
`forward_add_rms_norm` (mistralrs-core/src/layers.rs:328-413 -> core/src/cuda/ffi.rs)."""
import ctypes

import numpy as np

import torch

from . import lib

_TAG = {torch.float16: "f16", torch.bfloat16: "bf16", torch.float32: "f32"}
_DT_CODE = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def fused_glu(a, b, activation):
    """out = act(a) * b over the last dim (rows may be strided)."""
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("fused_glu: shape/dtype mismatch")
    cols = a.shape[-1]
    rows = a.numel() // cols
    a2, b2 = a.reshape(rows, cols), b.reshape(rows, cols)
    if a2.stride(1) != 1 or b2.stride(1) != 1:
        a2, b2 = a2.contiguous(), b2.contiguous()
    out = torch.empty(a.shape, dtype=a.dtype, device=a.device)
    getattr(lib(), f"fused_glu_{_TAG[a.dtype]}")(_p(a2), _p(b2), _p(out), ctypes.c_uint32(rows), ctypes.c_uint32(cols),
                                                  ctypes.c_uint32(a2.stride(0)), ctypes.c_uint32(b2.stride(0)),
                                                  ctypes.c_int(int(activation)), ctypes.c_void_p(_stream(a.device)))
    return out


def fused_split_glu(x, activation):
    """x [..., 2*split]: act(x[..., :split]) * x[..., split:]."""
    x = x.contiguous()
    split = x.shape[-1] // 2
    rows = x.numel() // (2 * split)
    out = torch.empty(*x.shape[:-1], split, dtype=x.dtype, device=x.device)
    getattr(lib(), f"fused_split_glu_{_TAG[x.dtype]}")(_p(x), _p(out), ctypes.c_uint32(rows), ctypes.c_uint32(split),
                                                        ctypes.c_int(int(activation)), ctypes.c_void_p(_stream(x.device)))
    return out


def apply_rotary_qk(q, k, cos, sin, positions=None, is_neox=True):
    """In-place RoPE on q [T, H, D] and k [T, KVH, D]; cos/sin [*, rot_half]."""
    if q.dtype != k.dtype or cos.dtype != q.dtype or sin.dtype != q.dtype:
        raise ValueError("apply_rotary_qk: dtype mismatch")
    T, H, D = q.shape
    KVH = k.shape[1]
    rot_half = cos.shape[-1]
    if positions is None:
        lib().rotary_embedding(_p(q), _p(k), _p(cos), _p(sin), ctypes.c_int(int(is_neox)), ctypes.c_int(D),
                               ctypes.c_int64(T), ctypes.c_int(rot_half), ctypes.c_int(H), ctypes.c_int(KVH),
                               ctypes.c_int64(q.stride(0)), ctypes.c_int64(k.stride(0)),
                               ctypes.c_uint32(_DT_CODE[q.dtype]), ctypes.c_int64(_stream(q.device)))
    else:
        if positions.dtype not in (torch.int32, torch.uint32):
            raise ValueError("positions must be u32/i32")
        lib().rotary_embedding_positions(_p(q), _p(k), _p(cos), _p(sin), _p(positions), ctypes.c_int(int(is_neox)),
                                         ctypes.c_int(D), ctypes.c_int64(T), ctypes.c_int(rot_half),
                                         ctypes.c_int(cos.shape[0]), ctypes.c_int(H), ctypes.c_int(KVH),
                                         ctypes.c_int64(q.stride(0)), ctypes.c_int64(k.stride(0)),
                                         ctypes.c_uint32(_DT_CODE[q.dtype]), ctypes.c_int64(_stream(q.device)))


def embedding_gather(w, ids: torch.Tensor, dtype=torch.bfloat16):
    """Rows `ids` of a ggml-block embedding table, dequantised to `dtype` — the
    `GgufMatMul::embedding_forward_raw` gather (REF gguf/mod.rs:790-845: dequantize().index_select()).
    w: quant.QTensor [vocab, cols]; ids: i32 tensor of any shape -> [..., cols]."""
    from . import GGML
    if ids.dtype != torch.int32:
        raise ValueError("embedding_gather expects i32 ids")
    flat = ids.reshape(-1).contiguous()
    out = torch.empty(flat.numel(), w.shape[1], dtype=dtype, device=w.data.device)
    rc = lib().mrs_embedding_gather(ctypes.c_int32(GGML[w.dtype]), _p(w.data), ctypes.c_int32(w.shape[1]), _p(flat),
                                    ctypes.c_int32(flat.numel()), _p(out),
                                    ctypes.c_int32({torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[dtype]),
                                    ctypes.c_void_p(_stream(w.data.device)))
    if rc != 0:
        raise RuntimeError(f"mrs_embedding_gather failed with cudaError {rc}")
    return out.reshape(*ids.shape, w.shape[1])


def rms_norm(x, weight, eps):
    x = x.contiguous()
    cols = x.shape[-1]
    out = torch.empty_like(x)
    getattr(lib(), f"mrs_rms_norm_{_TAG[x.dtype]}")(_p(x), _p(weight), _p(out), ctypes.c_int(x.numel() // cols),
                                                     ctypes.c_int(cols), ctypes.c_float(eps), ctypes.c_int64(_stream(x.device)))
    return out


def rms_norm_strided_4d(x, weight, eps):
    """Per-head RMSNorm of a (possibly strided) [B, H, S, D] view -> contiguous [B, H, S, D]
    (`rms_norm_strided_4d_*`, core/src/cuda/ffi.rs:183; the QK-norm fast path of layers.rs)."""
    if x.dim() != 4 or weight.shape[-1] != x.shape[-1]:
        raise ValueError("rms_norm_strided_4d expects [batch, heads, seq, head_dim] and a [head_dim] weight")
    b, h, s, d = x.shape
    out = torch.empty(b, h, s, d, dtype=x.dtype, device=x.device)
    sb, sh, ss, sd = x.stride()
    getattr(lib(), f"rms_norm_strided_4d_{_TAG[x.dtype]}")(
        _p(x), _p(weight.contiguous()), _p(out), ctypes.c_int64(sb), ctypes.c_int64(sh), ctypes.c_int64(ss),
        ctypes.c_int64(sd), ctypes.c_int(b), ctypes.c_int(h), ctypes.c_int(s), ctypes.c_int(d), ctypes.c_float(eps),
        ctypes.c_int64(_stream(x.device)))
    return out


def add_rms_norm(x, residual, weight, eps):
    """(sum, normed) = (x + residual, rmsnorm(x + residual)) — layers.rs:328 forward_add_rms_norm."""
    x, residual = x.contiguous(), residual.contiguous()
    cols = x.shape[-1]
    s, n = torch.empty_like(x), torch.empty_like(x)
    getattr(lib(), f"add_rms_norm_{_TAG[x.dtype]}")(_p(x), _p(residual), _p(weight), _p(s), _p(n),
                                                     ctypes.c_int(x.numel() // cols), ctypes.c_int(cols),
                                                     ctypes.c_float(eps), ctypes.c_int64(_stream(x.device)))
    return s, n


CUDA_TOPK_CHUNK_SIZE, CUDA_TOPK_MAX_K = 2048, 128   # REF mistralrs-core/src/ops.rs:12,18


def cuda_topk_logits_f32_packed(logits: torch.Tensor, k: int, temperature: float):
    """Mirror of `cuda_topk_logits_f32_packed[_batched]` (REF mistralrs-core/src/ops.rs:690-830): f32 logits
    [vocab] or [rows, vocab] -> (values [.., k], indices [.., k] i64, denom [..], global_max [..])."""
    if temperature <= 0.0 or not np.isfinite(temperature):
        raise ValueError("cuda_topk_logits_f32_packed requires a positive finite temperature")
    if logits.dtype != torch.float32:
        raise ValueError("cuda_topk_logits_f32_packed requires F32 logits")
    x = logits.contiguous().reshape(-1, logits.shape[-1])
    rows, ncols = x.shape
    k = min(k, ncols)
    if k == 0 or k > CUDA_TOPK_MAX_K:
        raise ValueError(f"cuda_topk_logits_f32_packed k={k} must be in [1, {CUDA_TOPK_MAX_K}]")
    nblocks = -(-ncols // CUDA_TOPK_CHUNK_SIZE)
    dev = x.device
    bv = torch.empty(rows * nblocks * k, dtype=torch.float32, device=dev)
    bi = torch.empty(rows * nblocks * k, dtype=torch.int32, device=dev)
    bm = torch.empty(rows * nblocks, dtype=torch.float32, device=dev)
    bs = torch.empty(rows * nblocks, dtype=torch.float32, device=dev)
    packed = torch.empty(rows, 2 * k + 2, dtype=torch.float32, device=dev)
    it = torch.full((rows,), 1.0 / temperature, dtype=torch.float32, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    lib().topk_large_f32_packed_batched(P(x), P(it), P(bv), P(bi), P(bm), P(bs), P(packed), ctypes.c_int(rows), ctypes.c_int(ncols),
                                        ctypes.c_int(k), ctypes.c_int(CUDA_TOPK_CHUNK_SIZE), ctypes.c_int(nblocks),
                                        ctypes.c_int64(torch.cuda.current_stream(dev).cuda_stream))
    shape = logits.shape[:-1]
    return (packed[:, :k].reshape(*shape, k), packed[:, k:2 * k].to(torch.int64).reshape(*shape, k),
            packed[:, 2 * k].reshape(shape), packed[:, 2 * k + 1].reshape(shape))


def cuda_top1_logits_f32(logits: torch.Tensor) -> torch.Tensor:
    """greedy token ids [rows] (u32 as i64) of f32 logits [rows, vocab] — `top1_large_f32_packed_batched`."""
    x = logits.contiguous().reshape(-1, logits.shape[-1])
    rows, ncols = x.shape
    nblocks = -(-ncols // CUDA_TOPK_CHUNK_SIZE)
    dev = x.device
    bv = torch.empty(rows * nblocks, dtype=torch.float32, device=dev)
    bi = torch.empty(rows * nblocks, dtype=torch.int32, device=dev)
    out = torch.empty(rows, dtype=torch.int32, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    lib().top1_large_f32_packed_batched(P(x), P(bv), P(bi), ctypes.c_void_p(0), P(out), ctypes.c_int(rows), ctypes.c_int(ncols),
                                        ctypes.c_int(CUDA_TOPK_CHUNK_SIZE), ctypes.c_int(nblocks),
                                        ctypes.c_int64(torch.cuda.current_stream(dev).cuda_stream))
    return out.to(torch.int64) & 0xFFFFFFFF
