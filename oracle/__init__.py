"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes front-end of oracle/liboracle.so (oracle/mrs_oracle.c: a plain-C restatement of the
reference's arithmetic, every function citing the reference file:line it follows) and of the
unmodified reference CUDA kernels built from /root/reference into oracle/_ref/ (build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package, and only as the checker or the CPU timing baseline.  The product package
(mistral.rs_b200/) never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

GGML = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8, "q8_1": 9,
        "q2_k": 10, "q3_k": 11, "q4_k": 12, "q5_k": 13, "q6_k": 14}
BLOCK_ELEMS = {"q4_0": 32, "q4_1": 32, "q5_0": 32, "q5_1": 32, "q8_0": 32, "q8_1": 32,
               "q2_k": 256, "q3_k": 256, "q4_k": 256, "q5_k": 256, "q6_k": 256}
BLOCK_BYTES = {"q4_0": 18, "q4_1": 20, "q5_0": 22, "q5_1": 24, "q8_0": 34, "q8_1": 36,
               "q2_k": 84, "q3_k": 110, "q4_k": 144, "q5_k": 176, "q6_k": 210}
# byte offsets of the f16 fields inside a block (used to synthesise finite scales)
F16_FIELDS = {"q4_0": [0], "q4_1": [0, 2], "q5_0": [0], "q5_1": [0, 2], "q8_0": [0],
              "q2_k": [80, 82], "q3_k": [108], "q4_k": [0, 2], "q5_k": [0, 2], "q6_k": [208]}
DT = {"f16": 0, "bf16": 1, "f32": 2}

_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("mrs_oracle.c", "mrs_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.mrs_f16_to_f32.restype = ctypes.c_float
        L.mrs_f16_to_f32.argtypes = [ctypes.c_uint16]
        L.mrs_glu_act.restype = ctypes.c_float
        L.mrs_glu_act.argtypes = [ctypes.c_float, ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def random_blocks(dtype, nblocks, rng, scale_exp=(-9, -7)):
    """Synthetic ggml blocks per SURVEY §8(d): quants uniform over their full bit range, f16
    scales 2^U(lo,hi) (dmin = d*U(0,0.5)); K-quant 6-bit scales/mins uniform."""
    bb = BLOCK_BYTES[dtype]
    raw = rng.integers(0, 256, size=(nblocks, bb), dtype=np.uint8)
    fields = F16_FIELDS[dtype]
    d = np.exp2(rng.uniform(scale_exp[0], scale_exp[1], size=nblocks)).astype(np.float16)
    raw[:, fields[0]:fields[0] + 2] = d.view(np.uint8).reshape(nblocks, 2)
    if len(fields) > 1:
        m = (d.astype(np.float32) * rng.uniform(0, 0.5, size=nblocks)).astype(np.float16)
        raw[:, fields[1]:fields[1] + 2] = m.view(np.uint8).reshape(nblocks, 2)
    return raw


def dequantize(dtype, blocks):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    n = blocks.size // BLOCK_BYTES[dtype] * BLOCK_ELEMS[dtype]
    out = np.empty(n, dtype=np.float32)
    rc = lib().mrs_dequantize(GGML[dtype], _p(blocks), _p(out), ctypes.c_int64(n))
    assert rc == 0
    return out


def quantize_q8_1(x, k_padded=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, k = x.reshape(-1, x.shape[-1]).shape
    k_padded = k_padded or (k + 511) // 512 * 512
    y = np.zeros(rows * (k_padded // 32) * 36, dtype=np.uint8)
    lib().mrs_quantize_q8_1(_p(x), _p(y), k, k_padded, rows)
    return y, k_padded // 32


def mmvq_q8_1(dtype, w, yq, ncols, nrows, stride_col_y, batch):
    out = np.empty(batch * nrows, dtype=np.float64)
    rc = lib().mrs_mmvq_q8_1(GGML[dtype], _p(w), _p(yq), _p(out), ncols, nrows, stride_col_y, batch)
    assert rc == 0
    return out.reshape(batch, nrows)


def matmul_exact(dtype, w, x, ncols, nrows):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ncols)
    out = np.empty(x.shape[0] * nrows, dtype=np.float64)
    rc = lib().mrs_matmul_exact(GGML[dtype], _p(w), _p(x), _p(out), ncols, nrows, x.shape[0])
    assert rc == 0
    return out.reshape(x.shape[0], nrows)


def qmatmul_cpu(dtype, w, x, ncols, nrows, threads=1):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ncols)
    out = np.empty(x.shape[0] * nrows, dtype=np.float32)
    rc = lib().mrs_qmatmul_cpu(GGML[dtype], _p(w), _p(x), _p(out), ncols, nrows, x.shape[0], threads)
    assert rc == 0
    return out.reshape(x.shape[0], nrows)


def round_dtype(a, dt):
    """Round an f32 array through f16/bf16 (round-to-nearest-even)."""
    a = np.asarray(a, dtype=np.float32)
    if dt == "f32":
        return a
    if dt == "f16":
        return a.astype(np.float16).astype(np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def fused_glu(a, b, act, dt):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty_like(a)
    lib().mrs_fused_glu(_p(a), _p(b), _p(out), ctypes.c_int64(a.size), act, DT[dt])
    return out


def rms_norm(x, w, eps, dt):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, cols = x.reshape(-1, x.shape[-1]).shape
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty_like(x)
    lib().mrs_rms_norm(_p(x), _p(w), _p(out), rows, cols, ctypes.c_float(eps), DT[dt])
    return out


def add_rms_norm(x, res, w, eps, dt):
    x = np.ascontiguousarray(x, dtype=np.float32)
    res = np.ascontiguousarray(res, dtype=np.float32)
    rows, cols = x.reshape(-1, x.shape[-1]).shape
    w = np.ascontiguousarray(w, dtype=np.float32)
    s, n = np.empty_like(x), np.empty_like(x)
    lib().mrs_add_rms_norm(_p(x), _p(res), _p(w), _p(s), _p(n), rows, cols, ctypes.c_float(eps), DT[dt])
    return s, n


def rotary(q, k, cos, sin, positions, is_neox, head_size, rot_half, num_heads, num_kv_heads, dt):
    q = np.ascontiguousarray(q, dtype=np.float32).copy()
    k = np.ascontiguousarray(k, dtype=np.float32).copy()
    cos = np.ascontiguousarray(cos, dtype=np.float32)
    sin = np.ascontiguousarray(sin, dtype=np.float32)
    tokens = q.shape[0]
    pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.uint32)
    lib().mrs_rotary(_p(q), _p(k), _p(cos), _p(sin), _p(pos), int(is_neox), head_size, ctypes.c_int64(tokens),
                     rot_half, num_heads, num_kv_heads, ctypes.c_int64(q.size // tokens),
                     ctypes.c_int64(k.size // tokens), DT[dt])
    return q, k


def llama3_rope_table(max_pos, head_dim, theta, scaling=None):
    cos = np.empty((max_pos, head_dim // 2), dtype=np.float32)
    sin = np.empty_like(cos)
    if scaling is None:
        lib().mrs_llama3_rope_table(_p(cos), _p(sin), max_pos, head_dim, ctypes.c_float(theta), 0,
                                    ctypes.c_float(1), ctypes.c_float(1), ctypes.c_float(1), 0)
    else:
        lib().mrs_llama3_rope_table(_p(cos), _p(sin), max_pos, head_dim, ctypes.c_float(theta), 1,
                                    ctypes.c_float(scaling["factor"]), ctypes.c_float(scaling["low_freq_factor"]),
                                    ctypes.c_float(scaling["high_freq_factor"]),
                                    int(scaling["original_max_position_embeddings"]))
    return cos, sin


def reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, num_heads, head_size, block_size, x, layout):
    """key/value: uint16 views [T, H*D]; caches modified in place (uint16)."""
    T = key.shape[0]
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    lib().mrs_reshape_and_cache(_p(key), _p(value), _p(key_cache), _p(value_cache), _p(sm), T, num_heads,
                                head_size, block_size, x, key.shape[1], value.shape[1], layout)


def paged_attention(q, key_cache, value_cache, block_tables, context_lens, num_kv_heads, head_size, block_size,
                    scale, layout, dt, softcap=0.0, x=8):
    q = np.ascontiguousarray(q, dtype=np.float32)
    S, H, D = q.shape
    bt = np.ascontiguousarray(block_tables, dtype=np.int32)
    cl = np.ascontiguousarray(context_lens, dtype=np.int32)
    out = np.empty((S, H, D), dtype=np.float32)
    lib().mrs_paged_attention(_p(q), _p(key_cache), _p(value_cache), _p(bt), _p(cl), _p(out), S, H, num_kv_heads,
                              head_size, block_size, bt.shape[1], H * D, ctypes.c_float(scale),
                              ctypes.c_float(softcap), layout, x, DT[dt])
    return out


def ref_lib(name):
    """The unmodified reference kernels compiled by oracle/build_ref.sh (None if not built)."""
    path = os.path.join(REF_DIR, f"libref_{name}.so")
    return ctypes.CDLL(path) if os.path.exists(path) else None
