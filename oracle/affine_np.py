"""ORACLE (test infrastructure): numpy restatement of the packed-affine view of ggml blocks,
w = scale * q - offset with unsigned 4- or 8-bit q and one (scale, offset) per 16 or 32 weights.
REF mistralrs-quant/src/gguf/packed_affine.rs:44-69 (payload width / group size per type), :94-135 (sizes), :138-160
(shape rule and N padding); block layouts mistralrs-quant/kernels/mmvq_gguf/mmvq_gguf.cu:134-225.
Written from the block layouts with whole-array numpy indexing — independently of csrc/affine.cuh's per-segment C — and
pinned by `scale * q - offset == oracle.dequantize(...)` for the ten types the C oracle decodes (itself pinned against
gguf-py vectors)."""
import numpy as np

# type -> (ggml code, block elems, block bytes, payload bits, group, min batch)    REF packed_affine.rs:926-939
GGUF_AFFINE_MIN_BATCH = 8
SPECS = {"q4_0": (2, 32, 18, 4, 32, 8), "q4_1": (3, 32, 20, 4, 32, 8), "q5_0": (6, 32, 22, 8, 32, 16), "q5_1": (7, 32, 24, 8, 32, 128),
         "q8_0": (8, 32, 34, 8, 32, 8), "q8_1": (9, 32, 36, 8, 32, 1), "q2_k": (10, 256, 84, 4, 16, 8), "q3_k": (11, 256, 110, 4, 16, 8),
         "q4_k": (12, 256, 144, 4, 32, 8), "q5_k": (13, 256, 176, 8, 32, 8), "q6_k": (14, 256, 210, 8, 16, 128),
         "q8_k": (15, 256, 292, 8, 32, 1)}


def supports_marlin_shape(n, k):
    return (k % 128 == 0 and n % 64 == 0) or (k % 64 == 0 and n % 128 == 0)


def padded_n_for_shape(n, k):
    tile = 64 if k % 128 == 0 else 128 if k % 64 == 0 else None
    if tile is None:
        return None
    p = -(-n // tile) * tile
    return p if supports_marlin_shape(p, k) else None


def _f16(b, off):
    return b[:, off:off + 2].copy().view(np.float16).astype(np.float32)[:, 0]


def decompose(dtype, blocks):
    """blocks: uint8 [nblocks, block_bytes] -> q uint8 [nblocks, elems], scale f32 [nblocks, elems/group], offset likewise."""
    _, elems, bb, bits, group, _ = SPECS[dtype]
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, bb)
    nb = b.shape[0]
    if dtype in ("q4_0", "q4_1"):
        qs = b[:, bb - 16:]
        q = np.concatenate([qs & 0xF, qs >> 4], axis=1)
        d = _f16(b, 0)
        off = 8 * d if dtype == "q4_0" else -_f16(b, 2)
        return q, d[:, None], off[:, None]
    if dtype in ("q5_0", "q5_1"):
        ho = 2 if dtype == "q5_0" else 4
        qh = b[:, ho:ho + 4].copy().view("<u4")[:, 0]
        qs = b[:, ho + 4:]
        bit = ((qh[:, None] >> np.arange(32, dtype=np.uint32)[None]) & 1).astype(np.uint8)
        q = np.concatenate([qs & 0xF, qs >> 4], axis=1) | (bit << 4)
        d = _f16(b, 0)
        off = 16 * d if dtype == "q5_0" else -_f16(b, 2)
        return q, d[:, None], off[:, None]
    if dtype in ("q8_0", "q8_1"):
        q = (b[:, bb - 32:].view(np.int8).astype(np.int16) + 128).astype(np.uint8)
        d = _f16(b, 0)
        return q, d[:, None], (128 * d)[:, None]
    if dtype == "q8_k":
        d = b[:, :4].copy().view("<f4")[:, 0]
        q = (b[:, 4:260].view(np.int8).astype(np.int16) + 128).astype(np.uint8)
        s = np.repeat(d[:, None], 8, axis=1)
        return q, s, 128 * s
    e = np.arange(256)
    if dtype == "q2_k":
        n, j, l = e // 128, (e % 128) // 32, e % 32
        q = (b[:, 16 + 32 * n + l] >> (2 * j)) & 3
        sm = b[:, :16]
        d, dmin = _f16(b, 80), _f16(b, 82)
        return q.astype(np.uint8), d[:, None] * (sm & 0xF), dmin[:, None] * (sm >> 4)
    if dtype == "q3_k":
        n, j, l = e // 128, (e % 128) // 32, e % 32
        q = ((b[:, 32 + 32 * n + l] >> (2 * j)) & 3) | (((b[:, l] >> (4 * n + j)) & 1) << 2)
        s = b[:, 96:108].astype(np.int32)
        i = np.arange(16)
        lo = np.where(i < 8, s[:, i % 8] & 0xF, s[:, i % 8] >> 4)
        hi = (s[:, 8 + i % 4] >> (2 * (i // 4))) & 3
        sc = _f16(b, 108)[:, None] * ((lo | (hi << 4)) - 32).astype(np.float32)
        return q.astype(np.uint8), sc, 4 * sc
    if dtype in ("q4_k", "q5_k"):
        c, half, l = e // 64, (e % 64) // 32, e % 32
        qo = 16 if dtype == "q4_k" else 48
        byte = b[:, qo + 32 * c + l]
        q = np.where(half == 1, byte >> 4, byte & 0xF)
        if dtype == "q5_k":
            q = q | (((b[:, 16 + l] >> (2 * c + half)) & 1) << 4)
        s = b[:, 4:16].astype(np.int32)
        i = np.arange(8)
        sc = np.where(i < 4, s[:, i % 4] & 63, (s[:, 4 + i] % 16) | ((s[:, i - 4] >> 6) << 4))      # i - 4 wraps for i < 4: masked out
        mn = np.where(i < 4, s[:, 4 + i % 4] & 63, (s[:, 4 + i] >> 4) | ((s[:, i] >> 6) << 4))
        return q.astype(np.uint8), _f16(b, 0)[:, None] * sc, _f16(b, 2)[:, None] * mn
    if dtype == "q6_k":
        n, j, l = e // 128, (e % 128) // 32, e % 32
        lo = b[:, 64 * n + 32 * (j & 1) + l]
        nib = np.where(j >= 2, lo >> 4, lo & 0xF)
        q = nib | (((b[:, 128 + 32 * n + l] >> (2 * j)) & 3) << 4)
        sc = _f16(b, 208)[:, None] * b[:, 192:208].view(np.int8).astype(np.float32)
        return q.astype(np.uint8), sc, 32 * sc
    raise KeyError(dtype)


def _round16(a, bf16):
    a = np.asarray(a, dtype=np.float32)
    if not bf16:
        return a.astype(np.float16)
    u = a.view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def _to_f32(a16, bf16):
    return (a16.astype(np.uint32) << 16).view(np.float32) if bf16 else a16.astype(np.float32)


def repack(dtype, w_blocks, n, k, padded_n, bf16):
    """Row-major packed arrays as csrc/affine.cuh lays them out: payload [padded_n, k*bits/8] u8, scales / offsets
    [padded_n, k/group] as 16-bit patterns (uint16)."""
    _, elems, bb, bits, group, _ = SPECS[dtype]
    q, sc, of = decompose(dtype, np.asarray(w_blocks, dtype=np.uint8).reshape(-1, bb))
    q = q.reshape(n, k)
    if bits == 4:
        pay = (q[:, 0::2] | (q[:, 1::2] << 4)).astype(np.uint8)
    else:
        pay = q
    s16 = _round16(sc.reshape(n, k // group), bf16).view(np.uint16)
    o16 = _round16(of.reshape(n, k // group), bf16).view(np.uint16)
    pad = padded_n - n
    z = lambda a: np.concatenate([a, np.zeros((pad,) + a.shape[1:], a.dtype)]) if pad else a
    return z(pay), z(s16), z(o16)


def weights(dtype, w_blocks, n, k, bf16):
    """The weights the packed GEMM multiplies by, f32 [n, k]: fma(q, scale16, -offset16) (float64 here: exact)."""
    _, elems, bb, bits, group, _ = SPECS[dtype]
    q, sc, of = decompose(dtype, np.asarray(w_blocks, dtype=np.uint8).reshape(-1, bb))
    s = _to_f32(_round16(sc.reshape(n, k // group), bf16).view(np.uint16 if bf16 else np.float16), bf16)
    o = _to_f32(_round16(of.reshape(n, k // group), bf16).view(np.uint16 if bf16 else np.float16), bf16)
    w = q.reshape(n, k).astype(np.float64) * np.repeat(s, group, axis=1).astype(np.float64) - np.repeat(o, group, axis=1).astype(np.float64)
    return w.astype(np.float32)
