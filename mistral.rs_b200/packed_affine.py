"""Host-side mirror of the reference's packed-affine GGUF path (REF mistralrs-quant/src/gguf/packed_affine.rs):
`AffineFormatSpec` (:44-69), `minimum_batch` (:76), `PackedAffinePlan` (:94-135), the shape rule (:138-160) and
`PackedAffine::{new, forward}` (:456-602, :604-770).  The repack and the GEMM are calls through the C ABI of
libmrs_b200.so under the reference's own symbol names (`mrs_gguf_affine_repack_*`, `marlin_affine_{u4,u8}_*`)."""
import ctypes
import os
from dataclasses import dataclass

import torch

from . import lib

MARLIN_N_TILE, MARLIN_WIDE_TILE, MARLIN_K_TILE, MARLIN_MAX_PARALLEL = 64, 128, 64, 16
GGUF_AFFINE_MIN_BATCH = 8          # REF gguf/mod.rs GGUF_AFFINE_MIN_BATCH
AFFINE_DTYPES = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "q8_1", "q2_k", "q3_k", "q4_k", "q5_k", "q6_k", "q8_k"]
# type -> (format code, source block size, payload bits, group size, minimum batch)
_SPECS = {"q4_0": (2, 32, 4, 32, GGUF_AFFINE_MIN_BATCH), "q4_1": (3, 32, 4, 32, GGUF_AFFINE_MIN_BATCH), "q5_0": (6, 32, 8, 32, 16),
          "q5_1": (7, 32, 8, 32, 128), "q8_0": (8, 32, 8, 32, GGUF_AFFINE_MIN_BATCH), "q8_1": (9, 32, 8, 32, 1),
          "q2_k": (10, 256, 4, 16, GGUF_AFFINE_MIN_BATCH), "q3_k": (11, 256, 4, 16, GGUF_AFFINE_MIN_BATCH),
          "q4_k": (12, 256, 4, 32, GGUF_AFFINE_MIN_BATCH), "q5_k": (13, 256, 8, 32, GGUF_AFFINE_MIN_BATCH), "q6_k": (14, 256, 8, 16, 128),
          "q8_k": (15, 256, 8, 32, 1)}
SOURCE_BLOCK_BYTES = {"q4_0": 18, "q4_1": 20, "q5_0": 22, "q5_1": 24, "q8_0": 34, "q8_1": 36, "q2_k": 84, "q3_k": 110, "q4_k": 144,
                      "q5_k": 176, "q6_k": 210, "q8_k": 292}
_DT = {torch.float16: "f16", torch.bfloat16: "bf16"}


class AffineFormatSpec(tuple):
    """(format_code, source_block_size, payload_bits, group_size, min_batch)"""
    format_code = property(lambda s: s[0])
    source_block_size = property(lambda s: s[1])
    payload_bits = property(lambda s: s[2])
    group_size = property(lambda s: s[3])
    min_batch = property(lambda s: s[4])

    @staticmethod
    def for_dtype(source_dtype):
        return AffineFormatSpec(_SPECS[source_dtype]) if source_dtype in _SPECS else None

    def supports_f32_input(self):
        return self.format_code in (9, 15)


def minimum_batch(dtype):
    spec = AffineFormatSpec.for_dtype(dtype)
    return None if spec is None else spec.min_batch


def supports_marlin_shape(n, k):
    return (k % MARLIN_WIDE_TILE == 0 and n % MARLIN_N_TILE == 0) or (k % MARLIN_K_TILE == 0 and n % MARLIN_WIDE_TILE == 0)


def padded_n_for_shape(n, k):
    if k % MARLIN_WIDE_TILE == 0:
        tile = MARLIN_N_TILE
    elif k % MARLIN_K_TILE == 0:
        tile = MARLIN_WIDE_TILE
    else:
        return None
    padded = -(-n // tile) * tile
    return padded if supports_marlin_shape(padded, k) else None


@dataclass(frozen=True)
class PackedAffinePlan:
    format: AffineFormatSpec
    n: int
    padded_n: int
    k: int
    payload_bytes: int
    metadata_values: int
    metadata_bytes: int
    workspace_len: int
    total_bytes: int

    @staticmethod
    def new(source_dtype, n, k):
        fmt = AffineFormatSpec.for_dtype(source_dtype)
        if fmt is None:
            return None
        padded_n = padded_n_for_shape(n, k)
        if padded_n is None or n == 0 or k == 0 or k % fmt.source_block_size or max(n, padded_n, k) >= 2 ** 31:
            return None
        payload = padded_n * k * fmt.payload_bits // 8
        values = k // fmt.group_size * padded_n
        ws = padded_n // MARLIN_N_TILE * MARLIN_MAX_PARALLEL
        return PackedAffinePlan(fmt, n, padded_n, k, payload, values, values * 2, ws, payload + 2 * values * 2 + ws * 4)


BACKEND_ENV = "MISTRALRS_GGUF_AFFINE_BACKEND"


def enabled():
    """Off unless MISTRALRS_GGUF_AFFINE_BACKEND is "on" / "auto": the repacked copy is a second full set of weights
    (REF packed_affine.rs:295-310).  Read on every call (the reference latches it once per process)."""
    return os.environ.get(BACKEND_ENV, "off") in ("on", "auto")


def should_dispatch(source_dtype, shape, flat_batch, act_dtype, device_type):
    """The dispatcher's decision, host logic only (REF gguf/mod.rs:326-366 `packed_affine_for`): backend enabled, a
    format with an affine form, batch at or above that format's minimum, 16-bit activations on CUDA, tile-compatible shape."""
    if not enabled():
        return False
    min_batch = minimum_batch(source_dtype)
    if min_batch is None or flat_batch < min_batch:
        return False
    return act_dtype in _DT and device_type == "cuda" and len(shape) == 2 and PackedAffinePlan.new(source_dtype, shape[0], shape[1]) is not None


class PackedAffine:
    """A ggml block tensor re-tiled once into payload + scales + offsets, then multiplied on the tensor cores."""

    @staticmethod
    def supports(source_dtype, shape, dtype, device):
        return (len(shape) == 2 and dtype in _DT and torch.device(device).type == "cuda"
                and PackedAffinePlan.new(source_dtype, shape[0], shape[1]) is not None)

    def __init__(self, data: torch.Tensor, source_dtype: str, shape, dtype: torch.dtype):
        """data: uint8 ggml blocks [n * k / block * block_bytes] on a CUDA device (a `quant.QTensor`'s .data)."""
        if not PackedAffine.supports(source_dtype, shape, dtype, data.device):
            raise ValueError(f"packed GGUF affine does not support {source_dtype} {tuple(shape)} {dtype} on {data.device}")
        plan = PackedAffinePlan.new(source_dtype, shape[0], shape[1])
        want = plan.n * (plan.k // plan.format.source_block_size) * SOURCE_BLOCK_BYTES[source_dtype]
        if data.dtype != torch.uint8 or data.numel() != want:
            raise ValueError(f"source must be uint8[{want}]")
        self.plan, self.dtype, self.device = plan, dtype, data.device
        self.n, self.padded_n, self.k = plan.n, plan.padded_n, plan.k
        self.payload = torch.empty(plan.payload_bytes, dtype=torch.uint8, device=data.device)
        self.scales = torch.empty(plan.metadata_values, dtype=dtype, device=data.device)
        self.offsets = torch.empty(plan.metadata_values, dtype=dtype, device=data.device)
        self.workspace = torch.zeros(plan.workspace_len, dtype=torch.int32, device=data.device)   # the reference's lock array; unused here
        data = data.contiguous()
        fn = getattr(lib(), f"mrs_gguf_affine_repack_{_DT[dtype]}")
        rc = fn(ctypes.c_int32(plan.format.format_code), ctypes.c_void_p(data.data_ptr()), ctypes.c_void_p(self.payload.data_ptr()),
                ctypes.c_void_p(self.scales.data_ptr()), ctypes.c_void_p(self.offsets.data_ptr()), ctypes.c_int32(plan.k),
                ctypes.c_int32(plan.n), ctypes.c_int32(plan.padded_n), ctypes.c_size_t(torch.cuda.current_stream(data.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"GGUF affine repack failed with status {rc}")

    def forward(self, xs: torch.Tensor) -> torch.Tensor:
        if xs.device != self.device:
            raise ValueError("packed GGUF affine input and weight are on different devices")
        if xs.dtype != self.dtype:
            raise ValueError(f"packed GGUF affine parameter dtype {self.dtype} does not match input {xs.dtype}")
        if xs.dim() == 0:
            raise ValueError("packed GGUF affine input must have at least one dimension")
        k = xs.shape[-1]
        m = xs.numel() // k if k else 0
        if m == 0 or k != self.k:
            raise ValueError(f"packed GGUF affine shape mismatch: input {tuple(xs.shape)}, weight [{self.n}, {self.k}]")
        xs = xs.contiguous()
        if xs.data_ptr() % 16:
            xs = xs.clone()
        out = torch.empty(*xs.shape[:-1], self.padded_n, dtype=xs.dtype, device=xs.device)
        fn = getattr(lib(), f"marlin_affine_u{self.plan.format.payload_bits}_{_DT[self.dtype]}")
        rc = fn(ctypes.c_void_p(xs.data_ptr()), ctypes.c_void_p(self.payload.data_ptr()), ctypes.c_void_p(self.scales.data_ptr()),
                ctypes.c_void_p(self.offsets.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_int32(m), ctypes.c_int32(self.k),
                ctypes.c_int32(self.padded_n), ctypes.c_int32(self.plan.format.group_size), ctypes.c_void_p(self.workspace.data_ptr()),
                ctypes.c_int64(torch.cuda.current_stream(xs.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"Marlin matmul failed with status {rc}")
        return out if self.padded_n == self.n else out.narrow(-1, 0, self.n).contiguous()
