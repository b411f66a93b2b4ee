"""GGUF archives -> device-resident ggml block tensors.

Python mirror of the reference's archive layer (`mistralrs-quant/src/gguf/archive.rs`:
`GgufArchive::{open, metadata, tensors, tensor_info, contains_tensor, tensor_data, load_qtensor}`)
over the C++ reader in `host/gguf_reader.hpp` (mmap, header/metadata/tensor catalogue, alignment,
split shards).  Tensor payloads are uploaded exactly as stored: the kernels consume ggml blocks
in file layout, so loading is one host->device copy per tensor and no repack.
"""
import ctypes
import os

import numpy as np
import torch

from .kv_index import host_lib

# GGUF metadata value types (spec) and the ggml tensor types this package has kernels or dtypes for
VT_NAMES = {0: "u8", 1: "i8", 2: "u16", 3: "i16", 4: "u32", 5: "i32", 6: "f32", 7: "bool", 8: "str", 9: "arr",
            10: "u64", 11: "i64", 12: "f64"}
GGML_NAMES = {0: "f32", 1: "f16", 2: "q4_0", 3: "q4_1", 6: "q5_0", 7: "q5_1", 8: "q8_0", 9: "q8_1", 10: "q2_k",
              11: "q3_k", 12: "q4_k", 13: "q5_k", 14: "q6_k", 15: "q8_k", 24: "i8", 25: "i16", 26: "i32",
              27: "i64", 28: "f64", 30: "bf16"}
_NP_DENSE = {"f32": np.float32, "f16": np.float16, "i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64,
             "f64": np.float64}

_bound = False


def _lib():
    global _bound
    L = host_lib()
    if not _bound:
        L.mrs_gguf_open.restype = ctypes.c_void_p
        L.mrs_gguf_tensor_data.restype = ctypes.c_void_p
        for f in ("alignment", "n_tensors", "n_metadata", "find_tensor", "meta_str", "meta_arr_str", "meta_arr_num"):
            getattr(L, f"mrs_gguf_{f}").restype = ctypes.c_int64
        _bound = True
    return L


class GgufTensorInfo:
    __slots__ = ("name", "ggml_type", "dtype", "dims", "shape", "shard", "offset", "nbytes", "index")

    def __repr__(self):
        return f"GgufTensorInfo({self.name!r}, {self.dtype}, shape={self.shape}, offset={self.offset}, nbytes={self.nbytes})"


class GgufArchive:
    """`GgufArchive::open(paths)` — one file or all shards of a split model (any order)."""

    def __init__(self, paths):
        if isinstance(paths, (str, os.PathLike)):
            paths = [paths]
        paths = [os.fspath(p) for p in paths]
        L = _lib()
        arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
        err = ctypes.create_string_buffer(512)
        self._h = L.mrs_gguf_open(arr, ctypes.c_int32(len(paths)), err, ctypes.c_int64(len(err)))
        if not self._h:
            raise ValueError(err.value.decode(errors="replace") or "cannot open GGUF archive")
        self._h = ctypes.c_void_p(self._h)
        self.alignment = int(L.mrs_gguf_alignment(self._h))
        self._tensors = {}
        name = ctypes.create_string_buffer(1024)
        ty, nd, shard = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        dims = (ctypes.c_int64 * 8)()
        off, nb = ctypes.c_int64(), ctypes.c_int64()
        for i in range(int(L.mrs_gguf_n_tensors(self._h))):
            L.mrs_gguf_tensor_info(self._h, ctypes.c_int64(i), name, ctypes.c_int64(len(name)), ctypes.byref(ty),
                                   ctypes.byref(nd), dims, ctypes.byref(shard), ctypes.byref(off), ctypes.byref(nb))
            t = GgufTensorInfo()
            t.name, t.ggml_type, t.index = name.value.decode(), ty.value, i
            t.dtype = GGML_NAMES.get(ty.value, f"ggml{ty.value}")
            t.dims = tuple(int(dims[d]) for d in range(nd.value))   # ggml order: dims[0] innermost
            t.shape = tuple(reversed(t.dims))                        # row-major shape ([N, K] for a linear)
            t.shard, t.offset, t.nbytes = shard.value, off.value, nb.value
            self._tensors[t.name] = t
        self._metadata = None

    def close(self):
        if getattr(self, "_h", None):
            _lib().mrs_gguf_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- metadata -------------------------------------------------------------------------
    def metadata(self):
        """dict key -> python value (ints, floats, bools, strings, lists), like `GgufArchive::metadata`."""
        if self._metadata is not None:
            return self._metadata
        L, h = _lib(), self._h
        out = {}
        key = ctypes.create_string_buffer(1024)
        vt, at, alen = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        for i in range(int(L.mrs_gguf_n_metadata(h))):
            L.mrs_gguf_meta_key(h, ctypes.c_int64(i), key, ctypes.c_int64(len(key)), ctypes.byref(vt), ctypes.byref(at),
                                ctypes.byref(alen))
            k = key.value
            t = VT_NAMES.get(vt.value)
            if t == "str":
                out[k.decode()] = self._str(L.mrs_gguf_meta_str, k)
            elif t in ("f32", "f64"):
                d = ctypes.c_double()
                L.mrs_gguf_meta_float(h, k, ctypes.byref(d))
                out[k.decode()] = d.value
            elif t == "arr":
                n = alen.value
                if VT_NAMES.get(at.value) == "str":
                    out[k.decode()] = [self._str(L.mrs_gguf_meta_arr_str, k, j) for j in range(n)]
                else:
                    fl = (ctypes.c_double * max(n, 1))()
                    it = (ctypes.c_int64 * max(n, 1))()
                    L.mrs_gguf_meta_arr_num(h, k, ctypes.c_int64(0), ctypes.c_int64(n), fl, it)
                    isf = VT_NAMES.get(at.value) in ("f32", "f64")
                    vals = list(fl[:n]) if isf else list(it[:n])
                    out[k.decode()] = [bool(v) for v in vals] if VT_NAMES.get(at.value) == "bool" else vals
            else:
                v = ctypes.c_int64()
                L.mrs_gguf_meta_int(h, k, ctypes.byref(v))
                out[k.decode()] = bool(v.value) if t == "bool" else v.value
        self._metadata = out
        return out

    def _str(self, fn, key, *idx):
        args = [self._h, key] + [ctypes.c_int64(i) for i in idx]
        n = int(fn(*args, None, ctypes.c_int64(0)))
        buf = ctypes.create_string_buffer(n + 1)
        fn(*args, buf, ctypes.c_int64(n + 1))
        return buf.raw[:n].decode(errors="replace")

    def metadata_value(self, key, default=None):
        return self.metadata().get(key, default)

    # ---- tensors --------------------------------------------------------------------------
    def tensors(self):
        return self._tensors

    def contains_tensor(self, name):
        return name in self._tensors

    def tensor_info(self, name) -> GgufTensorInfo:
        try:
            return self._tensors[name]
        except KeyError:
            raise KeyError(f"cannot find GGUF tensor `{name}`") from None

    def tensor_data(self, name) -> np.ndarray:
        """uint8 view of the tensor's bytes inside the mapping (valid while the archive is open)."""
        t = self.tensor_info(name)
        if t.nbytes < 0:
            raise ValueError(f"cannot determine the exact byte length of GGUF tensor `{name}` with dtype {t.ggml_type}")
        p = _lib().mrs_gguf_tensor_data(self._h, ctypes.c_int64(t.index))
        buf = (ctypes.c_uint8 * t.nbytes).from_address(p)
        a = np.frombuffer(buf, dtype=np.uint8)
        a.flags.writeable = False
        return a

    def load_dense(self, name, device, dtype=None) -> torch.Tensor:
        """f32 / f16 / bf16 / integer tensors as a torch tensor of `shape` on `device`."""
        t = self.tensor_info(name)
        raw = np.array(self.tensor_data(name))  # copy out of the mapping
        if t.dtype == "bf16":
            x = torch.from_numpy(raw.view(np.int16).reshape(t.shape)).view(torch.bfloat16)
        elif t.dtype in _NP_DENSE:
            x = torch.from_numpy(raw.view(_NP_DENSE[t.dtype]).reshape(t.shape))
        else:
            raise ValueError(f"GGUF tensor `{name}` has block dtype {t.dtype}; use load_qtensor")
        x = x.to(device)
        return x.to(dtype) if dtype is not None else x

    def load_qtensor(self, name, device):
        """`GgufArchive::load_qtensor`: the tensor's ggml blocks on `device`, as stored.
        Returns quant.QTensor(bytes, dtype name, (rows, cols))."""
        from .quant import QTensor
        t = self.tensor_info(name)
        if t.dtype in _NP_DENSE or t.dtype == "bf16":
            raise ValueError(f"GGUF tensor `{name}` is dense ({t.dtype}); use load_dense")
        if len(t.shape) < 1:
            raise ValueError(f"GGUF tensor `{name}` has no dimensions")
        cols = t.shape[-1]
        rows = int(np.prod(t.shape[:-1])) if len(t.shape) > 1 else 1
        data = torch.from_numpy(np.array(self.tensor_data(name))).to(device)
        return QTensor(data, t.dtype, (rows, cols))
