"""UQFF artifacts -> device-resident ggml block tensors.

UQFF is the reference's native quantized format: a directory of `<stem>-<n>.uqff` shards (each a
safetensors file whose entries follow naming conventions), `residual.safetensors` for the
unquantized tensors and the model's JSON assets
(REF docs/src/content/docs/reference/uqff-format.md; reader `mistralrs-quant/src/uqff/reader.rs`,
GGML-family layer entries `mistralrs-quant/src/gguf/mod.rs:142-161,261,781`):

    <key>.weight          U8   raw ggml blocks, as in GGUF
    <key>.weight.format   U8   scalar: 0 Gguf, 1 Unquant, 2 Hqq, 3 Fp8, 4 Afq, 5 F8Q8, 6 Mxfp4
    <key>.weight.dtype    U32  scalar: ggml type code (12 = Q4_K, 14 = Q6_K, 8 = Q8_0, ...)
    <key>.weight.shape    U32  vector [rows, cols]
    <key>.bias            optional
    uqff.version.{major,minor,patch}   U32 scalars; this reader speaks 1.0 .. 1.2

`<key>` is the layer's weight path (`model.layers.0.self_attn.q_proj`).  The container parsing is
C++ (`host/safetensors_reader.hpp`); blocks are uploaded as stored.
"""
import ctypes
import json
import os

import numpy as np
import torch

from .gguf_file import GGML_NAMES
from .kv_index import host_lib

UQFF_VERSION = (1, 2, 0)                      # REF mistralrs-quant/src/uqff/mod.rs:27-29
FORMATS = {0: "gguf", 1: "unquant", 2: "hqq", 3: "fp8", 4: "afq", 5: "f8q8", 6: "mxfp4"}  # REF lib.rs:1178-1186
_VERSION_KEYS = ("uqff.version.major", "uqff.version.minor", "uqff.version.patch")
_NP = {"BOOL": np.bool_, "U8": np.uint8, "I8": np.int8, "U16": np.uint16, "I16": np.int16, "F16": np.float16,
       "U32": np.uint32, "I32": np.int32, "F32": np.float32, "U64": np.uint64, "I64": np.int64, "F64": np.float64}

_bound = False


def _lib():
    global _bound
    L = host_lib()
    if not _bound:
        L.mrs_st_open.restype = ctypes.c_void_p
        L.mrs_st_tensor_data.restype = ctypes.c_void_p
        for f in ("n_tensors", "find", "n_metadata", "metadata"):
            getattr(L, f"mrs_st_{f}").restype = ctypes.c_int64
        _bound = True
    return L


class SafetensorsFile:
    """One safetensors container (mmap).  entries: name -> (dtype string, shape, offset, nbytes)."""

    def __init__(self, path):
        self.path = os.fspath(path)
        L = _lib()
        err = ctypes.create_string_buffer(512)
        h = L.mrs_st_open(self.path.encode(), err, ctypes.c_int64(len(err)))
        if not h:
            raise ValueError(err.value.decode(errors="replace") or f"cannot open {self.path}")
        self._h = ctypes.c_void_p(h)
        self.entries, self._index = {}, {}
        name, dt = ctypes.create_string_buffer(1024), ctypes.create_string_buffer(16)
        nd, dims = ctypes.c_int32(), (ctypes.c_int64 * 8)()
        off, nb = ctypes.c_int64(), ctypes.c_int64()
        for i in range(int(L.mrs_st_n_tensors(self._h))):
            L.mrs_st_tensor_info(self._h, ctypes.c_int64(i), name, ctypes.c_int64(len(name)), dt, ctypes.byref(nd), dims,
                                 ctypes.byref(off), ctypes.byref(nb))
            n = name.value.decode()
            self.entries[n] = (dt.value.decode(), tuple(int(dims[d]) for d in range(nd.value)), off.value, nb.value)
            self._index[n] = i
        self.metadata = {}
        k, v = ctypes.create_string_buffer(512), ctypes.create_string_buffer(4096)
        for i in range(int(L.mrs_st_n_metadata(self._h))):
            L.mrs_st_metadata(self._h, ctypes.c_int64(i), k, ctypes.c_int64(len(k)), v, ctypes.c_int64(len(v)))
            self.metadata[k.value.decode()] = v.value.decode(errors="replace")

    def close(self):
        if getattr(self, "_h", None):
            _lib().mrs_st_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def data(self, name) -> np.ndarray:
        """uint8 view of the entry's bytes (valid while the file is open)."""
        dt, shape, off, nb = self.entries[name]
        if nb == 0:
            return np.zeros(0, dtype=np.uint8)
        p = _lib().mrs_st_tensor_data(self._h, ctypes.c_int64(self._index[name]))
        a = np.frombuffer((ctypes.c_uint8 * nb).from_address(p), dtype=np.uint8)
        a.flags.writeable = False
        return a

    def array(self, name) -> np.ndarray:
        dt, shape, _, _ = self.entries[name]
        raw = self.data(name)
        if dt == "BF16":
            return raw.view(np.uint16).reshape(shape)     # caller reinterprets (numpy has no bf16)
        if dt not in _NP:
            raise ValueError(f"safetensors entry `{name}` has dtype {dt}, which numpy cannot represent")
        return raw.view(_NP[dt]).reshape(shape)


class UqffLayerInfo:
    __slots__ = ("key", "format", "dtype", "ggml_type", "shape", "nbytes", "has_bias")

    def __repr__(self):
        return f"UqffLayerInfo({self.key!r}, {self.format}, {self.dtype}, shape={self.shape})"


class UqffArchive:
    """All `.uqff` shards of one artifact (+ `residual.safetensors` and `config.json` found next to
    the first shard, as the reference's sibling-path lookup does)."""

    def __init__(self, paths, residual=None, config=None):
        if isinstance(paths, (str, os.PathLike)):
            p = os.fspath(paths)
            paths = sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith(".uqff")) if os.path.isdir(p) else [p]
        paths = [os.fspath(p) for p in paths]
        if not paths:
            raise ValueError("at least one UQFF shard is required")
        self.shards = [SafetensorsFile(p) for p in paths]
        self._where = {}
        versions = {}
        for si, sh in enumerate(self.shards):
            for name, (dt, shape, _, nb) in sh.entries.items():
                if name in _VERSION_KEYS:
                    if dt != "U32" or shape != () or nb != 4:
                        raise ValueError(f"UQFF version tensor `{name}` in `{sh.path}` must be a scalar U32.")
                    val = int(sh.array(name).reshape(-1)[0])
                    if name in versions and versions[name][0] != val:
                        raise ValueError(f"Conflicting UQFF version tensor `{name}` found in `{versions[name][1]}` "
                                         f"({versions[name][0]}) and `{sh.path}` ({val}).")
                    versions.setdefault(name, (val, sh.path))
                    continue
                if name in self._where:
                    raise ValueError(f"UQFF tensor `{name}` is duplicated across shards")
                self._where[name] = si
        if not all(k in versions for k in _VERSION_KEYS):
            raise ValueError("UQFF artifact has no version tag (pre-1.0 file); regenerate with `mistralrs quantize`.")
        self.version = tuple(versions[k][0] for k in _VERSION_KEYS)
        ours = ".".join(map(str, UQFF_VERSION))
        theirs = ".".join(map(str, self.version))
        if self.version[0] != UQFF_VERSION[0]:
            raise ValueError(f"UQFF version {theirs} is incompatible with this build ({ours}); regenerate with `mistralrs quantize`.")
        if self.version[1] > UQFF_VERSION[1]:
            raise ValueError(f"UQFF version {theirs} was written by a newer mistral.rs than this build ({ours}); upgrade mistral.rs.")
        base = os.path.dirname(os.path.abspath(paths[0]))
        rpath = residual if residual is not None else os.path.join(base, "residual.safetensors")
        self.residual = SafetensorsFile(rpath) if os.path.exists(rpath) else None
        cpath = config if config is not None else os.path.join(base, "config.json")
        self.config = json.load(open(cpath)) if os.path.exists(cpath) else None

    def close(self):
        for s in self.shards:
            s.close()
        if self.residual is not None:
            self.residual.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- catalogue ------------------------------------------------------------------------
    def contains(self, name):
        return name in self._where

    def layer_keys(self):
        """every self-describing layer: names ending in `.weight.format`, as `UqffReader` enumerates them"""
        return sorted(n[: -len(".weight.format")] for n in self._where if n.endswith(".weight.format"))

    def _entry(self, name):
        if name not in self._where:
            raise KeyError(f"cannot find UQFF tensor `{name}`")
        return self.shards[self._where[name]]

    def _scalar(self, name, dtype):
        sh = self._entry(name)
        dt, shape, _, _ = sh.entries[name]
        if dt != dtype or shape != ():
            raise ValueError(f"UQFF tensor `{name}` is not a {dtype.lower()} scalar.")
        return int(sh.array(name).reshape(-1)[0])

    def load_format(self, key) -> str:
        code = self._scalar(f"{key}.weight.format", "U8")
        if code not in FORMATS:
            raise ValueError(f"UQFF layer `{key}` has unknown format tag {code}")
        return FORMATS[code]

    def layer_info(self, key) -> UqffLayerInfo:
        info = UqffLayerInfo()
        info.key, info.format = key, self.load_format(key)
        info.has_bias = self.contains(f"{key}.bias")
        sh = self._entry(f"{key}.weight")
        wdt, wshape, _, wnb = sh.entries[f"{key}.weight"]
        info.nbytes = wnb
        if info.format == "gguf":
            info.ggml_type = self._scalar(f"{key}.weight.dtype", "U32")
            info.dtype = GGML_NAMES.get(info.ggml_type, f"ggml{info.ggml_type}")
            shp = self._entry(f"{key}.weight.shape")
            sdt, _, _, _ = shp.entries[f"{key}.weight.shape"]
            if sdt != "U32":
                raise ValueError(f"UQFF tensor `{key}.weight.shape` is not a u32 vector.")
            info.shape = tuple(int(v) for v in shp.array(f"{key}.weight.shape").reshape(-1))
            if wdt != "U8":
                raise ValueError(f"Expected U8 UQFF tensor `{key}.weight`, got {wdt}.")
        else:
            info.ggml_type, info.dtype, info.shape = None, wdt.lower(), wshape
        return info

    def load_qtensor(self, key, device):
        """GGML-family layer -> quant.QTensor(blocks on `device`, dtype name, (rows, cols))."""
        from . import BLOCK_BYTES, BLOCK_ELEMS
        from .quant import QTensor
        info = self.layer_info(key)
        if info.format != "gguf":
            raise NotImplementedError(f"UQFF layer `{key}` uses the {info.format} family; only GGML-family layers "
                                      "run on the block kernels")
        if info.dtype not in BLOCK_BYTES:
            raise NotImplementedError(f"UQFF layer `{key}` has ggml dtype {info.dtype}")
        rows = int(np.prod(info.shape[:-1])) if len(info.shape) > 1 else 1
        cols = info.shape[-1]
        want = rows * cols // BLOCK_ELEMS[info.dtype] * BLOCK_BYTES[info.dtype]
        if cols % BLOCK_ELEMS[info.dtype] or info.nbytes != want:
            raise ValueError(f"UQFF layer `{key}`: {info.nbytes} bytes of {info.dtype} do not match shape {info.shape}")
        data = torch.from_numpy(np.array(self._entry(f"{key}.weight").data(f"{key}.weight"))).to(device)
        return QTensor(data, info.dtype, (rows, cols))

    def load_tensor(self, name, device, dtype=None) -> torch.Tensor:
        """dense tensor from the shards or from residual.safetensors"""
        src = self.shards[self._where[name]] if name in self._where else self.residual
        if src is None or name not in src.entries:
            raise KeyError(f"cannot find UQFF tensor `{name}`")
        dt = src.entries[name][0]
        a = np.array(src.array(name))
        t = torch.from_numpy(a.view(np.int16)).view(torch.bfloat16) if dt == "BF16" else torch.from_numpy(a)
        t = t.to(device)
        return t.to(dtype) if dtype is not None else t


# ---- writing -------------------------------------------------------------------------------------
_ST = {np.dtype(v): k for k, v in _NP.items()}


def version_entries(version=UQFF_VERSION):
    """the three scalar U32 version tensors every shard carries (REF uqff/mod.rs:27-29, reader.rs version check)"""
    return {k: np.array(v, dtype=np.uint32) for k, v in zip(_VERSION_KEYS, version)}


def save_uqff(path, entries, version=UQFF_VERSION, metadata=None):
    """Write one `.uqff` shard: a safetensors container (8-byte little-endian header length, JSON header
    name -> {dtype, shape, data_offsets}, then the tensor bytes in header order, each 8-byte aligned through header
    padding only — tensors are packed back to back) holding `entries` (name -> numpy array; torch bf16 tensors are
    stored as BF16) plus the version tag.  The counterpart of `UqffArchive`; what `GgufMatMul.serialize_uqff` feeds."""
    entries = dict(entries)
    entries.update(version_entries(version))
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name in sorted(entries):
        a = entries[name]
        if isinstance(a, torch.Tensor):
            if a.dtype == torch.bfloat16:
                raw, st, shape = a.detach().cpu().contiguous().view(torch.int16).numpy().tobytes(), "BF16", tuple(a.shape)
            else:
                a = a.detach().cpu().numpy()
        if not isinstance(a, torch.Tensor):
            a = np.asarray(a)                       # (ascontiguousarray would turn a scalar into shape (1,))
            if not a.flags.c_contiguous:
                a = a.copy(order="C")
            if a.dtype not in _ST:
                raise ValueError(f"UQFF tensor `{name}`: unsupported dtype {a.dtype}")
            raw, st, shape = a.tobytes(), _ST[a.dtype], tuple(a.shape)
        header[name] = {"dtype": st, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    h = json.dumps(header, separators=(",", ":")).encode()
    h += b" " * (-len(h) % 8)
    with open(path, "wb") as f:
        f.write(len(h).to_bytes(8, "little"))
        f.write(h)
        for b in blobs:
            f.write(b)
