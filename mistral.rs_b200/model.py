"""Llama-family decode runner over the C ABI (`mrs_llama_decode_step`, include/mrs_b200_model.h).

Python here is only the harness: it allocates device memory with torch, fills the C structs with
raw pointers and drives CUDA-graph capture/replay.  The layer stack itself (which kernels run, in
which order) is C++ inside libmrs_b200.so.

Synthetic weights follow SURVEY §8(d): ggml blocks with uniformly random quants over their full
bit range, f16 scales d = 2^U(-9,-7), dmin = d*U(0,0.5); tensor-type map for Q4_K_M = the
llama.cpp recipe (Q4_K everywhere; Q6_K for `output`, and for attn_v + ffn_down on layers
i < n/8, i >= 7n/8 or (i - n/8) % 3 == 2).
"""
import ctypes
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import BLOCK_BYTES, BLOCK_ELEMS, GGML, kv_index, lib

F16_FIELDS = {"q4_0": [0], "q4_1": [0, 2], "q5_0": [0], "q5_1": [0, 2], "q8_0": [0],
              "q2_k": [80, 82], "q3_k": [108], "q4_k": [0, 2], "q5_k": [0, 2], "q6_k": [208]}


@dataclass
class LlamaConfig:
    hidden: int = 4096
    inter: int = 14336
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    head_dim: int = 128
    vocab: int = 128256
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: dict = None
    max_pos: int = 8192
    quant: str = "q4_k_m"      # "q4_k_m" | any ggml type name for a uniform model
    block_size: int = 16
    name: str = "llama-3-8b"
    rope_neox: bool = True     # rotate-half pairing; GGUF llama files use the interleaved pairing (False)
    rope_freq_factors: object = None   # optional per-frequency divisors (GGUF `rope_freqs.weight`, Llama-3.1 scaling)
    synth_scale_exp: tuple = (-9, -7)  # synthetic weights: block scales d = 2^U(lo, hi)

    # Real-size models take block super-scales 2^U(-15,-13): with the 6-bit sub-scales (1..63) and 4-bit
    # quants of the k-quants that gives |w| ~ 0.016, i.e. unit-scale activations at K = 4096 .. 28672.
    # SURVEY §8(d)'s 2^U(-9,-7) (|w| ~ 1) sends the 14336-wide down projection past f16's range, and the
    # Q8_1 block scale is a half (REF mmvq_gguf.cu:146-152): the reference itself would produce NaN.
    @staticmethod
    def llama3_8b(**kw):
        kw.setdefault("synth_scale_exp", (-15, -13))
        return LlamaConfig(rope_scaling=None, **kw)

    @staticmethod
    def llama3_70b(**kw):
        kw.setdefault("synth_scale_exp", (-15, -13))
        return LlamaConfig(hidden=8192, inter=28672, n_layers=80, n_heads=64, n_kv_heads=8, name="llama-3-70b", **kw)

    @staticmethod
    def tinyllama(**kw):
        return LlamaConfig(hidden=2048, inter=5632, n_layers=22, n_heads=32, n_kv_heads=4, head_dim=64, vocab=32000,
                           rope_theta=10000.0, name="tinyllama-1.1b", **kw)

    @staticmethod
    def tiny_test(**kw):
        d = dict(hidden=512, inter=1024, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, vocab=1024,
                 rope_theta=10000.0, max_pos=512, name="tiny-test")
        d.update(kw)
        return LlamaConfig(**d)


def tensor_type(cfg: LlamaConfig, name: str, layer: int = 0) -> str:
    """ggml type of a tensor under cfg.quant (llama.cpp Q4_K_M recipe)."""
    if cfg.quant != "q4_k_m":
        return cfg.quant
    if name == "output":
        return "q6_k"
    if name in ("attn_v", "ffn_down"):
        n = cfg.n_layers
        if layer < n // 8 or layer >= 7 * n // 8 or (layer - n // 8) % 3 == 2:
            return "q6_k"
    return "q4_k"


TENSOR_IDS = {"attn_q": 0, "attn_k": 1, "attn_v": 2, "attn_output": 3, "ffn_gate": 4, "ffn_up": 5, "ffn_down": 6,
              "attn_norm": 7, "ffn_norm": 8, "token_embd": 9, "output": 10, "output_norm": 11}


def synth_blocks(dtype: str, nblocks: int, seed: int, scale_exp=(-9, -7)) -> np.ndarray:
    """uint8 [nblocks, block_bytes] per SURVEY §8(d), numpy PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bb = BLOCK_BYTES[dtype]
    raw = rng.integers(0, 256, size=(nblocks, bb), dtype=np.uint8)
    f = F16_FIELDS[dtype]
    d = np.exp2(rng.uniform(scale_exp[0], scale_exp[1], size=nblocks)).astype(np.float16)
    raw[:, f[0]:f[0] + 2] = d.view(np.uint8).reshape(nblocks, 2)
    if len(f) > 1:
        m = (d.astype(np.float32) * rng.uniform(0, 0.5, size=nblocks)).astype(np.float16)
        raw[:, f[1]:f[1] + 2] = m.view(np.uint8).reshape(nblocks, 2)
    return raw


def tensor_seed(layer: int, name: str) -> int:
    return 0xB200 + layer * 16 + TENSOR_IDS[name]


class _QW(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("ggml_type", ctypes.c_int32), ("rows", ctypes.c_int32),
                ("cols", ctypes.c_int32)]


class _Layer(ctypes.Structure):
    _fields_ = [(n, _QW) for n in ("wq", "wk", "wv", "wo", "w_gate", "w_up", "w_down")] + \
               [("attn_norm", ctypes.c_void_p), ("ffn_norm", ctypes.c_void_p),
                ("k_cache", ctypes.c_void_p), ("v_cache", ctypes.c_void_p)]


_AR_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p)


class _Step(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("hidden", "n_layers", "n_heads", "n_kv_heads", "head_dim", "vocab",
                                              "block_size", "act_dtype")] + \
               [("rms_eps", ctypes.c_float), ("sm_scale", ctypes.c_float), ("rope_neox", ctypes.c_int32),
                ("pdl", ctypes.c_int32), ("layers", ctypes.POINTER(_Layer)), ("tok_embd", _QW), ("lm_head", _QW),
                ("final_norm", ctypes.c_void_p), ("rope_cos", ctypes.c_void_p), ("rope_sin", ctypes.c_void_p),
                ("batch", ctypes.c_int32), ("padded_tiles", ctypes.c_int32), ("max_blocks_per_seq", ctypes.c_int32),
                ("skip_mask", ctypes.c_int32), ("fused_attention", ctypes.c_int32), ("reserved0", ctypes.c_int32)] + \
               [(n, ctypes.c_void_p) for n in ("token_ids", "positions", "slot_mapping", "kv_indptr", "kv_indices",
                                               "kv_last_page_len", "request_indices", "kv_tile_indices", "o_indptr",
                                               "kv_chunk_size", "block_valid_mask", "x", "x2", "q", "k", "v",
                                               "attn_out", "act", "logits", "tmp_v", "tmp_s", "out_token",
                                               "attn_counters", "argmax_scratch")] + \
               [("all_reduce", _AR_FN), ("all_reduce_user", ctypes.c_void_p), ("tp", ctypes.c_void_p)]


class _TpCtx(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("peer_base", ctypes.c_void_p * 8),
                ("flags_offset", ctypes.c_int64), ("slot_offset", ctypes.c_int64 * 2), ("seq_counter", ctypes.c_void_p),
                ("ll_offset", ctypes.c_int64), ("ll_slot_stride", ctypes.c_int64), ("ll_src_stride", ctypes.c_int64)]


class PeerAllReduce:
    """Symmetric buffer + peer mappings for the in-graph peer-memory all-reduce (`mrs_tp_ctx`).
    One process per GPU; the rendezvous goes through torch.distributed._symmetric_memory (plumbing:
    allocation + IPC handle exchange), the data path is our own kernel over NVLink loads."""

    def __init__(self, elems: int, dtype, device, group=None, low_latency=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("peer-memory all-reduce supports up to 8 ranks (one NVSwitch domain)")
        slot = (elems * torch.empty(0, dtype=dtype).element_size() + 255) // 256 * 256
        # low-latency region: [2 slots][world sources][4 bytes per element] of {element pair, sequence number} words
        if low_latency is None:
            low_latency = os.environ.get("MRS_TP_LL", "1") != "0"
        ll_src = (elems * 4 + 255) // 256 * 256 if low_latency else 0
        ll_slot = self.world * ll_src
        self.low_latency = bool(low_latency)
        self.buf = symm.empty(256 + 2 * slot + 2 * ll_slot, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, group)
        self.seq = torch.zeros(2, dtype=torch.int32, device=device)   # [all-reduces so far, time-out flag]
        c = _TpCtx()
        c.world, c.rank = self.world, self.rank
        for r in range(self.world):
            c.peer_base[r] = int(self.handle.buffer_ptrs[r])
        c.flags_offset, c.slot_offset[0], c.slot_offset[1] = 0, 256, 256 + slot
        c.seq_counter = self.seq.data_ptr()
        if low_latency:
            c.ll_offset, c.ll_slot_stride, c.ll_src_stride = 256 + 2 * slot, ll_slot, ll_src
        self.ctx = c
        dist.barrier(group)

    def pointer(self):
        return ctypes.addressof(self.ctx)

    def timed_out(self):
        """True when a low-latency all-reduce gave up waiting for a peer (ranks out of step): results are invalid."""
        return bool(int(self.seq[1].item()) != 0)


def rope_tables(cfg: LlamaConfig):
    """cos/sin [max_pos, head_dim/2] f32 — REF mistralrs-core/src/layers.rs:1071-1160."""
    half = cfg.head_dim // 2
    inv = (1.0 / np.power(np.float32(cfg.rope_theta), np.arange(0, cfg.head_dim, 2, dtype=np.float32) / np.float32(cfg.head_dim))).astype(np.float32)
    sc = cfg.rope_scaling
    if sc is not None:
        low_wl = np.float32(sc["original_max_position_embeddings"]) / np.float32(sc["low_freq_factor"])
        high_wl = np.float32(sc["original_max_position_embeddings"]) / np.float32(sc["high_freq_factor"])
        out = []
        for f in inv:
            wl = np.float32(2 * np.pi) / f
            if wl < high_wl:
                out.append(f)
            elif wl > low_wl:
                out.append(f / np.float32(sc["factor"]))
            else:
                smooth = (np.float32(sc["original_max_position_embeddings"]) / wl - np.float32(sc["low_freq_factor"])) / \
                         (np.float32(sc["high_freq_factor"]) - np.float32(sc["low_freq_factor"]))
                out.append((1 - smooth) * f / np.float32(sc["factor"]) + smooth * f)
        inv = np.array(out, dtype=np.float32)
    if cfg.rope_freq_factors is not None:   # llama.cpp stores Llama-3.1's scaling as theta / factor per frequency
        inv = (inv / np.asarray(cfg.rope_freq_factors, dtype=np.float32)).astype(np.float32)
    freqs = np.arange(cfg.max_pos, dtype=np.float32)[:, None] * inv[None, :]
    assert freqs.shape[1] == half
    return np.cos(freqs).astype(np.float32), np.sin(freqs).astype(np.float32)


def compute_kv_shard(total_kv_heads: int, head_dim: int, rank: int, world: int):
    """Rows of a K / V projection kept by `rank`: (first_row, rows).  REF mistralrs-quant/src/distributed/layers.rs:2692-2715
    (`compute_kv_shard`): KV heads are split over the ranks; when there are more ranks than KV heads every head is
    REPLICATED on world / kv_heads consecutive ranks (each rank then holds exactly one)."""
    if world == 1:
        return 0, total_kv_heads * head_dim
    if world > total_kv_heads:
        if world % total_kv_heads:
            raise ValueError(f"tensor-parallel size {world} must be a multiple of the {total_kv_heads} KV heads to replicate them")
        replicate = world // total_kv_heads
        return (rank // replicate) * head_dim, head_dim
    if total_kv_heads % world:
        raise ValueError(f"tensor-parallel size {world} must divide the {total_kv_heads} KV heads")
    per = total_kv_heads // world
    return rank * per * head_dim, per * head_dim


def compute_n_kv_groups(total_kv_heads: int, n_heads: int, world: int) -> int:
    """Query heads per local KV head under KV-head replication.  REF distributed/layers.rs:2718-2733."""
    replicate = world // total_kv_heads if world > total_kv_heads else 1
    return (n_heads // total_kv_heads) // replicate if replicate else n_heads // total_kv_heads


class LlamaWeights:
    """Synthetic device-resident weights (optionally one tensor-parallel shard)."""

    def __init__(self, cfg: LlamaConfig, device, dtype=torch.bfloat16, tp_rank=0, tp_size=1, keep_host=False, fast_synth=False):
        """fast_synth: generate the (shard-shaped) blocks on the device with torch's RNG instead of numpy
        PCG64 on the host — same distribution, not the SURVEY §8(d) byte stream; for throughput runs of
        large models (config 3's 8 GB, config 5's 40 GB) where no oracle comparison is made."""
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.fast_synth = bool(fast_synth)
        if fast_synth and keep_host:
            raise ValueError("fast_synth weights have no host copy")
        self.host = {} if keep_host else None
        H, I = cfg.hidden, cfg.inter
        nq, nkv = cfg.n_heads * cfg.head_dim, cfg.n_kv_heads * cfg.head_dim
        assert cfg.n_heads % tp_size == 0 and I % (tp_size * 256) == 0
        compute_kv_shard(cfg.n_kv_heads, cfg.head_dim, tp_rank, tp_size)      # raises on an impossible KV head layout
        self.layers = []
        self.nbytes = 0
        for l in range(cfg.n_layers):
            L = {}
            for name, rows, cols, kind in (("attn_q", nq, H, "col"), ("attn_k", nkv, H, "kv"), ("attn_v", nkv, H, "kv"),
                                           ("attn_output", H, nq, "row"), ("ffn_gate", I, H, "col"),
                                           ("ffn_up", I, H, "col"), ("ffn_down", H, I, "row")):
                L[name] = self._qtensor(l, name, rows, cols, kind)
            for name in ("attn_norm", "ffn_norm"):
                L[name] = self._norm(l, name)
            self.layers.append(L)
        self.tok_embd = self._qtensor(0, "token_embd", cfg.vocab, H, "rep")
        self.output = self._qtensor(0, "output", cfg.vocab, H, "rep")
        self.output_norm = self._norm(0, "output_norm")
        cos, sin = rope_tables(cfg)
        self.rope_cos = torch.from_numpy(cos).to(device).to(dtype)
        self.rope_sin = torch.from_numpy(sin).to(device).to(dtype)

    # ---- real weights -------------------------------------------------------------------------
    GGUF_NAMES = {"attn_q": "col", "attn_k": "kv", "attn_v": "kv", "attn_output": "row", "ffn_gate": "col",
                  "ffn_up": "col", "ffn_down": "row"}

    @staticmethod
    def config_from_gguf(ar) -> "LlamaConfig":
        """Hyper-parameters from GGUF metadata (`llama.*` keys, as the reference's
        `quantized_llama.rs::PropsGGUF` reads them)."""
        md = ar.metadata()
        arch = md.get("general.architecture", "llama")

        def g(key, default=None):
            v = md.get(f"{arch}.{key}", default)
            if v is None:
                raise KeyError(f"GGUF metadata key `{arch}.{key}` is missing")
            return v
        hidden, n_heads = int(g("embedding_length")), int(g("attention.head_count"))
        head_dim = int(md.get(f"{arch}.attention.key_length", md.get(f"{arch}.rope.dimension_count", hidden // n_heads)))
        emb = ar.tensor_info("token_embd.weight")
        factors = None
        if ar.contains_tensor("rope_freqs.weight"):
            factors = ar.load_dense("rope_freqs.weight", "cpu").float().numpy()
        return LlamaConfig(hidden=hidden, inter=int(g("feed_forward_length")), n_layers=int(g("block_count")),
                           n_heads=n_heads, n_kv_heads=int(g("attention.head_count_kv", n_heads)), head_dim=head_dim,
                           vocab=int(md.get(f"{arch}.vocab_size", emb.shape[0])),
                           rms_eps=float(g("attention.layer_norm_rms_epsilon", 1e-5)),
                           rope_theta=float(g("rope.freq_base", 10000.0)), rope_scaling=None,
                           max_pos=int(g("context_length", 4096)), quant="gguf",
                           name=str(md.get("general.name", arch)), rope_neox=False, rope_freq_factors=factors)

    @classmethod
    def from_gguf(cls, ar, device, dtype=torch.bfloat16, tp_rank=0, tp_size=1, keep_host=False, max_pos=None):
        """Device-resident weights of a llama-architecture GGUF archive (`gguf_file.GgufArchive`):
        ggml blocks uploaded as stored (column/row TP shards cut on block boundaries, as for the
        synthetic model), norm vectors converted to the activation dtype, `output.weight` falling
        back to the tied `token_embd.weight`.  Tensor names: llama.cpp's (`blk.N.attn_q.weight` …)."""
        self = cls.__new__(cls)
        cfg = cls.config_from_gguf(ar)
        if max_pos is not None:
            cfg.max_pos = int(max_pos)
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.host = {} if keep_host else None
        self.layers, self.nbytes = [], 0
        if cfg.n_heads % tp_size:
            raise ValueError("tensor-parallel size must divide the head counts")
        compute_kv_shard(cfg.n_kv_heads, cfg.head_dim, tp_rank, tp_size)      # raises on an impossible KV head layout
        for l in range(cfg.n_layers):
            L = {}
            for name, kind in cls.GGUF_NAMES.items():
                L[name] = self._gguf_qtensor(ar, f"blk.{l}.{name}.weight", (l, name), kind)
            for name in ("attn_norm", "ffn_norm"):
                L[name] = self._gguf_norm(ar, f"blk.{l}.{name}.weight", (l, name))
            self.layers.append(L)
        self.tok_embd = self._gguf_qtensor(ar, "token_embd.weight", (0, "token_embd"), "rep")
        out_name = "output.weight" if ar.contains_tensor("output.weight") else "token_embd.weight"
        self.output = self._gguf_qtensor(ar, out_name, (0, "output"), "rep")
        self.output_norm = self._gguf_norm(ar, "output_norm.weight", (0, "output_norm"))
        cos, sin = rope_tables(cfg)
        self.rope_cos = torch.from_numpy(cos).to(device).to(dtype)
        self.rope_sin = torch.from_numpy(sin).to(device).to(dtype)
        return self

    UQFF_NAMES = {"attn_q": "self_attn.q_proj", "attn_k": "self_attn.k_proj", "attn_v": "self_attn.v_proj",
                  "attn_output": "self_attn.o_proj", "ffn_gate": "mlp.gate_proj", "ffn_up": "mlp.up_proj",
                  "ffn_down": "mlp.down_proj"}

    @staticmethod
    def config_from_hf(cj: dict) -> "LlamaConfig":
        """Hyper-parameters from a Hugging Face `config.json` (llama family), the source the
        reference's non-GGUF loaders read (`models/llama.rs::Config`)."""
        heads = int(cj["num_attention_heads"])
        sc = cj.get("rope_scaling")
        if sc is not None:
            kind = sc.get("rope_type", sc.get("type"))
            if kind != "llama3":
                raise NotImplementedError(f"rope_scaling type {kind!r} is not supported (llama3 only)")
            sc = {k: sc[k] for k in ("factor", "low_freq_factor", "high_freq_factor", "original_max_position_embeddings")}
        return LlamaConfig(hidden=int(cj["hidden_size"]), inter=int(cj["intermediate_size"]),
                           n_layers=int(cj["num_hidden_layers"]), n_heads=heads,
                           n_kv_heads=int(cj.get("num_key_value_heads", heads)),
                           head_dim=int(cj.get("head_dim") or int(cj["hidden_size"]) // heads), vocab=int(cj["vocab_size"]),
                           rms_eps=float(cj.get("rms_norm_eps", 1e-5)), rope_theta=float(cj.get("rope_theta", 10000.0)),
                           rope_scaling=sc, max_pos=int(cj.get("max_position_embeddings", 4096)), quant="uqff",
                           name=str(cj.get("_name_or_path") or cj.get("model_type", "llama")), rope_neox=True)

    @classmethod
    def from_uqff(cls, ar, device, dtype=torch.bfloat16, tp_rank=0, tp_size=1, keep_host=False, max_pos=None):
        """Device-resident weights of a UQFF artifact (`uqff_file.UqffArchive`): GGML-family layer
        entries uploaded as stored, norms from `residual.safetensors`, hyper-parameters from
        `config.json`; Hugging Face tensor paths (`model.layers.N.self_attn.q_proj` …) and the
        rotate-half RoPE pairing, so the fused attention path applies."""
        if ar.config is None:
            raise ValueError("UQFF artifact has no config.json next to its shards")
        self = cls.__new__(cls)
        cfg = cls.config_from_hf(ar.config)
        if max_pos is not None:
            cfg.max_pos = int(max_pos)
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.host = {} if keep_host else None
        self.layers, self.nbytes = [], 0
        if cfg.n_heads % tp_size:
            raise ValueError("tensor-parallel size must divide the head counts")
        compute_kv_shard(cfg.n_kv_heads, cfg.head_dim, tp_rank, tp_size)      # raises on an impossible KV head layout
        for l in range(cfg.n_layers):
            L = {}
            for name, kind in cls.GGUF_NAMES.items():
                L[name] = self._uqff_qtensor(ar, f"model.layers.{l}.{cls.UQFF_NAMES[name]}", (l, name), kind)
            L["attn_norm"] = self._uqff_norm(ar, f"model.layers.{l}.input_layernorm.weight", (l, "attn_norm"))
            L["ffn_norm"] = self._uqff_norm(ar, f"model.layers.{l}.post_attention_layernorm.weight", (l, "ffn_norm"))
            self.layers.append(L)
        if not ar.contains("model.embed_tokens.weight.format"):
            raise NotImplementedError("this UQFF artifact keeps dense token embeddings in residual.safetensors (UQFF <= 1.1); "
                                      "the embedding gather kernel takes ggml block types")
        self.tok_embd = self._uqff_qtensor(ar, "model.embed_tokens", (0, "token_embd"), "rep")
        tied = bool(ar.config.get("tie_word_embeddings", False)) or not ar.contains("lm_head.weight.format")
        self.output = self._uqff_qtensor(ar, "model.embed_tokens" if tied else "lm_head", (0, "output"), "rep")
        self.output_norm = self._uqff_norm(ar, "model.norm.weight", (0, "output_norm"))
        cos, sin = rope_tables(cfg)
        self.rope_cos = torch.from_numpy(cos).to(device).to(dtype)
        self.rope_sin = torch.from_numpy(sin).to(device).to(dtype)
        return self

    def _uqff_norm(self, ar, name, key):
        t = ar.load_tensor(name, self.device, self.dtype).reshape(-1).contiguous()
        if self.host is not None:
            self.host[key] = t.float().cpu().numpy()
        return t

    def _uqff_qtensor(self, ar, key_path, key, kind):
        q = ar.load_qtensor(key_path, "cpu")
        rows, cols = q.shape
        be, bb = BLOCK_ELEMS[q.dtype], BLOCK_BYTES[q.dtype]
        full = q.data.numpy().reshape(rows, cols // be, bb)
        full, rows, cols = self._shard(full, rows, cols, be, kind)
        flat = np.ascontiguousarray(full).reshape(-1)
        t = torch.from_numpy(np.array(flat)).to(self.device)
        self.nbytes += t.numel()
        if self.host is not None:
            self.host[key] = np.array(flat)
        return (t, q.dtype, rows, cols)

    def _gguf_norm(self, ar, name, key):
        t = ar.load_dense(name, self.device, self.dtype).reshape(-1).contiguous()
        if self.host is not None:
            self.host[key] = t.float().cpu().numpy()
        return t

    def _gguf_qtensor(self, ar, name, key, kind):
        info = ar.tensor_info(name)
        if info.dtype not in BLOCK_BYTES:
            raise NotImplementedError(f"GGUF tensor `{name}` is {info.dtype}: the decode kernels take ggml block types "
                                      f"({', '.join(sorted(BLOCK_BYTES))})")
        if len(info.shape) != 2:
            raise ValueError(f"GGUF tensor `{name}` must be a matrix, got shape {info.shape}")
        rows, cols = info.shape
        be, bb = BLOCK_ELEMS[info.dtype], BLOCK_BYTES[info.dtype]
        full = ar.tensor_data(name).reshape(rows, cols // be, bb)
        full, rows, cols = self._shard(full, rows, cols, be, kind)
        flat = np.ascontiguousarray(full).reshape(-1)
        t = torch.from_numpy(np.array(flat)).to(self.device)
        self.nbytes += t.numel()
        if self.host is not None:
            self.host[key] = np.array(flat)
        return (t, info.dtype, rows, cols)

    def _shard(self, full, rows, cols, be, kind):
        """kind: 'col' (rows sharded), 'kv' (rows sharded by KV head, replicated when ranks outnumber KV heads),
        'row' (K sharded on block boundaries), 'rep' (replicated)."""
        r, w = self.tp_rank, self.tp_size
        if kind == "kv" and w > 1:
            first, n = compute_kv_shard(self.cfg.n_kv_heads, self.cfg.head_dim, r, w)
            return full[first:first + n], n, cols
        if kind == "col" and w > 1:
            if rows % w:
                raise ValueError("column-parallel rows do not divide by the tensor-parallel size")
            return full[r * rows // w:(r + 1) * rows // w], rows // w, cols
        if kind == "row" and w > 1:
            nb = cols // be
            if nb % w:
                raise ValueError("row-parallel K does not split on block boundaries")
            return full[:, r * nb // w:(r + 1) * nb // w], rows, cols // w
        return full, rows, cols

    def _norm(self, layer, name):
        rng = np.random.Generator(np.random.PCG64(tensor_seed(layer, name)))
        w = (1.0 + 0.1 * rng.standard_normal(self.cfg.hidden)).astype(np.float32)
        t = torch.from_numpy(w).to(self.device).to(self.dtype)
        if self.host is not None:
            self.host[(layer, name)] = t.float().cpu().numpy()
        return t

    def _qtensor(self, layer, name, rows, cols, kind):
        """kind: 'col' (rows sharded), 'row' (K sharded on block boundaries), 'rep' (replicated).
        Sharding rules: REF mistralrs-quant/src/distributed/layers.rs:1167-1294 (column),
        :695-975 (row), gguf/weight_source.rs:809-818 (block-aligned K slices)."""
        dt = tensor_type(self.cfg, name, layer)
        be, bb = BLOCK_ELEMS[dt], BLOCK_BYTES[dt]
        if self.fast_synth:
            w = self.tp_size
            if kind == "kv" and w > 1:
                rows = compute_kv_shard(self.cfg.n_kv_heads, self.cfg.head_dim, self.tp_rank, w)[1]
            elif kind == "col" and w > 1:
                rows //= w
            elif kind == "row" and w > 1:
                assert (cols // be) % w == 0
                cols //= w
            nblocks = rows * cols // be
            gen = torch.Generator(device=self.device).manual_seed(tensor_seed(layer, name) * 64 + self.tp_rank)
            raw = torch.randint(0, 256, (nblocks, bb), dtype=torch.uint8, device=self.device, generator=gen)
            lo, hi = self.cfg.synth_scale_exp
            f = F16_FIELDS[dt]
            d = torch.exp2(torch.empty(nblocks, device=self.device).uniform_(lo, hi, generator=gen)).to(torch.float16)
            raw[:, f[0]:f[0] + 2] = d.view(torch.uint8).reshape(nblocks, 2)
            if len(f) > 1:
                m = (d.float() * torch.empty(nblocks, device=self.device).uniform_(0, 0.5, generator=gen)).to(torch.float16)
                raw[:, f[1]:f[1] + 2] = m.view(torch.uint8).reshape(nblocks, 2)
            t = raw.reshape(-1)
            self.nbytes += t.numel()
            return (t, dt, rows, cols)
        full = synth_blocks(dt, rows * cols // be, tensor_seed(layer, name), self.cfg.synth_scale_exp).reshape(rows, cols // be, bb)
        full, rows, cols = self._shard(full, rows, cols, be, kind)
        full = np.ascontiguousarray(full)
        t = torch.from_numpy(full.reshape(-1)).to(self.device)
        self.nbytes += t.numel()
        if self.host is not None:
            self.host[(layer, name)] = full.reshape(-1)
        return (t, dt, rows, cols)


def runner_split_pages(block_size, batch, n_kv_heads, max_ctx, sm_count=148, min_tokens=64):
    """Split-KV chunk (in pages) for the whole-token decode path: enough tiles that
    batch x kv_heads x tiles covers the SMs (one attention CTA per SM), chunks of at least
    `min_tokens`.  The reference's host policy (metadata.rs:61-86, kv_index.decode_split_pages) never
    goes below 256-token chunks, which leaves 16 CTAs for a batch-1 8-KV-head decode; the kernels
    accept any plan, this one only changes how finely the same sum is split."""
    tiles = max(1, sm_count // max(1, batch * n_kv_heads))
    tokens = max(min_tokens, -(-max_ctx // tiles))
    return max(1, -(-tokens // block_size))


class LlamaRunner:
    """Owns KV cache + scratch + per-step metadata for a batch of sequences and drives
    mrs_llama_decode_step / mrs_decode_advance (eagerly or as a captured CUDA graph)."""

    def __init__(self, weights: LlamaWeights, batch=1, max_ctx=512, pdl=False, sm_count=148, comm=None,
                 fused_attention=True, split_policy="sm_fill", split_min_tokens=64, peer_allreduce=None):
        cfg, dev, dt = weights.cfg, weights.device, weights.dtype
        self.w, self.cfg, self.dev, self.dt, self.B = weights, cfg, dev, dt, batch
        tp = weights.tp_size
        self.n_heads, self.n_kv = cfg.n_heads // tp, max(1, cfg.n_kv_heads // tp)   # (KV heads are replicated when tp > kv heads)
        bs = cfg.block_size
        self.max_blocks = -(-max_ctx // bs)
        nb = batch * self.max_blocks + 1
        self.pool = kv_index.BlockPool(nb)
        self.tables = [self.pool.get_new_blocks(self.max_blocks) for _ in range(batch)]
        self.block_tables = torch.tensor(self.tables, dtype=torch.int32, device=dev)
        self.context_lens = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)   # bit 0: a sequence ran out of context
        self.max_ctx = min(self.max_blocks * bs, cfg.max_pos)
        self.steps_taken = 0              # host-side count of advances since reset() (graph replays via replay())
        self.split_pages = (kv_index.decode_split_pages(bs, batch, self.n_kv, max_ctx, sm_count=sm_count)
                            if split_policy == "reference" else
                            runner_split_pages(bs, batch, self.n_kv, max_ctx, sm_count, split_min_tokens))
        self.padded_tiles = batch * -(-self.max_blocks // self.split_pages)
        if self.padded_tiles <= batch:        # a single chunk per request: unsplit plan
            self.split_pages, self.padded_tiles = 0, batch
        z = lambda *s, d=torch.int32: torch.zeros(*s, dtype=d, device=dev)
        self.meta = dict(token_ids=z(batch), positions=z(batch), slot_mapping=z(batch, d=torch.int64),
                         kv_indptr=z(batch + 1), kv_indices=z(batch * self.max_blocks), kv_last_page_len=z(batch),
                         request_indices=z(self.padded_tiles), kv_tile_indices=z(self.padded_tiles),
                         o_indptr=z(batch + 1), kv_chunk_size=z(1), block_valid_mask=z(self.padded_tiles, d=torch.uint8))
        a = lambda *s: torch.zeros(*s, dtype=dt, device=dev)
        H = cfg.hidden
        self.buf = dict(x=a(batch, H), x2=a(batch, H), q=a(batch, self.n_heads * cfg.head_dim),
                        k=a(batch, self.n_kv * cfg.head_dim), v=a(batch, self.n_kv * cfg.head_dim),
                        attn_out=a(batch, self.n_heads * cfg.head_dim), act=a(batch, cfg.inter // tp),
                        logits=a(batch, cfg.vocab), tmp_v=a(self.padded_tiles, self.n_heads, cfg.head_dim),
                        tmp_s=torch.zeros(self.padded_tiles, self.n_heads, dtype=torch.float32, device=dev),
                        out_token=self.meta["token_ids"],  # argmax feeds the next step directly
                        attn_counters=torch.zeros(batch * self.n_kv * 2, dtype=torch.int32, device=dev),
                        argmax_scratch=torch.zeros(16 * batch + 16, dtype=torch.uint8, device=dev))
        self.k_cache = [a(nb, self.n_kv, bs, cfg.head_dim) for _ in range(cfg.n_layers)]
        self.v_cache = [a(nb, self.n_kv, bs, cfg.head_dim) for _ in range(cfg.n_layers)]
        self._layers = (_Layer * cfg.n_layers)()
        for l, L in enumerate(weights.layers):
            for field_, name in (("wq", "attn_q"), ("wk", "attn_k"), ("wv", "attn_v"), ("wo", "attn_output"),
                                 ("w_gate", "ffn_gate"), ("w_up", "ffn_up"), ("w_down", "ffn_down")):
                t, ty, rows, cols = L[name]
                setattr(self._layers[l], field_, _QW(t.data_ptr(), GGML[ty], rows, cols))
            self._layers[l].attn_norm = L["attn_norm"].data_ptr()
            self._layers[l].ffn_norm = L["ffn_norm"].data_ptr()
            self._layers[l].k_cache = self.k_cache[l].data_ptr()
            self._layers[l].v_cache = self.v_cache[l].data_ptr()
        s = _Step()
        s.hidden, s.n_layers, s.n_heads, s.n_kv_heads, s.head_dim, s.vocab = H, cfg.n_layers, self.n_heads, self.n_kv, cfg.head_dim, cfg.vocab
        s.block_size, s.act_dtype = bs, {torch.float16: 0, torch.bfloat16: 1}[dt]
        s.rms_eps, s.sm_scale, s.rope_neox, s.pdl = cfg.rms_eps, 1.0 / float(np.sqrt(cfg.head_dim)), int(cfg.rope_neox), int(pdl)
        s.layers = ctypes.cast(self._layers, ctypes.POINTER(_Layer))
        t, ty, rows, cols = weights.tok_embd
        s.tok_embd = _QW(t.data_ptr(), GGML[ty], rows, cols)
        t, ty, rows, cols = weights.output
        s.lm_head = _QW(t.data_ptr(), GGML[ty], rows, cols)
        s.final_norm, s.rope_cos, s.rope_sin = weights.output_norm.data_ptr(), weights.rope_cos.data_ptr(), weights.rope_sin.data_ptr()
        s.batch, s.padded_tiles, s.max_blocks_per_seq = batch, self.padded_tiles, self.max_blocks
        s.fused_attention = int(fused_attention)
        for n, t in self.meta.items():
            setattr(s, n, t.data_ptr())
        for n, t in self.buf.items():
            setattr(s, n, t.data_ptr())
        self._ar_cb, self._peer = None, peer_allreduce
        if peer_allreduce is not None:     # in-graph peer-memory sum (takes precedence over the callback)
            s.tp = peer_allreduce.pointer()
        if comm is not None:
            self._ar_cb = _AR_FN(comm)
            s.all_reduce = self._ar_cb
        self.step_struct = s
        self.graph = None

    # ---- eager pieces -------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def advance(self):
        m = self.meta
        rc = lib().mrs_decode_advance(ctypes.c_void_p(self.block_tables.data_ptr()), ctypes.c_int(self.max_blocks),
                                      ctypes.c_void_p(self.context_lens.data_ptr()), ctypes.c_int(self.B),
                                      ctypes.c_int(self.cfg.block_size), ctypes.c_int(self.split_pages),
                                      ctypes.c_int(self.padded_tiles), *[ctypes.c_void_p(m[k].data_ptr()) for k in
                                      ("positions", "slot_mapping", "kv_indptr", "kv_indices", "kv_last_page_len",
                                       "request_indices", "kv_tile_indices", "o_indptr", "kv_chunk_size",
                                       "block_valid_mask")], ctypes.c_int(self.cfg.max_pos),
                                      ctypes.c_void_p(self.error_flag.data_ptr()), self._stream())
        assert rc == 0, rc
        self.steps_taken += 1

    def forward(self):
        rc = lib().mrs_llama_decode_step(ctypes.byref(self.step_struct), self._stream())
        if rc != 0:
            raise RuntimeError(f"mrs_llama_decode_step failed: cudaError {rc}")

    def step(self):
        """advance the KV metadata for the token in `token_ids`, run the stack, argmax -> token_ids."""
        if self.steps_taken >= self.max_ctx and not torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"LlamaRunner: context exhausted ({self.max_ctx} tokens: block table / RoPE table)")
        self.advance()
        self.forward()

    def reset(self, context_len=0):
        self.context_lens.fill_(context_len)
        self.error_flag.zero_()
        self.steps_taken = int(context_len)

    def replay(self):
        """one captured decode step; raises before a sequence would run past the allocated context
        (the kernels freeze such a sequence and set error_flag, but a caller should never get there)"""
        if self.steps_taken >= self.max_ctx:
            raise RuntimeError(f"LlamaRunner: context exhausted ({self.max_ctx} tokens: block table / RoPE table)")
        self.steps_taken += 1
        self.graph.replay()

    def check_overflow(self):
        if int(self.error_flag.item()) & 1:
            raise RuntimeError("LlamaRunner: a sequence ran past its allocated context (KV write skipped)")

    def capture(self):
        self.step(); self.reset()  # warm-up outside capture (module load, attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph(keep_graph=True)   # keep_graph: the kernel nodes can be counted afterwards
        with torch.cuda.graph(g):
            self.step()
        self.reset()
        self.graph = g
        return g

    def set_tokens(self, ids):
        self.meta["token_ids"].copy_(torch.as_tensor(ids, dtype=torch.int32, device=self.dev))

    def logits(self):
        return self.buf["logits"]


class LlamaPrefill:
    """Prompt processing (one sequence, positions 0..T-1) composed from the C-ABI ops, in the order
    of the reference's prefill forward (`models/llama.rs` Block::forward with seq_len > 1):
    RMSNorm -> tcgen05 dequant-GEMMs (`mmq.forward`, the `fast_mmq` path) -> RoPE -> causal prompt
    attention over the fresh q/k/v (`paged_attn.prefill_attention`, the reference's flash-attn call,
    paged_attention.rs:1413-1475) -> KV scatter into the paged HND cache -> o_proj -> add+RMSNorm ->
    GLU -> down -> add.  TTFT of BASELINE config 3 is the time of `forward` on a 4096-token prompt."""

    def __init__(self, weights: "LlamaWeights", max_tokens=4096, runner: "LlamaRunner" = None):
        """runner: write the prompt's K/V into sequence 0 of this decode runner's paged cache (its block
        table), so a generation is prefill -> `runner.reset(T)` -> decode-graph replays."""
        from . import ops, paged_attn, quant  # noqa: F401  (fail early when the extension is missing)
        cfg, dev, dt = weights.cfg, weights.device, weights.dtype
        if weights.tp_size != 1:
            raise NotImplementedError("LlamaPrefill runs single-GPU")
        self.w, self.cfg, self.dev, self.dt = weights, cfg, dev, dt
        bs = cfg.block_size
        self.max_tokens = int(max_tokens)
        self.nblocks = -(-self.max_tokens // bs)
        if runner is not None:
            if runner.max_blocks < self.nblocks:
                raise ValueError("LlamaPrefill: the runner's block table is shorter than max_tokens")
            self.table, self.k_cache, self.v_cache = list(runner.tables[0]), runner.k_cache, runner.v_cache
        else:
            nb = self.nblocks + 1                                   # block 0 stays the null block
            self.table = list(range(1, self.nblocks + 1))
            self.k_cache = [torch.zeros(nb, cfg.n_kv_heads, bs, cfg.head_dim, dtype=dt, device=dev) for _ in range(cfg.n_layers)]
            self.v_cache = [torch.zeros(nb, cfg.n_kv_heads, bs, cfg.head_dim, dtype=dt, device=dev) for _ in range(cfg.n_layers)]
        self._slots = torch.tensor([self.table[i // bs] * bs + i % bs for i in range(self.max_tokens)], dtype=torch.int64, device=dev)

    def forward(self, tokens, all_logits=False):
        """tokens: list[int] (1 < len <= max_tokens).  Returns logits [vocab] of the last token, or
        [T, vocab] with all_logits=True.  The KV cache of every layer holds rows 0..T-1 afterwards."""
        from . import mmq, ops, paged_attn, quant
        cfg, dev, dt, w = self.cfg, self.dev, self.dt, self.w
        T = len(tokens)
        if not 1 < T <= min(self.max_tokens, cfg.max_pos):
            raise ValueError(f"LlamaPrefill.forward: need 1 < tokens <= {min(self.max_tokens, cfg.max_pos)}, got {T}")
        H, KVH, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        slots = self._slots[:T]
        if torch.is_tensor(tokens):      # e.g. pinned host ids: the H2D copy is part of the call
            ids = tokens.to(device=dev, dtype=torch.int32, non_blocking=True)
        else:
            ids = torch.tensor(tokens, dtype=torch.int32, device=dev)
        x = torch.empty(T, cfg.hidden, dtype=dt, device=dev)
        t, ty, rows, cols = w.tok_embd
        rc = lib().mrs_embedding_gather(ctypes.c_int32(GGML[ty]), ctypes.c_void_p(t.data_ptr()), ctypes.c_int32(cols),
                                        ctypes.c_void_p(ids.data_ptr()), ctypes.c_int32(T), ctypes.c_void_p(x.data_ptr()),
                                        ctypes.c_int32({torch.float16: 0, torch.bfloat16: 1}[dt]),
                                        ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"mrs_embedding_gather failed with cudaError {rc}")
        qt = lambda e: quant.QTensor(e[0], e[1], (e[2], e[3]))
        scale = 1.0 / float(np.sqrt(D))
        h = ops.rms_norm(x, w.layers[0]["attn_norm"], cfg.rms_eps)
        for l, L in enumerate(w.layers):
            q = mmq.forward(qt(L["attn_q"]), h).view(T, H, D)
            k = mmq.forward(qt(L["attn_k"]), h).view(T, KVH, D)
            v = mmq.forward(qt(L["attn_v"]), h).view(T, KVH, D)
            ops.apply_rotary_qk(q, k, w.rope_cos, w.rope_sin, None, is_neox=cfg.rope_neox)
            attn = paged_attn.prefill_attention(q, k, v, scale)
            paged_attn.reshape_and_cache_flashinfer(k, v, self.k_cache[l], self.v_cache[l], slots)
            o = mmq.forward(qt(L["attn_output"]), attn.view(T, H * D))
            x, h2 = ops.add_rms_norm(o, x, L["ffn_norm"], cfg.rms_eps)       # x = o + x ; h2 = norm(x)
            gate = mmq.forward(qt(L["ffn_gate"]), h2)
            up = mmq.forward(qt(L["ffn_up"]), h2)
            d = mmq.forward(qt(L["ffn_down"]), ops.fused_glu(gate, up, 0))
            nxt = w.layers[l + 1]["attn_norm"] if l + 1 < cfg.n_layers else w.output_norm
            x, h = ops.add_rms_norm(d, x, nxt, cfg.rms_eps)                   # x = d + x ; h = next norm(x)
        if all_logits:
            return mmq.forward(qt(w.output), h)
        return quant.plain(qt(w.output), h[T - 1:T].contiguous())[0]
