"""Test helper: write llama-architecture GGUF files with gguf-py (the independent implementation
the reader is checked against)."""
import numpy as np

import gguf
from gguf import GGMLQuantizationType as QT

from mistralrs_b200 import BLOCK_BYTES, BLOCK_ELEMS, model as M

QTYPE = {"q4_0": QT.Q4_0, "q4_1": QT.Q4_1, "q5_0": QT.Q5_0, "q5_1": QT.Q5_1, "q8_0": QT.Q8_0, "q2_k": QT.Q2_K,
         "q3_k": QT.Q3_K, "q4_k": QT.Q4_K, "q5_k": QT.Q5_K, "q6_k": QT.Q6_K}


def add_blocks(w, name, dtype, rows, cols, blocks):
    """blocks: uint8 [rows * cols / block_elems, block_bytes] -> GGUF tensor [rows, cols] of `dtype`."""
    be, bb = BLOCK_ELEMS[dtype], BLOCK_BYTES[dtype]
    w.add_tensor(name, np.ascontiguousarray(blocks).reshape(rows, cols // be * bb), raw_dtype=QTYPE[dtype])


def llama_tensors(cfg):
    """name -> ('q', dtype, rows, cols, blocks) | ('f32', array): the synthetic model of model.py."""
    H, I = cfg.hidden, cfg.inter
    nq, nkv = cfg.n_heads * cfg.head_dim, cfg.n_kv_heads * cfg.head_dim
    out = {}

    def q(gname, layer, name, rows, cols):
        dt = M.tensor_type(cfg, name, layer)
        out[gname] = ("q", dt, rows, cols, M.synth_blocks(dt, rows * cols // BLOCK_ELEMS[dt], M.tensor_seed(layer, name)))

    def norm(gname, layer, name):
        rng = np.random.Generator(np.random.PCG64(M.tensor_seed(layer, name)))
        out[gname] = ("f32", (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32))
    q("token_embd.weight", 0, "token_embd", cfg.vocab, H)
    for l in range(cfg.n_layers):
        for name, rows, cols in (("attn_q", nq, H), ("attn_k", nkv, H), ("attn_v", nkv, H), ("attn_output", H, nq),
                                 ("ffn_gate", I, H), ("ffn_up", I, H), ("ffn_down", H, I)):
            q(f"blk.{l}.{name}.weight", l, name, rows, cols)
        norm(f"blk.{l}.attn_norm.weight", l, "attn_norm")
        norm(f"blk.{l}.ffn_norm.weight", l, "ffn_norm")
    norm("output_norm.weight", 0, "output_norm")
    q("output.weight", 0, "output", cfg.vocab, H)
    return out


def write_llama_gguf(path, cfg, tensors=None, names=None, split=None, alignment=None, extra_meta=True):
    """Write (a shard of) the synthetic llama model.  names: subset of tensor names for this file;
    split: (no, count, total_tensors)."""
    tensors = tensors if tensors is not None else llama_tensors(cfg)
    w = gguf.GGUFWriter(path, "llama")
    if alignment is not None:
        w.add_custom_alignment(alignment)
    w.add_string("general.name", cfg.name)
    w.add_uint32("llama.embedding_length", cfg.hidden)
    w.add_uint32("llama.feed_forward_length", cfg.inter)
    w.add_uint32("llama.block_count", cfg.n_layers)
    w.add_uint32("llama.attention.head_count", cfg.n_heads)
    w.add_uint32("llama.attention.head_count_kv", cfg.n_kv_heads)
    w.add_uint32("llama.rope.dimension_count", cfg.head_dim)
    w.add_uint32("llama.context_length", cfg.max_pos)
    w.add_float32("llama.attention.layer_norm_rms_epsilon", cfg.rms_eps)
    w.add_float32("llama.rope.freq_base", cfg.rope_theta)
    if extra_meta:
        w.add_bool("test.flag", True)
        w.add_int32("test.negative", -7)
        w.add_uint64("test.big", 2 ** 40 + 3)
        w.add_float64("test.pi", 3.141592653589793)
        w.add_array("tokenizer.ggml.tokens", [f"tok{i}" for i in range(16)] + ["ünï", ""])
        w.add_array("tokenizer.ggml.scores", [float(-i) * 0.5 for i in range(18)])
        w.add_array("tokenizer.ggml.token_type", [1, 2, 3, 1, 1, 6])
    if split is not None:
        no, count, total = split
        w.add_uint16("split.no", no)
        w.add_uint16("split.count", count)
        w.add_int32("split.tensors.count", total)
    for name in (names if names is not None else list(tensors)):
        t = tensors[name]
        if t[0] == "q":
            add_blocks(w, name, t[1], t[2], t[3], t[4])
        else:
            w.add_tensor(name, t[1])
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return tensors
