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


# ---- UQFF artifacts (safetensors shards following REF docs/.../reference/uqff-format.md) -------------
UQFF_KEYS = {"attn_q": "self_attn.q_proj", "attn_k": "self_attn.k_proj", "attn_v": "self_attn.v_proj",
             "attn_output": "self_attn.o_proj", "ffn_gate": "mlp.gate_proj", "ffn_up": "mlp.up_proj",
             "ffn_down": "mlp.down_proj"}
GGML_CODE = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8, "q2_k": 10, "q3_k": 11, "q4_k": 12, "q5_k": 13,
             "q6_k": 14}


def uqff_layer_entries(key, dtype, rows, cols, blocks):
    """the four entries of a GGML-family layer"""
    return {f"{key}.weight": np.ascontiguousarray(blocks).reshape(-1).astype(np.uint8),
            f"{key}.weight.format": np.array(0, dtype=np.uint8),
            f"{key}.weight.dtype": np.array(GGML_CODE[dtype], dtype=np.uint32),
            f"{key}.weight.shape": np.array([rows, cols], dtype=np.uint32)}


def uqff_version_entries(version=(1, 2, 0)):
    return {f"uqff.version.{n}": np.array(v, dtype=np.uint32) for n, v in zip(("major", "minor", "patch"), version)}


def write_llama_uqff(dirpath, cfg, n_shards=2, version=(1, 2, 0), tie=False):
    """Write the synthetic llama model of model.py as a UQFF artifact directory: `q-<n>.uqff` shards,
    residual.safetensors (norms, bf16) and config.json.  Returns name -> blocks / f32 norm arrays."""
    import json
    import os

    import torch
    from safetensors.numpy import save_file
    from safetensors.torch import save_file as save_torch
    t = llama_tensors(cfg)
    layers, resid, src = {}, {}, {}

    def q(key, gname):
        _, dt, rows, cols, blocks = t[gname]
        layers[key] = uqff_layer_entries(key, dt, rows, cols, blocks)
        src[key] = (dt, rows, cols, blocks)

    def norm(name, gname):
        resid[name] = torch.from_numpy(t[gname][1]).to(torch.bfloat16)
        src[name] = t[gname][1]
    q("model.embed_tokens", "token_embd.weight")
    for l in range(cfg.n_layers):
        for g, h in UQFF_KEYS.items():
            q(f"model.layers.{l}.{h}", f"blk.{l}.{g}.weight")
        norm(f"model.layers.{l}.input_layernorm.weight", f"blk.{l}.attn_norm.weight")
        norm(f"model.layers.{l}.post_attention_layernorm.weight", f"blk.{l}.ffn_norm.weight")
    norm("model.norm.weight", "output_norm.weight")
    if not tie:
        q("lm_head", "output.weight")
    os.makedirs(dirpath, exist_ok=True)
    keys = list(layers)
    per = -(-len(keys) // n_shards)
    paths = []
    for s in range(n_shards):
        entries = dict(uqff_version_entries(version))
        for k in keys[s * per:(s + 1) * per]:
            entries.update(layers[k])
        p = os.path.join(dirpath, f"q-{s}.uqff")
        save_file(entries, p, metadata={"uqff.producer": "mrs-b200 tests", "uqff.producer.mistralrs.version": "test"})
        paths.append(p)
    save_torch(resid, os.path.join(dirpath, "residual.safetensors"))
    json.dump({"model_type": "llama", "hidden_size": cfg.hidden, "intermediate_size": cfg.inter,
               "num_hidden_layers": cfg.n_layers, "num_attention_heads": cfg.n_heads,
               "num_key_value_heads": cfg.n_kv_heads, "head_dim": cfg.head_dim, "vocab_size": cfg.vocab,
               "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta, "max_position_embeddings": cfg.max_pos,
               "tie_word_embeddings": tie}, open(os.path.join(dirpath, "config.json"), "w"))
    return paths, src
