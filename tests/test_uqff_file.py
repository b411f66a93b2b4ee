"""UQFF artifacts: the C++ safetensors container reader against the `safetensors` package (the
independent implementation of the container), and the UQFF layer conventions / version rules of the
reference reader (REF docs/.../reference/uqff-format.md, mistralrs-quant/src/uqff/reader.rs:81-176).
CPU only; bytes and integers bit-exact."""
import json
import os
import struct

import numpy as np
import pytest
import torch
from safetensors import safe_open
from safetensors.numpy import save_file

from gguf_util import uqff_layer_entries, uqff_version_entries, write_llama_uqff
from mistralrs_b200 import BLOCK_BYTES, BLOCK_ELEMS, model as M, uqff_file


@pytest.fixture(scope="module")
def cfg():
    return M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=2)


@pytest.fixture(scope="module")
def artifact(tmp_path_factory, cfg):
    d = str(tmp_path_factory.mktemp("uqff"))
    paths, src = write_llama_uqff(d, cfg, n_shards=3)
    return d, paths, src


def test_container_matches_safetensors_package(tmp_path):
    rng = np.random.default_rng(3)
    tensors = {"a.f32": rng.standard_normal((3, 5)).astype(np.float32), "b.u8": rng.integers(0, 256, 1000, dtype=np.uint8),
               "c.i64": np.arange(7, dtype=np.int64), "scalar.u32": np.array(42, dtype=np.uint32),
               "empty": np.zeros((0, 4), dtype=np.float16), "ünï.f16": rng.standard_normal(9).astype(np.float16)}
    p = str(tmp_path / "t.safetensors")
    save_file(tensors, p, metadata={"k": "v", "quote": 'a "b" \\ c', "uni": "é中"})
    f = uqff_file.SafetensorsFile(p)
    with safe_open(p, framework="np") as ref:
        assert set(f.entries) == set(ref.keys()) == set(tensors)
        assert f.metadata == ref.metadata()
        for name in ref.keys():
            want = ref.get_tensor(name)
            dt, shape, off, nb = f.entries[name]
            assert shape == want.shape and nb == want.nbytes
            assert np.array_equal(f.array(name), want)
    raw = open(p, "rb").read()
    hlen = struct.unpack("<Q", raw[:8])[0]
    for name, (dt, shape, off, nb) in f.entries.items():   # offsets are absolute: header + data_offsets
        hdr = json.loads(raw[8:8 + hlen])
        assert off == 8 + hlen + hdr[name]["data_offsets"][0]
    f.close()


def test_container_rejects_malformed_files(tmp_path):
    p = str(tmp_path / "ok.safetensors")
    save_file({"x": np.arange(16, dtype=np.float32), "y": np.arange(4, dtype=np.uint8)}, p)
    raw = open(p, "rb").read()
    hlen = struct.unpack("<Q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])

    def rewrite(name, h, data=None):
        hb = json.dumps(h).encode()
        q = str(tmp_path / name)
        open(q, "wb").write(struct.pack("<Q", len(hb)) + hb + (raw[8 + hlen:] if data is None else data))
        return q
    with pytest.raises(ValueError, match="header length exceeds"):
        q = str(tmp_path / "len.safetensors")
        open(q, "wb").write(struct.pack("<Q", 10 ** 12) + raw[8:])
        uqff_file.SafetensorsFile(q)
    bad = json.loads(json.dumps(hdr)); bad["x"]["data_offsets"][1] += 4
    with pytest.raises(ValueError, match="its shape needs|extends past|hole or overlap"):
        uqff_file.SafetensorsFile(rewrite("off.safetensors", bad))
    bad = json.loads(json.dumps(hdr)); bad["x"]["dtype"] = "Q4"
    with pytest.raises(ValueError, match="unknown dtype"):
        uqff_file.SafetensorsFile(rewrite("dt.safetensors", bad))
    bad = json.loads(json.dumps(hdr)); bad["y"]["data_offsets"] = bad["x"]["data_offsets"][:1] + [bad["x"]["data_offsets"][0] + 4]
    with pytest.raises(ValueError, match="hole or overlap|trailing"):
        uqff_file.SafetensorsFile(rewrite("ov.safetensors", bad))
    with pytest.raises(ValueError, match="trailing bytes"):
        uqff_file.SafetensorsFile(rewrite("tail.safetensors", hdr, raw[8 + hlen:] + b"\0" * 8))
    with pytest.raises(ValueError, match="JSON"):
        q = str(tmp_path / "json.safetensors")
        open(q, "wb").write(struct.pack("<Q", 5) + b"{oops" + raw[8 + hlen:])
        uqff_file.SafetensorsFile(q)
    with pytest.raises(ValueError, match="cannot open"):
        uqff_file.SafetensorsFile(str(tmp_path / "missing.safetensors"))


def test_layer_catalogue_and_blocks(artifact, cfg):
    d, paths, src = artifact
    with uqff_file.UqffArchive(d) as ar:
        assert ar.version == (1, 2, 0) and len(ar.shards) == 3
        assert ar.residual is not None and ar.config["hidden_size"] == cfg.hidden
        quant_keys = [k for k, v in src.items() if isinstance(v, tuple)]
        assert ar.layer_keys() == sorted(quant_keys)
        for key in quant_keys:
            dt, rows, cols, blocks = src[key]
            info = ar.layer_info(key)
            assert (info.format, info.dtype, info.shape, info.has_bias) == ("gguf", dt, (rows, cols), False)
            q = ar.load_qtensor(key, "cpu")
            assert q.dtype == dt and tuple(q.shape) == (rows, cols)
            assert np.array_equal(q.data.numpy(), blocks.reshape(-1))
        n = ar.load_tensor("model.norm.weight", "cpu")
        assert n.dtype == torch.bfloat16
        assert torch.equal(n.float(), torch.from_numpy(src["model.norm.weight"]).to(torch.bfloat16).float())
        with pytest.raises(KeyError, match="cannot find UQFF tensor"):
            ar.layer_info("model.layers.9.mlp.up_proj")
    # explicit shard list in any order gives the same catalogue
    with uqff_file.UqffArchive(list(reversed(paths))) as ar2:
        assert ar2.layer_keys() == sorted(quant_keys)


def test_version_rules(tmp_path):
    # REF reader.rs:81-176: scalar U32 version entries are mandatory; a different major or a newer
    # minor is rejected; conflicting copies across shards are rejected
    layer = uqff_layer_entries("l", "q8_0", 2, 32, np.zeros((2, 34), dtype=np.uint8))

    def art(name, entries_list):
        d = tmp_path / name
        d.mkdir()
        for i, e in enumerate(entries_list):
            save_file(e, str(d / f"q-{i}.uqff"))
        return str(d)
    assert uqff_file.UqffArchive(art("ok", [{**layer, **uqff_version_entries((1, 0, 7))}])).version == (1, 0, 7)
    with pytest.raises(ValueError, match="no version tag"):
        uqff_file.UqffArchive(art("nov", [layer]))
    with pytest.raises(ValueError, match="incompatible with this build"):
        uqff_file.UqffArchive(art("maj", [{**layer, **uqff_version_entries((2, 0, 0))}]))
    with pytest.raises(ValueError, match="written by a newer"):
        uqff_file.UqffArchive(art("min", [{**layer, **uqff_version_entries((1, 3, 0))}]))
    with pytest.raises(ValueError, match="Conflicting UQFF version tensor"):
        uqff_file.UqffArchive(art("conf", [{**layer, **uqff_version_entries((1, 2, 0))},
                                           {"m.weight": np.zeros(4, dtype=np.uint8), **uqff_version_entries((1, 1, 0))}]))
    bad = {**layer, **uqff_version_entries((1, 2, 0))}
    bad["uqff.version.minor"] = np.array([2], dtype=np.uint32)       # a vector, not a scalar
    with pytest.raises(ValueError, match="must be a scalar U32"):
        uqff_file.UqffArchive(art("vec", [bad]))
    with pytest.raises(ValueError, match="duplicated across shards"):
        uqff_file.UqffArchive(art("dup", [{**layer, **uqff_version_entries()}, {**layer, **uqff_version_entries()}]))


def test_non_gguf_families_are_reported_not_misread(tmp_path):
    d = tmp_path / "afq"
    d.mkdir()
    e = {"l.weight": np.zeros((4, 8), dtype=np.uint32), "l.weight.format": np.array(4, dtype=np.uint8),
         "l.weight.bits": np.array(4, dtype=np.uint8), "l.weight.group_size": np.array(64, dtype=np.uint8),
         **uqff_version_entries()}
    save_file(e, str(d / "afq4-0.uqff"))
    with uqff_file.UqffArchive(str(d)) as ar:
        assert ar.layer_info("l").format == "afq"
        with pytest.raises(NotImplementedError, match="afq family"):
            ar.load_qtensor("l", "cpu")


def test_llama_weights_from_uqff_cpu(artifact, cfg):
    d, paths, src = artifact
    with uqff_file.UqffArchive(d) as ar:
        w = M.LlamaWeights.from_uqff(ar, "cpu", keep_host=True)
        c = w.cfg
        for f in ("hidden", "inter", "n_layers", "n_heads", "n_kv_heads", "head_dim", "vocab", "max_pos"):
            assert getattr(c, f) == getattr(cfg, f), f
        assert c.rope_neox is True and c.quant == "uqff"
        for l in range(cfg.n_layers):
            for name, hf in M.LlamaWeights.UQFF_NAMES.items():
                t, ty, rows, cols = w.layers[l][name]
                dt, r, cc, blocks = src[f"model.layers.{l}.{hf}"]
                assert (ty, rows, cols) == (dt, r, cc) and np.array_equal(t.numpy(), blocks.reshape(-1))
        assert np.array_equal(w.output[0].numpy(), src["lm_head"][3].reshape(-1))
        assert np.array_equal(w.tok_embd[0].numpy(), src["model.embed_tokens"][3].reshape(-1))
        ws = M.LlamaWeights.from_uqff(ar, "cpu", tp_rank=1, tp_size=2)
        t, ty, rows, cols = ws.layers[0]["attn_output"]
        dt, r, cc, blocks = src["model.layers.0.self_attn.o_proj"]
        be, bb = BLOCK_ELEMS[dt], BLOCK_BYTES[dt]
        full = blocks.reshape(r, cc // be, bb)
        nb = cc // be
        assert (rows, cols) == (r, cc // 2)
        assert np.array_equal(t.numpy(), np.ascontiguousarray(full[:, nb // 2:]).reshape(-1))


def test_tied_and_llama3_rope_scaling(tmp_path, cfg):
    d = str(tmp_path / "tied")
    write_llama_uqff(d, cfg, n_shards=1, tie=True)
    cj = json.load(open(os.path.join(d, "config.json")))
    cj["rope_scaling"] = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                          "original_max_position_embeddings": 128}
    json.dump(cj, open(os.path.join(d, "config.json"), "w"))
    with uqff_file.UqffArchive(d) as ar:
        w = M.LlamaWeights.from_uqff(ar, "cpu")
        assert torch.equal(w.output[0], w.tok_embd[0])
        assert w.cfg.rope_scaling["factor"] == 8.0
        plain = M.rope_tables(M.LlamaWeights.config_from_hf({**cj, "rope_scaling": None}))[0]
        assert not np.array_equal(M.rope_tables(w.cfg)[0], plain)      # scaled frequencies differ


def test_layer_serialize_roundtrip_and_layout(tmp_path):   # gguf/mod.rs:755-806 serialize_uqff / deserialize_uqff
    import oracle
    from safetensors.numpy import load_file
    from mistralrs_b200 import quant, uqff_file
    rng = np.random.default_rng(4)
    n, k = 24, 512
    blocks = oracle.random_blocks("q6_k", n * k // 256, rng)
    bias = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(torch.bfloat16)
    layer = quant.GgufMatMul(quant.QTensor(torch.from_numpy(blocks.reshape(-1)), "q6_k", (n, k)), bias)
    key = "model.layers.3.mlp.down_proj"
    entries = layer.serialize_uqff(key)
    want = uqff_layer_entries(key, "q6_k", n, k, blocks)          # the documented four entries
    assert set(entries) == set(want) | {f"{key}.bias"}
    for name, a in want.items():
        assert entries[name].dtype == a.dtype and entries[name].shape == a.shape and np.array_equal(entries[name], a), name
    plain = quant.GgufMatMul(quant.QTensor(torch.from_numpy(blocks.reshape(-1)), "q6_k", (n, k)))
    other = plain.serialize_uqff("lm_head")
    assert "lm_head.bias" not in other
    path = tmp_path / "q-0.uqff"
    uqff_file.save_uqff(path, {**entries, **other}, metadata={"writer": "test"})
    # an independent reader agrees on every tensor (bf16 bias aside, which numpy cannot represent)
    try:
        ref = load_file(str(path))
    except (TypeError, ValueError):
        ref = None
    if ref is not None:
        assert np.array_equal(ref[f"{key}.weight"], blocks.reshape(-1)) and int(ref["uqff.version.minor"]) == uqff_file.UQFF_VERSION[1]
    with uqff_file.UqffArchive([path]) as ar:
        assert ar.version == uqff_file.UQFF_VERSION and ar.layer_keys() == ["lm_head", key]
        back = quant.GgufMatMul.deserialize_uqff(ar, key, "cpu")
        assert back.w.dtype == "q6_k" and tuple(back.w.shape) == (n, k) and torch.equal(back.w.data, layer.w.data)
        assert back.b.dtype == torch.bfloat16 and torch.equal(back.b, bias)
        assert quant.GgufMatMul.deserialize_uqff(ar, "lm_head", "cpu").b is None
    with pytest.raises(ValueError):
        uqff_file.save_uqff(tmp_path / "bad.uqff", {"x": np.zeros(2, dtype=np.complex64)})
