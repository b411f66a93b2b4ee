"""GGUF archive reader (host/gguf_reader.hpp + gguf_file.py) against gguf-py 0.19's writer/reader —
the independent implementation of the format — and the reference archive layer's error behaviour
(REF mistralrs-quant/src/gguf/archive.rs).  CPU only; bytes and integers bit-exact."""
import struct

import gguf
import numpy as np
import pytest
import torch

from gguf_util import llama_tensors, write_llama_gguf
from mistralrs_b200 import BLOCK_BYTES, BLOCK_ELEMS, gguf_file, model as M


@pytest.fixture(scope="module")
def cfg():
    return M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=2)


@pytest.fixture(scope="module")
def tiny(tmp_path_factory, cfg):
    path = str(tmp_path_factory.mktemp("gguf") / "tiny.gguf")
    tensors = write_llama_gguf(path, cfg)
    return path, tensors


def test_catalogue_matches_gguf_py(tiny):
    path, tensors = tiny
    ref = gguf.GGUFReader(path)
    with gguf_file.GgufArchive(path) as ar:
        assert ar.alignment == 32
        assert set(ar.tensors()) == {t.name for t in ref.tensors} == set(tensors)
        for rt in ref.tensors:
            info = ar.tensor_info(rt.name)
            assert info.ggml_type == int(rt.tensor_type)
            assert info.dims == tuple(int(d) for d in rt.shape)          # ggml order, innermost first
            assert info.shape == tuple(reversed(info.dims))
            assert info.offset == int(rt.data_offset) and info.offset % ar.alignment == 0
            assert info.nbytes == int(rt.n_bytes)
            assert np.array_equal(ar.tensor_data(rt.name), np.asarray(rt.data).view(np.uint8).reshape(-1))
        with pytest.raises(KeyError, match="cannot find GGUF tensor"):
            ar.tensor_info("blk.99.attn_q.weight")


def test_tensor_bytes_are_the_blocks_we_wrote(tiny):
    path, tensors = tiny
    with gguf_file.GgufArchive(path) as ar:
        for name, t in tensors.items():
            if t[0] == "q":
                assert ar.tensor_info(name).dtype == t[1]
                assert ar.tensor_info(name).shape == (t[2], t[3])
                assert np.array_equal(ar.tensor_data(name), t[4].reshape(-1))
                q = ar.load_qtensor(name, "cpu")
                assert q.dtype == t[1] and tuple(q.shape) == (t[2], t[3])
                assert q.data.numel() == t[2] * t[3] // BLOCK_ELEMS[t[1]] * BLOCK_BYTES[t[1]]
                with pytest.raises(ValueError, match="load_qtensor"):
                    ar.load_dense(name, "cpu")
            else:
                assert torch.equal(ar.load_dense(name, "cpu"), torch.from_numpy(t[1]))
                with pytest.raises(ValueError, match="load_dense"):
                    ar.load_qtensor(name, "cpu")


def test_metadata_all_value_types(tiny, cfg):
    path, _ = tiny
    with gguf_file.GgufArchive(path) as ar:
        md = ar.metadata()
    assert md["general.architecture"] == "llama" and md["general.name"] == cfg.name
    assert md["llama.embedding_length"] == cfg.hidden and md["llama.block_count"] == cfg.n_layers
    assert md["llama.rope.freq_base"] == pytest.approx(cfg.rope_theta)
    assert md["test.flag"] is True and md["test.negative"] == -7 and md["test.big"] == 2 ** 40 + 3
    assert md["test.pi"] == 3.141592653589793
    assert md["tokenizer.ggml.tokens"] == [f"tok{i}" for i in range(16)] + ["ünï", ""]
    assert md["tokenizer.ggml.scores"] == [float(-i) * 0.5 for i in range(18)]
    assert md["tokenizer.ggml.token_type"] == [1, 2, 3, 1, 1, 6]
    # every key gguf-py sees, we see
    ref = gguf.GGUFReader(path)
    assert set(ref.fields) - {"GGUF.version", "GGUF.tensor_count", "GGUF.kv_count"} == set(md)


def test_custom_alignment(tmp_path, cfg):
    path = str(tmp_path / "a64.gguf")
    write_llama_gguf(path, cfg, alignment=64, extra_meta=False)
    ref = gguf.GGUFReader(path)
    with gguf_file.GgufArchive(path) as ar:
        assert ar.alignment == 64
        for rt in ref.tensors:
            assert ar.tensor_info(rt.name).offset == int(rt.data_offset)
            assert ar.tensor_info(rt.name).offset % 64 == 0


def test_split_shards_any_order_and_errors(tmp_path, cfg):
    tensors = llama_tensors(cfg)
    names = list(tensors)
    a, b = names[: len(names) // 2], names[len(names) // 2:]
    p0, p1 = str(tmp_path / "m-00001-of-00002.gguf"), str(tmp_path / "m-00002-of-00002.gguf")
    write_llama_gguf(p0, cfg, tensors, a, split=(0, 2, len(names)))
    write_llama_gguf(p1, cfg, tensors, b, split=(1, 2, len(names)), extra_meta=False)
    for order in ([p0, p1], [p1, p0]):
        with gguf_file.GgufArchive(order) as ar:
            assert set(ar.tensors()) == set(names)
            assert {ar.tensor_info(n).shard for n in a} == {0} and {ar.tensor_info(n).shard for n in b} == {1}
            for n in (a[1], b[0], b[-1]):
                t = tensors[n]
                want = t[4].reshape(-1) if t[0] == "q" else t[1].view(np.uint8)
                assert np.array_equal(ar.tensor_data(n), want)
            assert ar.metadata()["test.big"] == 2 ** 40 + 3       # model metadata comes from shard 0
    with pytest.raises(ValueError, match="declares 2 shards, but 1"):
        gguf_file.GgufArchive(p0)
    with pytest.raises(ValueError, match="appears twice"):
        gguf_file.GgufArchive([p0, p0])
    p1dup = str(tmp_path / "dup.gguf")
    write_llama_gguf(p1dup, cfg, tensors, a[:1] + b, split=(1, 2, len(names)), extra_meta=False)
    with pytest.raises(ValueError, match="duplicated across shards"):
        gguf_file.GgufArchive([p0, p1dup])
    p1short = str(tmp_path / "short.gguf")
    write_llama_gguf(p1short, cfg, tensors, b[:-1], split=(1, 2, len(names)), extra_meta=False)
    with pytest.raises(ValueError, match=f"declares {len(names)} tensors, but {len(names) - 1}"):
        gguf_file.GgufArchive([p0, p1short])


def test_malformed_files(tmp_path, tiny):
    path, _ = tiny
    raw = open(path, "rb").read()

    def variant(name, data):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        return p
    with pytest.raises(ValueError, match="bad magic"):
        gguf_file.GgufArchive(variant("magic.gguf", b"GGML" + raw[4:]))
    with pytest.raises(ValueError, match="unsupported GGUF version 1"):
        gguf_file.GgufArchive(variant("v1.gguf", raw[:4] + struct.pack("<I", 1) + raw[8:]))
    with pytest.raises(ValueError, match="big-endian"):
        gguf_file.GgufArchive(variant("be.gguf", b"FUGG" + raw[4:]))
    with pytest.raises(ValueError, match="truncated"):
        gguf_file.GgufArchive(variant("cut_meta.gguf", raw[:200]))
    with pytest.raises(ValueError, match="extends past the end"):
        gguf_file.GgufArchive(variant("cut_data.gguf", raw[: len(raw) - 1000]))
    with pytest.raises(ValueError, match="too small"):
        gguf_file.GgufArchive(variant("tiny.gguf", raw[:10]))
    with pytest.raises(ValueError, match="cannot open"):
        gguf_file.GgufArchive(str(tmp_path / "missing.gguf"))


def test_llama_weights_from_gguf_cpu(tiny, cfg):
    # config + every tensor of the runner's weight structure, incl. tensor-parallel shards cut on
    # block boundaries exactly like the synthetic model's (model.LlamaWeights._qtensor)
    path, tensors = tiny
    with gguf_file.GgufArchive(path) as ar:
        c = M.LlamaWeights.config_from_gguf(ar)
        for f in ("hidden", "inter", "n_layers", "n_heads", "n_kv_heads", "head_dim", "vocab", "max_pos"):
            assert getattr(c, f) == getattr(cfg, f), f
        assert c.rms_eps == pytest.approx(cfg.rms_eps) and c.rope_theta == pytest.approx(cfg.rope_theta)
        assert c.rope_neox is False and c.quant == "gguf"
        w = M.LlamaWeights.from_gguf(ar, "cpu", keep_host=True)
        for l in range(cfg.n_layers):
            for name in M.LlamaWeights.GGUF_NAMES:
                t, ty, rows, cols = w.layers[l][name]
                src = tensors[f"blk.{l}.{name}.weight"]
                assert (ty, rows, cols) == (src[1], src[2], src[3])
                assert np.array_equal(t.numpy(), src[4].reshape(-1))
            assert torch.equal(w.layers[l]["attn_norm"].float(),
                               torch.from_numpy(tensors[f"blk.{l}.attn_norm.weight"][1]).to(torch.bfloat16).float())
        assert np.array_equal(w.output[0].numpy(), tensors["output.weight"][4].reshape(-1))
        assert w.rope_cos.shape == (cfg.max_pos, cfg.head_dim // 2)
        # TP=2: column shards = row halves, row shards = halves of every row's blocks
        for rank in (0, 1):
            ws = M.LlamaWeights.from_gguf(ar, "cpu", tp_rank=rank, tp_size=2)
            t, ty, rows, cols = ws.layers[1]["ffn_gate"]
            src = tensors["blk.1.ffn_gate.weight"]
            be, bb = BLOCK_ELEMS[ty], BLOCK_BYTES[ty]
            full = src[4].reshape(src[2], src[3] // be, bb)
            assert (rows, cols) == (src[2] // 2, src[3])
            assert np.array_equal(t.numpy(), full[rank * rows:(rank + 1) * rows].reshape(-1))
            t, ty, rows, cols = ws.layers[1]["ffn_down"]
            src = tensors["blk.1.ffn_down.weight"]
            be, bb = BLOCK_ELEMS[ty], BLOCK_BYTES[ty]
            full = src[4].reshape(src[2], src[3] // be, bb)
            nb = src[3] // be
            assert (rows, cols) == (src[2], src[3] // 2)
            assert np.array_equal(t.numpy(), np.ascontiguousarray(full[:, rank * nb // 2:(rank + 1) * nb // 2]).reshape(-1))


def test_tied_output_falls_back_to_embedding(tmp_path, cfg):
    tensors = llama_tensors(cfg)
    names = [n for n in tensors if n != "output.weight"]
    p = str(tmp_path / "tied.gguf")
    write_llama_gguf(p, cfg, tensors, names, extra_meta=False)
    with gguf_file.GgufArchive(p) as ar:
        w = M.LlamaWeights.from_gguf(ar, "cpu")
        assert torch.equal(w.output[0], w.tok_embd[0]) and w.output[1] == w.tok_embd[1]
