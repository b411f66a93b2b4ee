"""Decode-backend selection (host logic): REF mistralrs-core/src/flashinfer/mod.rs:257-273, perf_flags.rs:9-29."""
from mistralrs_b200 import paged_attn as pa


def test_group_sizes_and_head_dims(monkeypatch):
    monkeypatch.delenv(pa.FLASHINFER_DECODE_ENV, raising=False)
    assert pa.flashinfer_decode_enabled()                                           # on by default
    assert pa.flashinfer_supports_layer(32, 8, 128, 128)                            # Llama-3-8B / Mistral-7B: group 4
    assert pa.flashinfer_supports_layer(64, 8, 128, 128)                            # Llama-3-70B: group 8
    assert pa.flashinfer_supports_layer(32, 4, 64, 64)                              # TinyLlama: group 8, head 64
    for g in (1, 2, 3, 4, 6, 8, 16):
        assert pa.supports_flashinfer_group_size(16 * g, 16)
    for q, kv in ((40, 8), (56, 8), (96, 8), (32, 0), (30, 8)):                     # groups 5, 7, 12; no kv heads; not divisible
        assert not pa.supports_flashinfer_group_size(q, kv)
    assert not pa.flashinfer_supports_layer(32, 8, 96, 96) and not pa.flashinfer_supports_layer(32, 8, 192, 128)
    assert pa.flashinfer_supports_layer(8, 8, 512, 512) and pa.flashinfer_supports_layer(8, 4, 256, 256)


def test_env_flag(monkeypatch):
    for v, want in (("0", False), ("false", False), ("off", False), ("no", False), ("FALSE", False), ("1", True), ("on", True),
                    ("yes", True), ("TRUE", True), ("maybe", True), ("", True)):
        monkeypatch.setenv(pa.FLASHINFER_DECODE_ENV, v)
        assert pa.flashinfer_decode_enabled() is want, v
    monkeypatch.setenv(pa.FLASHINFER_DECODE_ENV, "0")
    assert not pa.flashinfer_supports_layer(32, 8, 128, 128)
