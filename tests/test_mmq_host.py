"""fast_mmq's entry points (`plain`, `fused_qkv`, `fused_glu`, `fused_ffn`; REF gguf/fast_mmq.rs:760-826) as compositions of
the prefill GEMM: host plumbing only, with the GEMM and the GLU kernel replaced by CPU stand-ins (the kernels themselves
are covered by the GPU suite)."""
import types

import pytest
import torch

from mistralrs_b200 import mmq, ops


def _w(n, k, seed):
    g = torch.Generator().manual_seed(seed)
    return types.SimpleNamespace(shape=(n, k), dense=torch.randn(n, k, generator=g))


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(mmq, "forward", lambda w, xs: xs @ w.dense.T)
    monkeypatch.setattr(ops, "fused_glu", lambda a, b, act: torch.nn.functional.silu(a) * b)


def test_compositions(cpu_kernels):
    x = torch.randn(2, 9, 32)
    q, k, v, g, u, d = _w(48, 32, 1), _w(16, 32, 2), _w(16, 32, 3), _w(64, 32, 4), _w(64, 32, 5), _w(32, 64, 6)
    assert torch.equal(mmq.plain(q, x), x @ q.dense.T)
    yq, yk, yv = mmq.fused_qkv(q, k, v, x)
    assert yq.shape == (2, 9, 48) and torch.equal(yk, x @ k.dense.T) and torch.equal(yv, x @ v.dense.T)
    glu = torch.nn.functional.silu(x @ g.dense.T) * (x @ u.dense.T)
    assert torch.equal(mmq.fused_glu(g, u, x, 0), glu)
    assert torch.equal(mmq.fused_ffn(g, u, d, x, 0), glu @ d.dense.T)


def test_shape_mismatch_messages(cpu_kernels):
    x = torch.randn(4, 32)
    with pytest.raises(ValueError, match="fused_glu: gate/up shape mismatch"):
        mmq.fused_glu(_w(64, 32, 1), _w(48, 32, 2), x, 0)
    with pytest.raises(ValueError, match="fused_ffn: gate/up shape mismatch"):
        mmq.fused_ffn(_w(64, 32, 1), _w(48, 32, 2), _w(32, 64, 3), x, 0)
