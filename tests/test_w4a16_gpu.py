"""Small-batch W4A16 (csrc/w4a16.cu) behind the reference's Marlin symbols: `gptq_marlin_repack` /
`awq_marlin_repack` + `marlin_{gptq,awq}_4bit_{f16,bf16}` driven exactly as `gptq_linear` /
`marlin_matmul` drive them (repack, marlin_permute_scales, forward), against the numpy oracle
(oracle/gptq.py) — and, in test_ref_golden_gpu.py, against the reference's own Marlin kernel."""
import numpy as np
import pytest
import torch

from oracle import gptq as og
from mistralrs_b200 import gptq

pytestmark = pytest.mark.gpu


def _mk(K, N, group, seed):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N))
    g = K if group == -1 else group
    scales = np.exp2(rng.uniform(-8, -6, size=(K // g, N))).astype(np.float16)
    return rng, q, scales, g


def _tol(ref, x, wd, ulp):
    mag = np.abs(x.astype(np.float64)) @ np.abs(wd).astype(np.float64)
    return ulp * np.abs(ref) * 1.01 + 2e-6 * mag + 1e-6


@pytest.mark.parametrize("M", [1, 7, 32, 33, 64, 100, 300])
def test_gptq_marlin_boundary_batches(cuda, M):
    K, N, group = 1024, 576, 128          # 4.5 row tiles (ragged last tile), split-K over a cluster
    rng, q, scales, g = _mk(K, N, group, 1)
    qweight = og.pack_gptq(q)
    x = rng.standard_normal((M, K)).astype(np.float16)
    layer = gptq.GptqMarlinLayer(torch.from_numpy(qweight).to(cuda), torch.from_numpy(scales).to(cuda), group_size=group)
    y = layer.forward_raw(torch.from_numpy(x).to(cuda)).float().cpu().numpy()
    wd = og.dequant_gptq(qweight, scales, None, g)
    ref = og.gemm(x, wd)
    tol = _tol(ref, x, wd, 2.0 ** -11)
    assert (np.abs(y - ref) <= tol).all(), float((np.abs(y - ref) / tol).max())


@pytest.mark.parametrize("K,N,group", [(4096, 4096, 128), (512, 256, 32), (1024, 192, 64), (2048, 128, -1), (14336, 256, 128)])
def test_gptq_marlin_shapes_and_groups(cuda, K, N, group):
    M = 32
    rng, q, scales, g = _mk(K, N, group, 3)
    qweight = og.pack_gptq(q)
    x = (0.5 * rng.standard_normal((M, K))).astype(np.float16)
    layer = gptq.GptqMarlinLayer(torch.from_numpy(qweight).to(cuda), torch.from_numpy(scales).to(cuda), group_size=group)
    y = layer.forward_raw(torch.from_numpy(x).to(cuda)).float().cpu().numpy()
    wd = og.dequant_gptq(qweight, scales, None, g)
    ref = og.gemm(x, wd)
    tol = _tol(ref, x, wd, 2.0 ** -11)
    assert (np.abs(y - ref) <= tol).all(), float((np.abs(y - ref) / tol).max())


def test_gptq_act_order_follows_the_reference_flow(cuda):
    # desc_act checkpoints: the reference sorts the weight rows by group (perm = argsort(g_idx)) at
    # repack time and runs Marlin with contiguous groups on the UNPERMUTED activations
    # (gptq_cuda.rs:573-590, marlin_backend.rs): y = x . W[perm] with scales[k' / group]
    K, N, group, M = 1024, 256, 128, 16
    rng, q, scales, g = _mk(K, N, group, 5)
    g_idx = (rng.permutation(K) // group).astype(np.int32)
    qweight = og.pack_gptq(q)
    x = rng.standard_normal((M, K)).astype(np.float16)
    layer = gptq.GptqMarlinLayer(torch.from_numpy(qweight).to(cuda), torch.from_numpy(scales).to(cuda),
                                 g_idx=torch.from_numpy(g_idx).to(cuda), group_size=group)
    y = layer.forward_raw(torch.from_numpy(x).to(cuda)).float().cpu().numpy()
    perm = np.argsort(g_idx, kind="stable")
    wd = ((q[perm] - 8).astype(np.float32) * scales.astype(np.float32)[np.arange(K) // group]).astype(np.float16).astype(np.float32)
    ref = og.gemm(x, wd)
    tol = _tol(ref, x, wd, 2.0 ** -11)
    assert (np.abs(y - ref) <= tol).all(), float((np.abs(y - ref) / tol).max())


@pytest.mark.parametrize("K,N,group,M", [(512, 384, 128, 40), (448, 160, 64, 9), (1152, 320, 32, 32), (4096, 1024, 128, 32)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_awq_marlin_boundary(cuda, dt, K, N, group, M):
    """zero points + scale rows staged per iteration: several groups per 128-k iteration (32, 64), a K tail of
    one 64-k chunk (448, 1152 % 128 == 64... 1152 = 9 x 128), ragged last row tile (160, 320; N stays a multiple of the scale permutation width the layer applies), split-K (4096)"""
    rng, q, scales, g = _mk(K, N, group, 2)
    z = rng.integers(0, 16, size=(K // group, N))
    qweight, qzeros = og.pack_awq(q), og.pack_awq(z)
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(cuda).to(dt)
    layer = gptq.GptqMarlinLayer(torch.from_numpy(qweight).to(cuda), torch.from_numpy(scales).to(cuda),
                                 qzeros=torch.from_numpy(qzeros).to(cuda), group_size=group, is_awq=True)
    y = layer.forward_raw(x).float().cpu().numpy()
    xs = x.float().cpu().numpy()
    sc = scales if dt == torch.float16 else torch.from_numpy(scales).to(torch.bfloat16).float().numpy()
    gidx = np.arange(K) // group
    wd = (q - z[gidx]).astype(np.float32) * sc.astype(np.float32)[gidx]
    wd = torch.from_numpy(wd).to(dt).float().numpy()           # one rounding of the dequantised weight in dt
    ref = xs.astype(np.float64) @ wd.astype(np.float64)
    ulp = 2.0 ** -11 if dt == torch.float16 else 2.0 ** -8
    tol = _tol(ref, xs, wd, ulp)
    assert (np.abs(y - ref) <= tol).all(), float((np.abs(y - ref) / tol).max())


@pytest.mark.parametrize("M,N,K", [(32, 32000, 4096), (5, 200, 512), (130, 384, 1024)])
def test_dense_linear(cuda, M, N, K):
    rng = np.random.default_rng(7)
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(cuda).to(torch.float16)
    w = torch.from_numpy((0.05 * rng.standard_normal((N, K))).astype(np.float32)).to(cuda).to(torch.float16)
    y = gptq.dense_linear(x, w).float()
    ref = x.double() @ w.double().t()
    mag = x.double().abs() @ w.double().abs().t()
    tol = 2.0 ** -11 * ref.abs() * 1.01 + 2e-6 * mag + 1e-6
    assert bool(((y.double() - ref).abs() <= tol).all()), float(((y.double() - ref).abs() / tol).max())
