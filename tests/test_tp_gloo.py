"""Tensor-parallel host logic on CPU (gloo, world_size 2): the column / row sharding rules of
the reference (distributed/layers.rs:1167-1294 column = output rows; :695-975 row = K columns on
quant-block boundaries, outputs sum-all-reduced) applied to ggml blocks, checked with the CPU
oracle: gathered column shards and all-reduced row shards must reproduce the unsharded product."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    import oracle
    from mistralrs_b200 import model as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=2, hidden=512, inter=1024)
    full = M.LlamaWeights(cfg, torch.device("cpu"), keep_host=True)
    shard = M.LlamaWeights(cfg, torch.device("cpu"), tp_rank=rank, tp_size=world, keep_host=True)
    rng = np.random.default_rng(5)
    ok = True
    for layer in range(cfg.n_layers):
        # column parallel (ffn_gate): rows [rank*N/w, (rank+1)*N/w) -> all-gather == full
        ty = M.tensor_type(cfg, "ffn_gate", layer)
        x = rng.standard_normal((1, cfg.hidden)).astype(np.float32)
        xq, st = oracle.quantize_q8_1(x)
        y_full = oracle.mmvq_q8_1(ty, full.host[(layer, "ffn_gate")], xq, cfg.hidden, cfg.inter, st, 1)
        y_part = oracle.mmvq_q8_1(ty, shard.host[(layer, "ffn_gate")], xq, cfg.hidden, cfg.inter // world, st, 1)
        parts = [torch.zeros(1, cfg.inter // world, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(y_part))
        ok &= bool(np.array_equal(torch.cat(parts, dim=1).numpy(), y_full))
        # row parallel (ffn_down): K slice on block boundaries, partial sums all-reduced
        ty = M.tensor_type(cfg, "ffn_down", layer)
        a = rng.standard_normal((1, cfg.inter)).astype(np.float32)
        aq, st = oracle.quantize_q8_1(a)
        y_full = oracle.mmvq_q8_1(ty, full.host[(layer, "ffn_down")], aq, cfg.inter, cfg.hidden, st, 1)
        kl = cfg.inter // world
        a_loc = np.ascontiguousarray(a[:, rank * kl:(rank + 1) * kl])
        aq_loc, st_loc = oracle.quantize_q8_1(a_loc)
        y_part = torch.from_numpy(oracle.mmvq_q8_1(ty, shard.host[(layer, "ffn_down")], aq_loc, kl, cfg.hidden, st_loc, 1))
        dist.all_reduce(y_part)
        ok &= bool(np.allclose(y_part.numpy(), y_full, rtol=1e-12, atol=1e-9))
    # replicated tensors are identical on every rank
    ok &= bool(np.array_equal(full.host[(0, "output")], shard.host[(0, "output")]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_column_and_row_parallel_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_kv_head_and_block_alignment_rules():
    sys.path.insert(0, ROOT)
    from mistralrs_b200 import model as M
    cfg = M.LlamaConfig.llama3_8b()
    for w in (2, 4, 8):   # K shards stay multiples of the 256 super-block
        assert (cfg.inter // w) % 256 == 0 and (cfg.n_heads * cfg.head_dim // w) % 256 == 0
        assert cfg.n_kv_heads % w == 0
    c70 = M.LlamaConfig.llama3_70b()
    assert (c70.inter // 8) % 256 == 0 and c70.n_kv_heads // 8 == 1


def test_kv_head_shard_rule():
    """`compute_kv_shard` / `compute_n_kv_groups` (REF mistralrs-quant/src/distributed/layers.rs:2692-2733): KV heads split
    over the ranks, replicated on consecutive ranks when the ranks outnumber them."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from mistralrs_b200 import model as M
    D = 128
    assert M.compute_kv_shard(8, D, 0, 1) == (0, 8 * D)
    assert [M.compute_kv_shard(8, D, r, 2) for r in range(2)] == [(0, 4 * D), (4 * D, 4 * D)]
    assert [M.compute_kv_shard(8, D, r, 8) for r in range(8)] == [(r * D, D) for r in range(8)]
    # 2 KV heads on 8 ranks: each head on 4 consecutive ranks
    assert [M.compute_kv_shard(2, D, r, 8)[0] // D for r in range(8)] == [0, 0, 0, 0, 1, 1, 1, 1]
    assert all(M.compute_kv_shard(2, D, r, 8)[1] == D for r in range(8))
    assert M.compute_n_kv_groups(2, 32, 8) == 4 and M.compute_n_kv_groups(8, 32, 8) == 4 and M.compute_n_kv_groups(8, 32, 1) == 4
    with pytest.raises(ValueError):
        M.compute_kv_shard(8, D, 0, 3)          # does not divide
    with pytest.raises(ValueError):
        M.compute_kv_shard(3, D, 0, 8)          # cannot replicate evenly


def test_kv_heads_replicated_when_ranks_outnumber_them():
    """one KV head, two ranks: both ranks hold the whole K / V projection, the query heads are still split"""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from mistralrs_b200 import model as M
    cfg = M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=1, hidden=512, inter=1024, n_heads=8, n_kv_heads=1)
    full = M.LlamaWeights(cfg, torch.device("cpu"), keep_host=True)
    shards = [M.LlamaWeights(cfg, torch.device("cpu"), tp_rank=r, tp_size=2, keep_host=True) for r in range(2)]
    for name in ("attn_k", "attn_v"):
        for sh in shards:
            assert np.array_equal(sh.host[(0, name)], full.host[(0, name)])
            assert sh.layers[0][name][2] == cfg.head_dim
    q_rows = [sh.layers[0]["attn_q"][2] for sh in shards]
    assert q_rows == [cfg.n_heads * cfg.head_dim // 2] * 2
    assert np.array_equal(np.concatenate([sh.host[(0, "attn_q")] for sh in shards]), full.host[(0, "attn_q")])


# ---- the layer classes (distributed.py) with biases, under gloo: the CUDA linear replaced by an oracle stand-in ----
def _layer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    import oracle
    from mistralrs_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class CpuLinear:                       # stand-in for quant.GgufMatMul: exact product with the dequantised shard
        def __init__(self, data, dtype, shape):
            self.w = torch.from_numpy(oracle.dequantize(dtype, data.numpy()).reshape(shape).astype(np.float64))

        def forward(self, xs):
            return xs @ self.w.T

    ok = True
    for dtype, n, k in (("q4_k", 96, 512), ("q6_k", 64, 1024), ("q8_0", 48, 192)):
        rng = np.random.default_rng(11)                                   # same bytes on every rank
        blocks = torch.from_numpy(oracle.random_blocks(dtype, n * k // oracle.BLOCK_ELEMS[dtype], rng).reshape(-1))
        bias = torch.from_numpy(rng.standard_normal(n))
        x = torch.from_numpy(rng.standard_normal((3, k)))
        full = D.ReplicatedLayer(blocks, dtype, (n, k), bias, CpuLinear).forward(x)
        col = D.ColumnParallelLayer(blocks, dtype, (n, k), rank, world, bias, CpuLinear)
        parts = [torch.zeros(3, n // world, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, col.forward(x))
        ok &= bool(torch.equal(torch.cat(parts, dim=1), full))            # rows are disjoint: bit-identical
        row = D.RowParallelLayer(blocks, dtype, (n, k), rank, world, D.SumAllReduce(), bias, CpuLinear)
        y = row.forward(row.input_slice(x).contiguous())
        ok &= bool(torch.allclose(y, full, rtol=1e-12, atol=1e-12))       # bias once, after the reduce
        ok &= row.shape == (n, k // world) and col.shape == (n // world, k)
    single = D.SumAllReduce(world_size=1)
    t = torch.ones(2)
    ok &= single.is_noop() and single.sum_all_reduce(t) is t and not D.SumAllReduce().is_noop()
    custom = D.SumAllReduce(world_size=world, reduce_fn=lambda v: v * world)   # a plugged-in exchange is what gets called
    ok &= bool(torch.equal(custom.sum_all_reduce(torch.ones(2)), torch.full((2,), float(world))))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_parallel_layers_with_bias_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + 3) % 2000
    procs = [ctx.Process(target=_layer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_rules_reject_what_the_reference_rejects():
    import oracle
    from mistralrs_b200 import distributed as D
    blocks = torch.from_numpy(oracle.random_blocks("q4_k", 6 * 3, np.random.default_rng(0)).reshape(-1))   # [6, 768]
    with pytest.raises(ValueError, match="block boundaries"):
        D.shard_k_blocks(blocks, "q4_k", (6, 768), 0, 2)          # 3 blocks of 256 do not split in two
    with pytest.raises(ValueError, match="do not divide"):
        D.shard_rows(blocks, "q4_k", (6, 768), 0, 4)
    s, shape = D.shard_k_blocks(blocks, "q4_k", (6, 768), 1, 3)
    assert shape == (6, 256) and s.numel() == 6 * 144
    assert torch.equal(s.reshape(6, 144), blocks.reshape(6, 3, 144)[:, 1])
