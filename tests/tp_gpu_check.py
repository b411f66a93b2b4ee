"""Run under torchrun (N ranks, NCCL): tensor-parallel decode of a tiny model must match the
single-GPU decode of the same synthetic weights (the all-reduce sums bf16 partials, so a few
bf16 ulps of difference are expected — same as the reference's RowParallelLayer)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.load_package()
from mistralrs_b200 import model as M  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=4, hidden=1024, inter=2048, n_heads=16, n_kv_heads=8)
    full = M.LlamaWeights(cfg, dev)
    shard = M.LlamaWeights(cfg, dev, tp_rank=rank, tp_size=world)
    bufs = {}

    def comm(buf, count, dtype, stream, user):
        dist.all_reduce(bufs[buf])

    ref = M.LlamaRunner(full, batch=2, max_ctx=64)
    tp = M.LlamaRunner(shard, batch=2, max_ctx=64, comm=comm)
    for n in ("x", "x2"):
        bufs[tp.buf[n].data_ptr()] = tp.buf[n]
    tp.capture()   # exercises NCCL inside CUDA-graph capture
    # the product path: in-graph peer-memory all-reduce fused with the residual add, PDL chain intact
    peer = M.PeerAllReduce(2 * cfg.hidden, torch.bfloat16, dev, low_latency=True)
    tp2 = M.LlamaRunner(shard, batch=2, max_ctx=64, pdl=True, peer_allreduce=peer)
    tp2.capture()
    assert peer.low_latency
    # the flags + pull protocol (the low-latency region switched off): same arithmetic, must give the same bits
    peer3 = M.PeerAllReduce(2 * cfg.hidden, torch.bfloat16, dev, low_latency=False)
    tp3 = M.LlamaRunner(shard, batch=2, max_ctx=64, pdl=True, peer_allreduce=peer3)
    tp3.capture()
    results, last = {}, {}
    for name, runner in (("nccl", tp), ("peer", tp2), ("peer-pull", tp3)):
        ref.reset(); runner.reset()
        ref.set_tokens([11, 400]); runner.set_tokens([11, 400])
        worst = 0.0
        for step in range(8):
            ref.step(); runner.graph.replay()
            torch.cuda.synchronize()
            a, b = ref.logits().float(), runner.logits().float()
            worst = max(worst, ((a - b).abs().max() / a.abs().max()).item())
            runner.set_tokens(ref.meta["token_ids"].cpu().tolist())
        results[name] = worst
        last[name] = runner.logits().float().clone()
    # every rank must hold identical logits on the peer path (rank-ordered f32 sums)
    mine = tp2.logits().float().clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(mine, other)) and bool(torch.equal(last["peer"], last["peer-pull"]))
    ok = max(results.values()) <= 4 * 2.0 ** -7 and same
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"TP{world} vs TP1 worst rel logit diff: nccl {results['nccl']:.3e}, peer-memory low-latency {results['peer']:.3e} / flags+pull {results['peer-pull']:.3e}, "
              f"ranks and both protocols bit-identical: {same} -> {'OK' if t.item() == 1.0 else 'FAIL'}", flush=True)
    rc = 0 if t.item() == 1.0 else 1
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(rc)   # skip NCCL teardown (hangs while captured collectives are alive)


if __name__ == "__main__":
    main()
