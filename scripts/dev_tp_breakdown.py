"""Dev (GPU box, torchrun N ranks): per-token graph time of the tensor-parallel decode stack, split by the
skip masks (full / no attention / no GEMVs) and with the all-reduces removed (sharded weights, no TP context)."""
import os, sys
import torch, torch.distributed as dist
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
g.load_package()
from mistralrs_b200 import model as M
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
cfg = M.LlamaConfig.llama3_8b()
w = M.LlamaWeights(cfg, dev, tp_rank=rank, tp_size=world, fast_synth=True)
peer = M.PeerAllReduce(cfg.hidden, w.dtype, dev)


def timed(run, mask, reps=30):
    run.step_struct.skip_mask = mask
    run.reset(); run.context_lens.fill_(256)
    run.step(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        run.step()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize(); dist.barrier()
    run.context_lens.fill_(256)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gr.replay()
    e1.record(); torch.cuda.synchronize()
    run.step_struct.skip_mask = 0
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for name, pa in (("peer all-reduce", peer), ("no all-reduce (sharded GEMVs only; wrong sums, timing only)", None)):
    run = M.LlamaRunner(w, batch=1, max_ctx=400, pdl=True, peer_allreduce=pa)
    if pa is None:
        run._ar_cb = None
        run.step_struct.tp = None
    full, gemv, attn = timed(run, 0), timed(run, 1), timed(run, 2)
    if rank == 0:
        print(f"tp{world} {name}: full {full:8.1f} us  no-attention {gemv:8.1f} us  no-GEMV {attn:8.1f} us  -> {1e6/full:6.1f} tok/s", flush=True)
dist.destroy_process_group()
