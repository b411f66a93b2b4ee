import os, sys, time
import torch, torch.distributed as dist
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(1024, device=dev, dtype=torch.bfloat16)
dist.all_reduce(t); torch.cuda.synchronize(); log("eager all_reduce ok", t[0].item())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(t)
g.replay(); torch.cuda.synchronize(); log("graph all_reduce ok", t[0].item())
import __graft_entry__ as ge
ge.load_package()
from mistralrs_b200 import model as M
cfg = M.LlamaConfig.tiny_test(quant="q4_k_m", n_layers=2, hidden=1024, inter=2048, n_heads=16, n_kv_heads=8)
shard = M.LlamaWeights(cfg, dev, tp_rank=rank, tp_size=world)
bufs = {}
calls = [0]
def comm(buf, count, dtype, stream, user):
    calls[0] += 1
    dist.all_reduce(bufs[buf])
tp = M.LlamaRunner(shard, batch=1, max_ctx=64, comm=comm)
for n in ("x", "x2"): bufs[tp.buf[n].data_ptr()] = tp.buf[n]
log("runner built")
tp.set_tokens([5]); tp.step(); torch.cuda.synchronize(); log("eager TP step ok, callbacks", calls[0], "tok", tp.meta["token_ids"].cpu().tolist())
tp.capture(); log("capture ok")
tp.graph.replay(); torch.cuda.synchronize(); log("replay ok", tp.meta["token_ids"].cpu().tolist())
dist.destroy_process_group(); log("done")
