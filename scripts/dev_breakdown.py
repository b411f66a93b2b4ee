"""Dev: per-token time breakdown of the decode graph (full / GEMV-only / attention-only), optionally
sweeping the split-KV chunk size.  usage: dev_breakdown.py [layers] [min_tokens ...]
MRS_DEV_LIB=path runs against another build of libmrs_b200.so (A/B)."""
import sys, os
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
pkg = g.load_package()
if os.environ.get("MRS_DEV_LIB"):
    pkg.LIB_PATH = os.environ["MRS_DEV_LIB"]
from mistralrs_b200 import model as M
import ctypes
if os.environ.get("MRS_CTAS"):
    pkg.lib().mrs_set_mmvq_ctas_per_sm(ctypes.c_int(int(os.environ["MRS_CTAS"])))
if os.environ.get("MRS_FLAGS"):
    pkg.lib().mrs_set_mmvq_flags(ctypes.c_int(int(os.environ["MRS_FLAGS"], 0)))
dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sweep = [int(a) for a in sys.argv[2:]] or [64]
cfg = M.LlamaConfig.llama3_8b(); cfg.n_layers = layers
w = M.LlamaWeights(cfg, dev)


def timed(run, mask, reps=30):
    """mask: 0 full step, 1 GEMVs only (attention skipped), 2 attention only"""
    run.step_struct.skip_mask = mask
    run.reset(); run.context_lens.fill_(256)
    run.step(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        run.step()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    run.context_lens.fill_(256)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gr.replay()
    e1.record(); torch.cuda.synchronize()
    run.step_struct.skip_mask = 0
    return e0.elapsed_time(e1) / reps * 1e3


for pdl in (1, 0):
    for mt in sweep:
        run = M.LlamaRunner(w, batch=1, max_ctx=400, pdl=bool(pdl), fused_attention=True, split_min_tokens=mt)
        full, gemv, attn = timed(run, 0), timed(run, 1), timed(run, 2)
        print(f"layers={layers} pdl={pdl} split_min_tokens={mt} tiles={run.padded_tiles}: full {full:8.1f} us  "
              f"gemv-only {gemv:8.1f} us  attn-only {attn:8.1f} us  -> {1e6/full:6.1f} tok/s", flush=True)
