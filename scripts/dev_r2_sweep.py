"""Dev (round 2): one process, weights built once; per-token graph time (full / GEMV-only / attention-only)
under the decode-GEMV knobs: CTAs per SM (2 | 3) and the long-segment variant (on | off)."""
import ctypes, os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
pkg = g.load_package()
if os.environ.get("MRS_DEV_LIB"):
    pkg.LIB_PATH = os.environ["MRS_DEV_LIB"]     # A/B another build of libmrs_b200.so
from mistralrs_b200 import model as M
dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = M.LlamaConfig.llama3_8b(); cfg.n_layers = layers
w = M.LlamaWeights(cfg, dev)
L = pkg.lib()


def timed(run, mask, reps=30):
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


for name, mflags, aflags in (("default (mma attention, cluster merge)", 128 << 8, 0), ("mma attention, counter merge", 128 << 8, 2),
                             ("simt attention", 128 << 8, 1)):
    L.mrs_set_mmvq_flags(ctypes.c_int(mflags))
    L.mrs_set_attn_flags(ctypes.c_int(aflags))
    run = M.LlamaRunner(w, batch=1, max_ctx=400, pdl=True)
    full, gemv, attn = timed(run, 0), timed(run, 1), timed(run, 2)
    print(f"layers={layers} {name}: full {full:8.1f} us  gemv-only {gemv:8.1f} us  attn-only {attn:8.1f} us  -> {1e6/full:6.1f} tok/s", flush=True)
L.mrs_set_mmvq_flags(ctypes.c_int(128 << 8)); L.mrs_set_attn_flags(ctypes.c_int(0))
