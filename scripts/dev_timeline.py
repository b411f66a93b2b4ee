"""Dev: in-kernel timeline (globaltimer stamps) of the first decode GEMVs of a token, from a
-DMRS_TIMELINE build of mmvq.cu (MRS_DEV_LIB=path).  Prints per-launch phase medians in us."""
import sys, os, ctypes
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
pkg = g.load_package()
pkg.LIB_PATH = os.environ["MRS_DEV_LIB"]
from mistralrs_b200 import model as M, lib
dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pdl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
NL = 16
cfg = M.LlamaConfig.llama3_8b(); cfg.n_layers = layers
w = M.LlamaWeights(cfg, dev)
run = M.LlamaRunner(w, batch=1, max_ctx=400, pdl=bool(pdl), fused_attention=True)
run.reset(); run.context_lens.fill_(256)
run.step(); torch.cuda.synchronize()
buf = torch.zeros(NL * 320 * 16, dtype=torch.int64, device=dev)
L = lib()
L.mrs_mmvq_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.mrs_mmvq_timeline(ctypes.c_void_p(buf.data_ptr()), NL)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    run.step()
for _ in range(5): gr.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(NL, 320, 16)
prev_end = None
names = ["entry", "init", "issue1", "waited", "pass0", "pass1", "pass2", "full0", "loopend", "exit", "prodend"]
print("pdl", pdl, "| columns: median over CTAs, us since the launch's first CTA entry; (max) where shown")
for i in range(NL):
    grid = int(t[i, 0, 15])
    if grid == 0: continue
    a = t[i, :grid].astype(np.float64)
    K = int(t[i, 0, 14]) & 0xffffffff; vrows = int(t[i, 0, 14]) >> 32
    t0 = a[:, 0].min()
    rel = (a[:, :11] - t0) / 1e3
    med = np.median(rel, axis=0)
    mx = rel.max(axis=0)
    gap = (t0 - prev_end) / 1e3 if prev_end is not None else float("nan")
    print(f"#{i:2d} K={K:5d} vrows={vrows:6d} grid={grid:3d} gap_prev_exit->entry {gap:6.2f} | " +
          " ".join(f"{n}={med[j]:5.2f}" for j, n in enumerate(names)) +
          f" | entry_max={mx[0]:5.2f} waited_max={mx[3]:5.2f} exit_max={mx[9]:5.2f}")
    prev_end = a[:, 9].max()
