"""Dev script (GPU box): time our decode GEMV against the reference's own kernel
(oracle/_ref/libref_mmvq.so, unmodified reference source) on Llama-3-8B shapes.
Inputs larger than L2 are cycled so every launch streams from HBM."""
import ctypes, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.load_package()
import oracle
from mistralrs_b200 import quant, lib

if os.environ.get("MRS_CTAS"):
    lib().mrs_set_mmvq_ctas_per_sm(ctypes.c_int(int(os.environ["MRS_CTAS"])))
dev = torch.device("cuda:0")
ref = oracle.ref_lib("mmvq")
peak = 6582.5
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

def cur():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def bench(fn, iters=20):
    """Time `iters` back-to-back launches replayed from one CUDA graph (no CPU launch cost)."""
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def run(dtype, N, K, batch=1, mode="plain"):
    bb, be = oracle.BLOCK_BYTES[dtype], oracle.BLOCK_ELEMS[dtype]
    wbytes = N * K // be * bb
    nmats = 2 if mode == "glu" else 1
    copies = max(2, int(400e6 // (wbytes * nmats)) + 1)   # > 126 MB L2 in rotation
    rng = np.random.default_rng(1)
    base = torch.from_numpy(oracle.random_blocks(dtype, N * K // be, rng).reshape(-1)).to(dev)
    ws = [quant.QTensor(base.clone(), dtype, (N, K)) for _ in range(copies * nmats)]
    x = torch.randn(batch, K, device=dev).to(torch.bfloat16)
    kp = (K + 511) // 512 * 512
    scratch = torch.empty(batch * kp // 32 * 36, dtype=torch.uint8, device=dev)
    out = torch.empty(batch, N, dtype=torch.bfloat16, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    res = {}
    for name, L in (("ours", lib()), ("ref", ref)):
        if L is None: continue
        qf = getattr(L, "launch_mmvq_gguf_quantize_q8_1_bf16")
        qf(P(x), P(scratch), K, kp, batch, cur())
        if mode == "plain":
            f = getattr(L, f"launch_mmvq_gguf_{dtype}_bf16_plain")
            fn = lambda i: f(P(ws[i % copies].data), P(scratch), P(out), K, N, kp // 32, N, batch, cur())
        else:
            f = getattr(L, f"launch_mmvq_gguf_{dtype}_bf16_fused_glu")
            fn = lambda i: f(P(ws[2 * (i % copies)].data), P(ws[2 * (i % copies) + 1].data), P(scratch), P(out), K, N, kp // 32, N, batch, 0, cur())
        us = bench(fn)
        res[name] = (us, wbytes * nmats / us / 1e3, out.float().clone())
    line = f"{dtype:5s} {mode:5s} N={N:6d} K={K:6d} b={batch} bytes={wbytes*nmats/1e6:7.1f}MB"
    for name, (us, gbs, _) in res.items():
        line += f" | {name}: {us:8.1f} us {gbs:7.0f} GB/s ({gbs/peak*100:4.1f}%)"
    if "ref" in res:
        d = (res["ours"][2] - res["ref"][2]).abs().max().item()
        line += f" | maxdiff vs ref {d:.3g}"
    print(line, flush=True)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "one":
    run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    sys.exit(0)
if __name__ == "__main__":
    shapes = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336), (128256, 4096)]
    for dtype in ("q4_k", "q6_k", "q8_0"):
        for (N, K) in shapes:
            run(dtype, N, K)
    run("q4_k", 14336, 4096, mode="glu")
    run("q4_k", 4096, 4096, batch=4)
    run("q4_k", 4096, 4096, batch=8)
    for dtype in ("q2_k", "q3_k", "q5_k", "q4_0", "q4_1", "q5_0", "q5_1"):
        run(dtype, 4096, 4096)
