"""Dev (GPU box): tcgen05 dequant-GEMM throughput vs cuBLAS bf16 on pre-dequantised weights."""
import sys, os, json
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
g.load_package()
import oracle
from mistralrs_b200 import mmq, quant
dev = torch.device("cuda:0")
def bench(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for fmt in ("auto", "tc"):
    mmq.set_path(fmt)
    for dtype in ("q8_0", "q4_k", "q6_k"):
        for (M, N, K) in [(4096, 4096, 4096), (4096, 14336, 4096), (4096, 4096, 14336)]:
            rng = np.random.default_rng(0)
            wb = oracle.random_blocks(dtype, N * K // oracle.BLOCK_ELEMS[dtype], rng)
            w = quant.QTensor(torch.from_numpy(wb.reshape(-1)).to(dev), dtype, (N, K))
            x = torch.randn(M, K, device=dev).to(torch.bfloat16)
            ms = bench(lambda: mmq.forward(w, x))
            wd = torch.from_numpy(oracle.dequantize(dtype, wb).reshape(N, K)).to(dev).to(torch.bfloat16)
            ms_ref = bench(lambda: x @ wd.t()) if fmt == "auto" else float("nan")
            y = mmq.forward(w, x).float(); yr = (x.float() @ wd.float().t())
            err = ((y - yr).abs().max() / yr.abs().max()).item()
            fl = 2.0 * M * N * K
            print(f"{fmt} {dtype} M={M} N={N} K={K}: ours {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s | cuBLAS bf16 (dense W) {ms_ref:7.3f} ms {fl/ms_ref/1e9:7.1f} TF/s | rel diff {err:.2e}", flush=True)
