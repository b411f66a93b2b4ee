"""Dev (GPU box): W4A16 swap-AB kernel at Mistral-7B shapes, decode batch sizes; weights rotated
through > L2 so every launch streams from HBM.  Also the config-4 decode step (per-step graph)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.load_package()
from mistralrs_b200 import gptq, gptq_model as G, lib

dev = torch.device("cuda:0")
peak = 6582.5
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def bench(fn, iters=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(iters): fn(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run_w4(N, K, M, group=128):
    wbytes = N * K // 2 + (K // group) * N * 2
    copies = max(2, int(400e6 // wbytes) + 1)
    layers = []
    for c in range(copies):
        qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev)
        sc = torch.rand(K // group, N, device=dev).to(torch.float16) * 0.01
        tiles = torch.empty(K // 16, N * 2, dtype=torch.int32, device=dev)
        lib().gptq_marlin_repack(ctypes.c_void_p(qw.data_ptr()), ctypes.c_void_p(0), ctypes.c_void_p(tiles.data_ptr()), K, N, 4,
                                 ctypes.c_int64(torch.cuda.current_stream().cuda_stream))
        layers.append((tiles, sc))
    x = torch.randn(M, K, device=dev).to(torch.float16)
    y = torch.empty(M, N, dtype=torch.float16, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    def fn(i):
        t, s = layers[i % copies]
        rc = lib().mrs_w4a16_gemm(P(x), P(t), P(s), ctypes.c_void_p(0), P(y), M, K, N, group, 0, 0,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    us = bench(fn)
    print(f"w4a16 N={N:6d} K={K:6d} M={M:3d}: {us:8.1f} us  {wbytes/us/1e3:7.0f} GB/s ({wbytes/us/1e3/peak*100:5.1f}% of HBM)  "
          f"{2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)


def run_dense(N, K, M):
    wbytes = N * K * 2
    copies = max(2, int(400e6 // wbytes) + 1)
    ws = [torch.randn(N, K, device=dev).to(torch.float16) * 0.02 for _ in range(copies)]
    x = torch.randn(M, K, device=dev).to(torch.float16)
    def fn(i):
        gptq.dense_linear(x, ws[i % copies])
    us = bench(fn)
    print(f"dense N={N:6d} K={K:6d} M={M:3d}: {us:8.1f} us  {wbytes/us/1e3:7.0f} GB/s ({wbytes/us/1e3/peak*100:5.1f}% of HBM)", flush=True)


if __name__ == "__main__":
    for M in (32, 1, 8, 64):
        for (N, K) in [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]:
            run_w4(N, K, M)
    run_dense(32000, 4096, 32)
    # whole decode step, config 4
    cfg = G.GptqConfig.mistral_7b()
    if len(sys.argv) > 1:
        cfg.n_layers = int(sys.argv[1])
    w = G.GptqWeights(cfg, dev)
    for layout in ("hnd", "vllm"):
        for B in (32,):
            run = G.GptqRunner(w, batch=B, max_ctx=400, cache_layout=layout)
            for mask, name in ((0, "full"), (1, "linears-only"), (2, "attention-only")):
                run.step_struct.skip_mask = mask
                run.reset(256); run.step(); torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    run.step()
                for _ in range(3): gr.replay()
                run.reset(256)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): gr.replay()
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 20 * 1e3
                print(f"config4 {layout} batch={B} layers={cfg.n_layers} {name}: {us:9.1f} us/step -> {B*1e6/us:9.0f} tok/s", flush=True)
            run.step_struct.skip_mask = 0
