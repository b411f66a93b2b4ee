"""Dev (GPU box): in-kernel timeline of CTA (0, 0) of the int4 swap-AB kernel from the -DMRS_WA_TRACE build
(`make -C mistral.rs_b200/csrc trace` -> gpurun_wa_trace.so).  Prints SM-clock deltas per iteration."""
import ctypes, os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
pkg = g.load_package()
pkg.LIB_PATH = os.path.join(R, "gpurun_wa_trace.so")
from mistralrs_b200 import lib
dev = torch.device("cuda:0")
N, K, M, group = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 128
qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev)
sc = torch.rand(K // group, N, device=dev).to(torch.float16) * 0.01
tiles = torch.empty(K // 16, N * 2, dtype=torch.int32, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: torch.cuda.current_stream().cuda_stream
lib().gptq_marlin_repack(P(qw), ctypes.c_void_p(0), P(tiles), K, N, 4, ctypes.c_int64(st()))
x = torch.randn(M, K, device=dev).to(torch.float16)
y = torch.empty(M, N, dtype=torch.float16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    rc = lib().mrs_w4a16_gemm(P(x), P(tiles), P(sc), ctypes.c_void_p(0), P(y), M, K, N, group, 0, 0, ctypes.c_void_p(st()))
    assert rc == 0
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.int64)
rc = lib().mrs_debug_wa_trace(buf.ctypes.data_as(ctypes.c_void_p), 4096)
assert rc == 0
t0 = buf[0]
rel = lambda v: int(v - t0)
print(f"shape N={N} K={K} M={M}: setup done +{rel(buf[1])}  scale table filled +{rel(buf[2])}  epilogue begin +{rel(buf[3])}  end +{rel(buf[4])}  (SM clocks)")
nit = int(np.count_nonzero(buf[1000:1000 + 256:2]))
print("it | producer issue | w2: raw ready, dequant done, A stage free, arrived | w9: same | MMA: operands ready, issued+committed")
for i in range(nit):
    w2 = [rel(buf[200 + 4 * i + j]) for j in range(4)]
    w9 = [rel(buf[600 + 4 * i + j]) for j in range(4)]
    mm = [rel(buf[1000 + 2 * i + j]) for j in range(2)]
    print(f"{i:3d} | {rel(buf[100 + i]):7d} | {w2[0]:7d} {w2[1]:7d} {w2[2]:7d} {w2[3]:7d} | {w9[0]:7d} {w9[1]:7d} {w9[2]:7d} {w9[3]:7d} | {mm[0]:7d} {mm[1]:7d}")
