"""Dev (GPU box): a handful of launches of the int4 swap-AB kernel at one Mistral shape (for ncu)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.load_package()
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
for _ in range(6):
    rc = lib().mrs_w4a16_gemm(P(x), P(tiles), P(sc), ctypes.c_void_p(0), P(y), M, K, N, group, 0, 0, ctypes.c_void_p(st()))
    assert rc == 0
torch.cuda.synchronize()
print("ok")
