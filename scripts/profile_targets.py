"""Small launch sets for `ncu --set full` captures (one target per invocation)."""
import sys, os
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
g.load_package()
import oracle
from mistralrs_b200 import mmq, quant, model as M
dev = torch.device("cuda:0")
what = sys.argv[1]
rng = np.random.default_rng(0)
if what == "mmq":
    dtype, Mm, N, K = sys.argv[2], 4096, 4096, 4096
    wb = oracle.random_blocks(dtype, N * K // oracle.BLOCK_ELEMS[dtype], rng)
    w = quant.QTensor(torch.from_numpy(wb.reshape(-1)).to(dev), dtype, (N, K))
    x = torch.randn(Mm, K, device=dev).to(torch.bfloat16)
    for _ in range(4):
        mmq.forward(w, x)
    torch.cuda.synchronize()
elif what == "decode":
    cfg = M.LlamaConfig.llama3_8b(); cfg.n_layers = 4
    w = M.LlamaWeights(cfg, dev)
    run = M.LlamaRunner(w, batch=1, max_ctx=400, pdl=True)
    run.context_lens.fill_(255)
    for _ in range(3):
        run.step()
    torch.cuda.synchronize()
elif what == "w4a16":
    import ctypes
    from mistralrs_b200 import lib
    N, K, Mm, group = 28672, 4096, 32, 128
    qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev)
    sc = torch.rand(K // group, N, device=dev).to(torch.float16) * 0.01
    tiles = torch.empty(K // 16, N * 2, dtype=torch.int32, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lambda: torch.cuda.current_stream().cuda_stream
    lib().gptq_marlin_repack(P(qw), ctypes.c_void_p(0), P(tiles), K, N, 4, ctypes.c_int64(st()))
    x = torch.randn(Mm, K, device=dev).to(torch.float16)
    y = torch.empty(Mm, N, dtype=torch.float16, device=dev)
    for _ in range(6):
        assert lib().mrs_w4a16_gemm(P(x), P(tiles), P(sc), ctypes.c_void_p(0), P(y), Mm, K, N, group, 0, 0, ctypes.c_void_p(st())) == 0
    torch.cuda.synchronize()
elif what == "prefill_attn":      # optional second argument "mma": keep the call on the mma.sync kernel
    import ctypes
    from mistralrs_b200 import lib, paged_attn
    lib().mrs_prefill_attn_tc_debug(ctypes.c_int32(0 if (len(sys.argv) > 2 and sys.argv[2] == "mma") else 1), ctypes.c_uint32(0), ctypes.c_uint32(0))
    T, H, KVH, D = 4096, 32, 8, 128
    q = torch.randn(T, H, D, device=dev).to(torch.bfloat16)
    k = torch.randn(T, KVH, D, device=dev).to(torch.bfloat16)
    v = torch.randn(T, KVH, D, device=dev).to(torch.bfloat16)
    for _ in range(3):
        paged_attn.prefill_attention(q, k, v, D ** -0.5)
    torch.cuda.synchronize()
elif what == "affine":            # packed-affine GGUF GEMM (a5): optional source type, default q4_k
    from mistralrs_b200 import packed_affine as PA
    dtype = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
    Mm, N, K = 4096, 4096, 4096
    wb = oracle.random_blocks(dtype, N * K // oracle.BLOCK_ELEMS[dtype], rng)
    packed = PA.PackedAffine(torch.from_numpy(wb.reshape(-1)).to(dev), dtype, (N, K), torch.bfloat16)
    x = torch.randn(Mm, K, device=dev).to(torch.bfloat16)
    for _ in range(4):
        packed.forward(x)
    torch.cuda.synchronize()
