import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as g
g.load_package()
import oracle
from mistralrs_b200 import kv_index, paged_attn
from test_paged_attn_gpu import _setup
from util import to_dev
cuda = torch.device("cuda:0")
for (S, H, KVH, D, BS, ctx) in [(1, 8, 8, 128, 16, [16]), (1, 8, 8, 128, 16, [40]), (1, 32, 8, 128, 16, [384])]:
    dt = "bf16"
    q, kct, vct, ku, vu, bt, tables = _setup(cuda, S, H, KVH, D, BS, ctx, "hnd", dt, seed=1)
    scale = 1.0 / np.sqrt(D)
    want = oracle.paged_attention(q, ku, vu, bt, ctx, KVH, D, BS, scale, 1, dt)
    indptr, indices, last = kv_index.make_paged_kv_tensors(tables, ctx, BS, bt.size)
    req, tile, o_indptr, chunk, mask = kv_index.make_paged_kv_decode_tensors(tables, ctx, BS, None, S)
    print("indptr", indptr, "indices", indices[:6], "last", last, "req", req, "tile", tile, "o_indptr", o_indptr, "chunk", chunk, "mask", mask)
    d = lambda a: to_dev(np.ascontiguousarray(a), cuda)
    out = paged_attn.flashinfer_decode(to_dev(q, cuda, dt), kct, vct, d(indptr), d(indices), d(last), d(req), d(tile), d(o_indptr), d(chunk), d(mask), scale).float().cpu().numpy()
    err = np.abs(out - want)
    print(ctx, "max err", err.max(), "per head", err.max(axis=2)[0][:8])
    # dense torch reference
    kd, vd = paged_attn.gather_kv_cache_flashinfer(kct, vct, d(bt), d(np.array([0, ctx[0]], dtype=np.int32)), ctx[0], torch.bfloat16)
    qf = to_dev(q, cuda, dt).float()[0]
    ref = torch.stack([torch.softmax((kd.float()[:, h // (H // KVH)] @ qf[h]) * scale, 0) @ vd.float()[:, h // (H // KVH)] for h in range(H)])
    print("   torch-vs-kernel", (ref.cpu().numpy() - out[0]).__abs__().max(), "torch-vs-oracle", np.abs(ref.cpu().numpy() - want[0]).max())
