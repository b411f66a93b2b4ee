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
