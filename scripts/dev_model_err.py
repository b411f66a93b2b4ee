import sys, os
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
g.load_package()
import oracle
from oracle.model import OracleLlama
from mistralrs_b200 import model as M
cuda = torch.device("cuda:0")
for quant in ("q8_0", "q4_k", "q6_k"):
  for dtn, tdt in (("bf16", torch.bfloat16), ("f16", torch.float16)):
    for nl in (1, 2):
        cfg = M.LlamaConfig.tiny_test(quant=quant, n_layers=nl)
        w = M.LlamaWeights(cfg, cuda, dtype=tdt, keep_host=True)
        run = M.LlamaRunner(w, batch=2, max_ctx=64)
        cos, sin = M.rope_tables(cfg)
        ref = OracleLlama(cfg, w.host, M.tensor_type, cos, sin, dtn)
        toks = [17, 900]; run.set_tokens(toks); errs = []
        for pos in range(4):
            run.step(); torch.cuda.synchronize()
            got = run.logits().float().cpu().numpy(); want = ref.step(toks, pos)
            errs.append(float(np.abs(got - want).max() / np.abs(want).max()))
            toks = np.argmax(want, axis=1).tolist(); run.set_tokens(toks)
        srt = np.sort(want[0])[::-1]
        print(quant, dtn, "layers", nl, "rel err", ["%.2e" % e for e in errs], "scale %.3g top-gap %.3g finite %s" % (np.abs(want).max(), srt[0]-srt[1], np.isfinite(want).all()), flush=True)
