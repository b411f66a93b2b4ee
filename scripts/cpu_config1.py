"""BASELINE config 1 (TinyLlama-1.1B Q4_K_M, CPU, 128-token prompt + 64 generated): the CPU port of
the reference's CPU path (oracle/: Q8_K / Q8_0 activations + integer block dots) end to end — full
layer stack incl. RMSNorm, RoPE and attention — on synthetic weights of that shape.
usage: python scripts/cpu_config1.py [prompt_len] [gen_len]     (plumbing + CPU tok/s; no GPU)"""
import os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import __graft_entry__ as g
g.load_package()
from mistralrs_b200 import model as M
from oracle.model import OracleLlama

prompt_len = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gen_len = int(sys.argv[2]) if len(sys.argv) > 2 else 64
threads = len(os.sched_getaffinity(0))
cfg = M.LlamaConfig.tinyllama()
t0 = time.perf_counter()
host = {}
H, I = cfg.hidden, cfg.inter
nq, nkv = cfg.n_heads * cfg.head_dim, cfg.n_kv_heads * cfg.head_dim
from mistralrs_b200 import BLOCK_ELEMS
for l in range(cfg.n_layers):
    for name, rows, cols in (("attn_q", nq, H), ("attn_k", nkv, H), ("attn_v", nkv, H), ("attn_output", H, nq),
                             ("ffn_gate", I, H), ("ffn_up", I, H), ("ffn_down", H, I)):
        dt = M.tensor_type(cfg, name, l)
        host[(l, name)] = M.synth_blocks(dt, rows * cols // BLOCK_ELEMS[dt], M.tensor_seed(l, name)).reshape(-1)
    for name in ("attn_norm", "ffn_norm"):
        rng = np.random.Generator(np.random.PCG64(M.tensor_seed(l, name)))
        host[(l, name)] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
for name in ("token_embd", "output"):
    dt = M.tensor_type(cfg, name, 0)
    host[(0, name)] = M.synth_blocks(dt, cfg.vocab * H // BLOCK_ELEMS[dt], M.tensor_seed(0, name)).reshape(-1)
rng = np.random.Generator(np.random.PCG64(M.tensor_seed(0, "output_norm")))
host[(0, "output_norm")] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
cos, sin = M.rope_tables(cfg)
ref = OracleLlama(cfg, host, M.tensor_type, cos, sin, "f32", cpu_path=True, threads=threads)
print(f"weights: {sum(v.nbytes for v in host.values())/1e9:.2f} GB synthetic in {time.perf_counter()-t0:.1f}s; threads={threads}", flush=True)
toks = [1000 + (131 + i) % 2048 for i in range(prompt_len)]      # bench.rs prompt pattern, case 0
t0 = time.perf_counter()
for pos, t in enumerate(toks):                                    # token-by-token prompt (CPU path has no batched prefill here)
    logits = ref.step([t], pos)
t_prompt = time.perf_counter() - t0
nxt = int(np.argmax(logits[0]))
t0 = time.perf_counter()
for i in range(gen_len - 1):
    logits = ref.step([nxt], prompt_len + i)
    nxt = int(np.argmax(logits[0]))
t_gen = time.perf_counter() - t0
print(f"config1 TinyLlama-1.1B Q4_K_M CPU port: prompt {prompt_len} tok in {t_prompt:.2f}s ({prompt_len/t_prompt:.2f} tok/s), "
      f"decode {(gen_len-1)/t_gen:.2f} tok/s over {gen_len-1} tokens, threads={threads}")
