#!/bin/bash
# First GPU session for the packed-affine path: parity, then a timing next to the block-format GEMM, then ncu.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_run_affine.sh'
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzz_affine_gpu.py -q -rxX 2>&1 | tail -60 > gpurun_out/affine_tests.log
timeout 300 python - > gpurun_out/affine_timing.log 2>&1 <<'PY'
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g; g.load_package()
import oracle
from mistralrs_b200 import packed_affine as PA, mmq, quant
dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
M = N = K = 4096
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
def timed(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for dt in ("q4_k", "q8_0", "q6_k", "q4_0"):
    wb = oracle.random_blocks(dt, N * K // oracle.BLOCK_ELEMS[dt], rng)
    w = quant.QTensor(torch.from_numpy(wb.reshape(-1)).to(dev), dt, (N, K))
    p = PA.PackedAffine(w.data, dt, (N, K), torch.bfloat16)
    ta, tb = timed(lambda: p.forward(x)), timed(lambda: mmq.forward(w, x))
    fl = 2.0 * M * N * K
    print(f"{dt}: packed-affine {ta:.3f} ms ({fl / ta / 1e9:.0f} TFLOP/s)   block GEMM {tb:.3f} ms ({fl / tb / 1e9:.0f} TFLOP/s)")
PY
timeout 600 ncu --set full --clock-control none -k regex:mmq_tc_kernel -c 1 -o gpurun_out/affine_q4k python scripts/profile_targets.py affine q4_k > gpurun_out/affine_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/affine_q4k.ncu-rep > gpurun_out/affine_q4k.summary.txt 2>&1 && rm -f gpurun_out/affine_q4k.ncu-rep
