"""ORACLE (test infrastructure): GPTQ / AWQ int4 linear as the reference's Marlin path computes it.
REF: mistralrs-quant/src/gptq/gptq_cuda.rs:357-398 (forward_raw -> marlin_matmul, activations cast
to F16), :451-623 (tensor shapes), kernels/marlin/marlin_kernel.cuh:93-140 (dequant: kU4B8 ->
w = (q - 8) * s, zero points ignored for GPTQ; kU4 + zero points for AWQ), and the exllama
fallback kernels/gptq/q_gemm.cu:1413-1447.  Parity status: UNPINNED against the reference kernel
(its Marlin tile repack + scale permutation are not restated here; the reference has no numeric
test for this path either) — the formulas are pinned only by the cited source lines."""
import numpy as np

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]  # nibble position i of an AWQ int32 holds column 8*c + AWQ_ORDER[i]


def pack_gptq(q, bits=4):
    """q [K, N] ints in 0..15 -> qweight [K/8, N] int32 (nibbles along K)."""
    K, N = q.shape
    out = np.zeros((K // 8, N), dtype=np.uint32)
    for j in range(8):
        out |= q[j::8].astype(np.uint32) << np.uint32(4 * j)
    return out.view(np.int32)


def pack_awq(q):
    """q [R, N] ints in 0..15 -> [R, N/8] int32 (nibbles along N in AWQ order)."""
    R, N = q.shape
    out = np.zeros((R, N // 8), dtype=np.uint32)
    for i, j in enumerate(AWQ_ORDER):
        out |= q[:, j::8].astype(np.uint32) << np.uint32(4 * i)
    return out.view(np.int32)


def dequant_gptq(qweight, scales, g_idx, group):
    qw = qweight.view(np.uint32)
    K = qw.shape[0] * 8
    q = np.empty((K, qw.shape[1]), dtype=np.int32)
    for j in range(8):
        q[j::8] = (qw >> np.uint32(4 * j)) & 0xF
    g = g_idx if g_idx is not None else np.arange(K) // group
    w = (q - 8).astype(np.float32) * scales.astype(np.float32)[g]
    return w.astype(np.float16).astype(np.float32)      # Marlin holds the dequantised weight in f16


def dequant_awq(qweight, scales, qzeros, group):
    def unpack(a):
        a = a.view(np.uint32)
        out = np.empty((a.shape[0], a.shape[1] * 8), dtype=np.int32)
        for i, j in enumerate(AWQ_ORDER):
            out[:, j::8] = (a >> np.uint32(4 * i)) & 0xF
        return out
    q, z = unpack(qweight), unpack(qzeros)
    g = np.arange(q.shape[0]) // group
    w = (q - z[g]).astype(np.float32) * scales.astype(np.float32)[g]
    return w.astype(np.float16).astype(np.float32)


def gemm(x_f16, w_deq):
    """f64 product of f16 activations and the f16-rounded dequantised weights -> [M, N]."""
    return x_f16.astype(np.float64) @ w_deq.astype(np.float64)
