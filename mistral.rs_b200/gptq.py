"""GPTQ / AWQ int4 linear — mirror of `GptqLayer` (REF mistralrs-quant/src/gptq/gptq_cuda.rs:
forward_raw :357-398, gptq_linear :451-623).  Differences from the reference, by design: the
checkpoint tensors are consumed as stored (no Marlin repack, no scale permutation, no argsort of
g_idx) by the tcgen05 dequant-GEMM; only bits == 4 is implemented.  Like the reference the
activations are computed in F16 (`quantized_act_type`) and tensor parallelism is rejected
(`distributed/layers.rs:776-788`)."""
import ctypes

import torch

from . import lib


class GptqLayer:
    def __init__(self, qweight, scales, qzeros=None, g_idx=None, bits=4, group_size=128, is_awq=False, bias=None,
                 world_size=1):
        if world_size > 1:
            raise ValueError("GPTQ/AWQ layers do not support tensor parallelism")
        if bits != 4:
            raise ValueError("only 4-bit GPTQ/AWQ is implemented")
        if not qweight.is_cuda:
            raise ValueError("GPTQ is only supported on CUDA")   # gptq_cpu.rs bails the same way
        self.is_awq = is_awq
        if is_awq:
            self.k, self.n = qweight.shape[0], qweight.shape[1] * 8
            if qzeros is None:
                raise ValueError("AWQ needs qzeros")
        else:
            self.k, self.n = qweight.shape[0] * 8, qweight.shape[1]
        if scales.shape != (self.k // group_size, self.n) or scales.dtype != torch.float16:
            raise ValueError(f"scales must be f16 [{self.k // group_size}, {self.n}]")
        self.qweight, self.scales, self.qzeros, self.g_idx = qweight.contiguous(), scales.contiguous(), qzeros, g_idx
        self.group_size, self.bias = group_size, bias

    def quantized_act_type(self):
        return torch.float16

    def forward(self, a: torch.Tensor) -> torch.Tensor:
        """QuantMethod::forward: cast to quantized_act_type, forward_raw, cast back."""
        orig = a.dtype
        y = self.forward_raw(a.to(torch.float16))
        return y.to(orig)

    def forward_raw(self, a: torch.Tensor) -> torch.Tensor:
        if not a.is_cuda:
            raise ValueError("Expected CUDA input to GptqLayer")
        if a.shape[-1] != self.k or a.dtype != torch.float16:
            raise ValueError("GptqLayer: bad input shape/dtype")
        x = a.reshape(-1, self.k).contiguous()
        out = torch.empty(x.shape[0], self.n, dtype=torch.float16, device=a.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
        rc = lib().mrs_gptq_gemm(P(x), P(self.qweight), P(self.scales), P(self.qzeros), P(self.g_idx), P(out),
                                 ctypes.c_int(x.shape[0]), ctypes.c_int(self.k), ctypes.c_int(self.n),
                                 ctypes.c_int(self.group_size), ctypes.c_int(int(self.is_awq)),
                                 ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"mrs_gptq_gemm failed with cudaError {rc}")
        out = out.reshape(*a.shape[:-1], self.n)
        return out + self.bias if self.bias is not None else out
