"""Does tcgen05 kind::tf32 truncate or round fp32 operands?  Single-pass GEMM on full-precision inputs compared
with float64 references built from truncated and from round-to-nearest-even TF32 operands."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200 import _lib
lib = _lib.load()
M, N, K = 256, 128, 64
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, generator=g, device="cuda")
B = torch.randn(N, K, generator=g, device="cuda")
C = torch.empty(M, N, device="cuda")
_lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), K, 1, B.data_ptr(), K, 1, C.data_ptr(), N, None, 0, 1,
                                torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
def trunc(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)
def rne(x):
    i = x.view(torch.int32).to(torch.int64)
    i = i + 0x0FFF + ((i >> 13) & 1)
    return (i & ~0x1FFF).to(torch.int32).view(torch.float32)
for name, f in (("truncate", trunc), ("round-nearest-even", rne)):
    ref = f(A).double() @ f(B).double().t()
    print(name, "max abs diff", float((C.double() - ref).abs().max()))
print("vs exact fp32 inputs", float((C.double() - A.double() @ B.double().t()).abs().max()))
