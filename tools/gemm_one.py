"""One GEMM shape in a loop (for ncu): python tools/gemm_one.py M N K [mode]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200 import _lib
lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
st = torch.cuda.current_stream().cuda_stream
A, B, C = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.empty(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
for _ in range(10):
    _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), K, 1, B.data_ptr(), K, 1, C.data_ptr(), N, bias.data_ptr(), 1, mode, st))
torch.cuda.synchronize()
import ctypes as ct
ph = (ct.c_longlong * 16)()
lib.hrp_debug_gemm_phases.argtypes = [ct.POINTER(ct.c_longlong)]
names = ["start", "setup done", "first loads issued", "first stage full", "loaders done", "accumulator ready",
         "tile staged", "tile written", "mma kb0 issued", "mma last issued"]
for rep in range(3):
    _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), K, 1, B.data_ptr(), K, 1, C.data_ptr(), N, bias.data_ptr(), 1, mode, st))
    lib.hrp_debug_gemm_phases(ph)
    t = list(ph)
    print(" | ".join(f"{n} {t[i] - t[0]}" for i, n in enumerate(names)))
