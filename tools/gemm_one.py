"""One GEMM shape in a loop (for ncu): python tools/gemm_one.py M N K [mode]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200 import _lib
lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
akc = int(sys.argv[5]) if len(sys.argv) > 5 else 1   # 1: A is [M, K] (K-contiguous), 0: A is [K, M]
bkc = int(sys.argv[6]) if len(sys.argv) > 6 else 1
st = torch.cuda.current_stream().cuda_stream
A = torch.randn((M, K) if akc else (K, M), device="cuda")
B = torch.randn((N, K) if bkc else (K, N), device="cuda")
C = torch.empty(M, N, device="cuda")
sam, sak = (K, 1) if akc else (1, M)
sbn, sbk = (K, 1) if bkc else (1, N)
bias = torch.randn(N, device="cuda")
for _ in range(10):
    _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), sam, sak, B.data_ptr(), sbn, sbk, C.data_ptr(), N, bias.data_ptr(), 1, mode, st))
torch.cuda.synchronize()
import ctypes as ct
ph = (ct.c_longlong * 16)()
lib.hrp_debug_gemm_phases.argtypes = [ct.POINTER(ct.c_longlong)]
names = ["start", "setup done", "first loads issued", "first stage full", "loaders done", "accumulator ready",
         "tile staged", "tile written", "mma kb0 issued", "mma last issued", "L4 before wait", "L4 stage free", "L4 stashed",
         "L4 arrived", "M4 stage full", "M4 committed"]
for rep in range(3):
    _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), sam, sak, B.data_ptr(), sbn, sbk, C.data_ptr(), N, bias.data_ptr(), 1, mode, st))
    lib.hrp_debug_gemm_phases(ph)
    t = list(ph)
    print(" | ".join(f"{n} {t[i] - t[0]}" for i, n in enumerate(names)))
