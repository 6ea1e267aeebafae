"""Time the tcgen05 GEMM on the MLP's shapes (CUDA events; L2-warm back-to-back and L2-flushed)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(name, M, N, K, akc, bkc, mode=3, iters=30):
    A = torch.randn((M, K) if akc else (K, M), device="cuda")
    B = torch.randn((N, K) if bkc else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    sam, sak = (K, 1) if akc else (1, M)
    sbn, sbk = (K, 1) if bkc else (1, N)
    call = lambda: _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), sam, sak, B.data_ptr(), sbn, sbk, C.data_ptr(), N, None, 1, mode, st))
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / iters * 1e3
    cold = 0.0
    for _ in range(iters):
        flush.zero_(); e0.record(); call(); e1.record(); torch.cuda.synchronize(); cold += e0.elapsed_time(e1) * 1e3
    print(f"{name:28s} M={M} N={N} K={K} mode={mode}: warm {warm:6.2f} us  cold {cold/iters:6.2f} us  ({2*M*N*K/warm/1e6:.1f} TFLOP/s warm)")
for mode in (3, 1):
    run("forward L1", 4096, 256, 60, True, True, mode)
    run("forward hidden", 4096, 256, 256, True, True, mode)
    run("dgrad", 4096, 256, 256, True, False, mode)
    run("forward hidden 16k rows", 16384, 256, 256, True, True, mode)
    run("forward H=512", 4096, 512, 512, True, True, mode)
