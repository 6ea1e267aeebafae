"""Time the two tcgen05 GEMM kernels on the MLP's shapes (CUDA events): the TMA-fed kernel on pre-split operands
(hrp_gemm_tma.cu) against the register-staged one (hrp_mlp_tc.cu), back to back (programmatic dependent launch chains
the launches, L2 warm) and one at a time; then the TMA kernel's phase clocks of CTA (0,0,0)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
lib.hrp_debug_tma_gemm.argtypes = [i32, i32, i32, vp, vp, i64, i64, vp, vp, i64, i64, vp, vp, i32, vp, i32, i32, i32, vp]
lib.hrp_debug_split_lo.argtypes = [vp, vp, i64, vp]
lib.hrp_debug_tma_gemm_clocks.argtypes = [i32, vp]


def timeit(call, iters=50):
    for _ in range(5):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        call()
    e1.record(); torch.cuda.synchronize()
    chain = e0.elapsed_time(e1) / iters * 1e3
    one = 0.0
    for _ in range(iters):
        e0.record(); call(); e1.record(); torch.cuda.synchronize(); one += e0.elapsed_time(e1) * 1e3
    return chain, one / iters


def run(name, M, N, K, akc, bkc, splits=1, bn=0, lo_out=True):
    A = torch.randn((M, K) if akc else (K, M), device="cuda")
    B = torch.randn((N, K) if bkc else (K, N), device="cuda")
    A_lo, B_lo = torch.empty_like(A), torch.empty_like(B)
    lib.hrp_debug_split_lo(A.data_ptr(), A_lo.data_ptr(), A.numel(), st)
    lib.hrp_debug_split_lo(B.data_ptr(), B_lo.data_ptr(), B.numel(), st)
    Cm, C_lo = torch.empty(splits, M, N, device="cuda"), torch.empty(splits, M, N, device="cuda")
    bias = torch.randn(N, device="cuda")
    sam, sak = (K, 1) if akc else (1, M)
    sbn, sbk = (K, 1) if bkc else (1, N)
    new = lambda: _lib.check(lib.hrp_debug_tma_gemm(M, N, K, A.data_ptr(), A_lo.data_ptr(), sam, sak, B.data_ptr(), B_lo.data_ptr(),
                                                    sbn, sbk, Cm.data_ptr(), C_lo.data_ptr() if lo_out and splits == 1 else None, N,
                                                    bias.data_ptr() if splits == 1 else None, 1 if splits == 1 else 0, splits, bn, st))
    os.environ["HRP_NO_TMA"] = "1"
    old = lambda: _lib.check(lib.hrp_gemm_strided(M, N, K, A.data_ptr(), sam, sak, B.data_ptr(), sbn, sbk, Cm.data_ptr(), N,
                                                  bias.data_ptr(), 1, 3, st))
    t_old = timeit(old) if splits == 1 else (float("nan"), float("nan"))
    t_new = timeit(new)
    lib.hrp_debug_tma_gemm_clocks(1, None)
    new()
    clk = np.zeros(16, dtype=np.int64)
    lib.hrp_debug_tma_gemm_clocks(0, clk.ctypes.data)
    d = clk - clk[0]
    flops = 2.0 * M * N * K
    print(f"{name:30s} M={M} N={N} K={K} splits={splits} bn={bn or 'auto'}: TMA kernel chain {t_new[0]:6.2f} us, alone {t_new[1]:6.2f} us "
          f"({flops / t_new[0] / 1e6:.0f} TFLOP/s useful) | register-staged chain {t_old[0]:6.2f}, alone {t_old[1]:6.2f}")
    print(f"    phase clocks (cycles from kernel start, CTA 0): setup {d[1]}, predecessor done {d[2]}, first stage landed {d[3]}, "
          f"second {d[4]}, last MMA issued {d[5]}, accumulator complete {d[6]}, tile staged {d[7]}, stores drained {d[8]}")


run("forward L1", 4096, 256, 60, True, True)
run("forward hidden", 4096, 256, 256, True, True)
run("forward hidden, bn 128", 4096, 256, 256, True, True, bn=128)
run("forward [actor|critic]", 4096, 512, 256, True, True)
run("forward [actor|critic], bn 64", 4096, 512, 256, True, True, bn=64)
run("d(h2): K = 2H, W as MN-major B", 4096, 256, 512, True, False)
run("d(h1)", 4096, 256, 256, True, False)
run("[dWa1;dWc1] split 16", 512, 256, 4096, False, False, splits=16)
run("dW2 split 16", 256, 256, 4096, False, False, splits=16)
run("dW1 split 16", 256, 60, 4096, False, False, splits=16)
run("forward hidden H=512", 4096, 512, 512, True, True)
