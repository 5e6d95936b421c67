"""GPU experiment: tcgen05 GEMM against torch fp32 (not part of the product path)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200 import _lib

lib = _lib.load()
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
for (M, N, K) in [(128, 128, 32), (128, 64, 64), (256, 128, 256), (4096, 128, 256), (4096, 51, 128), (4096, 204, 128), (1000, 32, 96)]:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.1
    bias = torch.randn(N, device="cuda")
    want = torch.relu(A.double() @ W.double().T + bias.double()).float()
    for split in (1, 3):
        Cout = torch.full((M, N), float("nan"), device="cuda")
        rc = lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()), C.c_void_p(Cout.data_ptr()),
                               M, N, K, K, K, N, 1, split, None)
        torch.cuda.synchronize()
        err = float((Cout - want).abs().max() / want.abs().max())
        print(f"M={M} N={N} K={K} split={split} rc={rc} rel_err={err:.3e} nan={int(torch.isnan(Cout).sum())}", flush=True)
# timing of the actor's fc1 shape, 8 agents' worth of rows
M, N, K = 32768, 128, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); Cout = torch.empty(M, N, device="cuda")
for split in (1, 3):
    for _ in range(3):
        lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None, C.c_void_p(Cout.data_ptr()), M, N, K, K, K, N, 0, split, None)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None, C.c_void_p(Cout.data_ptr()), M, N, K, K, K, N, 0, split, None)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print(f"split={split}: {us:.1f} us, {2 * M * N * K / us / 1e6:.1f} TFLOP/s (fp32-equivalent)")

print("fixed vs per-slab cost (M=32768, N=128, split=3):")
for K in (32, 64, 128, 256, 512):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda")
    for _ in range(3):
        lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None, C.c_void_p(Cout.data_ptr()), M, N, K, K, K, N, 0, 3, None)
    a.record()
    for _ in range(20):
        lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None, C.c_void_p(Cout.data_ptr()), M, N, K, K, K, N, 0, 3, None)
    b.record(); torch.cuda.synchronize()
    print(f"  K={K}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
