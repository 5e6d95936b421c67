"""GPU experiment: stand-alone time of the tcgen05 weight-gradient GEMM (maexp_tc_gemm_tn, csrc/tc_gemm.cu) per launch, back to
back, for the shapes of a C51 team update (one agent-network per launch)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200 import _lib
lib = _lib.load()
for (M, N, K) in [(204, 128, 256), (128, 256, 256), (64, 128, 384), (128, 256, 64)]:
    A = torch.randn((K, M)).cuda(); B = torch.randn((K, N)).cuda(); out = torch.zeros((M, N), device="cuda")
    def run():
        _lib.check(lib.maexp_tc_gemm_tn(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(out.data_ptr()), None, 0, M, N, K, M, N, N, None, None))
    for _ in range(10): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): run()
    b.record(); torch.cuda.synchronize()
    print(M, N, K, "slabs", K // 32, "%.2f us per launch (back to back, includes launch overhead)" % (a.elapsed_time(b) * 5))
