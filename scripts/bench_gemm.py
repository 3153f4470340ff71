"""Micro-benchmark of the tcgen05 GEMM primitive (iisan_gemm_bf16) on the SAN shapes; CUDA-event timing."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iisan_b200 import _lib
lib = _lib.load()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
def run(M, N, K, a_mn=False, splitk=1, bias=True, out_bf16=True, out_f32=True, iters=50, rot=4):
    A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(rot)]
    B = torch.randn(N, K, device="cuda").bfloat16()
    if a_mn:
        A = [a.t().contiguous() for a in A]; B = B.t().contiguous()
    bv = torch.randn(N, device="cuda") if bias else None
    of = [torch.zeros(M, N, device="cuda") for _ in range(rot)] if out_f32 else [None] * rot
    ob = [torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(rot)] if out_bf16 else [None] * rot
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def go(i):
        a = A[i % rot]
        _lib.check(lib.iisan_gemm_bf16(M, N, K, p(a), a.stride(0), int(a_mn), p(B), B.stride(0), int(a_mn), p(of[i % rot]), N,
                                       p(ob[i % rot]), N, p(bv), 0, splitk, st), "gemm")
    for i in range(5): go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): go(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    byt = M * K * 2 + N * K * 2 + (M * N * 4 if out_f32 else 0) + (M * N * 2 if out_bf16 else 0)
    print(f"M={M} N={N} K={K} mn={int(a_mn)} splitk={splitk} f32={int(out_f32)} bf16={int(out_bf16)}: {us:8.1f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s  {byt/us/1e3:8.1f} GB/s")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
run(5632, 768, 64, iters=n)
run(5632, 768, 64, out_f32=False, iters=n)
run(5632, 64, 768, out_f32=False, iters=n)
run(5632, 768, 768, out_f32=False, iters=n)
run(768, 64, 5632, a_mn=True, splitk=15, bias=False, out_bf16=False, iters=n)
run(8192, 8192, 8192, out_f32=False, iters=max(3, n // 10), rot=1)
