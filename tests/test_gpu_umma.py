"""tcgen05 GEMM primitive vs a plain PyTorch fp32 reference of the same product on bf16-rounded inputs.
Tolerance: fp32 accumulation of exact bf16 products -> 1e-3 relative of the row scale (order of summation only);
bf16 outputs add one rounding (2^-8 relative)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(M, N, K, a_mn, b_mn, bias=False, relu=False, splitk=1, out_bf16=False, seed=0):
    from iisan_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bvec = torch.randn(N, device="cuda", generator=g) if bias else None
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bvec
    if relu:
        ref = ref.relu()
    A_s = A.t().contiguous() if a_mn else A
    B_s = B.t().contiguous() if b_mn else B
    out = torch.zeros(M, N, device="cuda")
    outb = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if out_bf16 else None
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.iisan_gemm_bf16(M, N, K, p(A_s), A_s.stride(0), int(a_mn), p(B_s), B_s.stride(0), int(b_mn), p(out), N, p(outb), N,
                             p(bvec), int(relu), splitk, st)
    _lib.check(rc, "iisan_gemm_bf16")
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item() / scale
    assert err < 1e-3, f"fp32 out rel err {err}"
    if out_bf16:
        errb = (outb.float() - ref).abs().max().item() / scale
        assert errb < 1e-2, f"bf16 out rel err {errb}"
    return err


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 64, 768), (256, 768, 64), (5632, 64, 768), (5632, 768, 64),
                                   (5632, 768, 768), (88, 64, 768), (200, 24, 96), (130, 256, 192), (128, 128, 128)])
def test_umma_gemm_k_major(M, N, K):
    _gemm(M, N, K, False, False, bias=True, relu=False, out_bf16=True)


@pytest.mark.parametrize("M,N,K", [(64, 768, 5632), (768, 64, 5632), (64, 64, 128), (128, 256, 200), (768, 768, 5632), (24, 96, 300)])
def test_umma_gemm_mn_major_splitk(M, N, K):
    _gemm(M, N, K, True, True, splitk=1)
    _gemm(M, N, K, True, True, splitk=8)


def test_umma_gemm_relu_bias():
    _gemm(512, 64, 768, False, False, bias=True, relu=True, out_bf16=True)


@pytest.mark.parametrize("M,N,K", [(5632, 768, 768), (5632, 768, 64), (130, 256, 192), (128, 64, 64), (300, 24, 96)])
def test_umma_gemm_data_gradient_majors(M, N, K):
    """A K-major x B MN-major: dx = dy W with the nn.Linear weight [out = K, in = N] read in place."""
    _gemm(M, N, K, False, True, out_bf16=True)


def test_umma_gemm_persistent_many_tiles():
    """More tiles than SMs: every CTA walks several tiles through its two TMEM accumulators."""
    _gemm(8192, 1024, 256, False, False, bias=True, out_bf16=True)
    _gemm(4096, 2048, 128, True, True, splitk=2)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_umma_gemm_multicast_clusters(a_mn, b_mn):
    """>= 4 tiles of 128 x 256 per SM: clusters of two CTAs, the column tile is multicast (odd row-tile count included)."""
    _gemm(8192, 2560, 128, a_mn, b_mn, bias=not a_mn, out_bf16=not a_mn)
    _gemm(8064, 2560, 192, a_mn, b_mn)          # 63 row tiles: the last pair has a phantom tile
