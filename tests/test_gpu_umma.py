"""GPU test (-m gpu) of the hand-written tcgen05 plumbing: UMMA descriptors, TMEM, mbarrier commit, and the
shifted-window operand addressing the implicit-GEMM convolutions rely on."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N,shift,rows", [(16, 32, 0, 128), (32, 32, 0, 136), (32, 32, 1, 136), (64, 32, 3, 136),
                                            (32, 224, 2, 130), (288, 32, 0, 128), (80, 64, 5, 140)])
def test_umma_selftest(K, N, shift, rows):
    from crfp_b200 import _lib as L
    g = torch.Generator().manual_seed(K * 1000 + N + shift)
    A = torch.randn(rows, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = A[shift:shift + 128].float() @ B.float().t()
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().crfp_selftest_umma(rows, K, N, shift, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), st), "selftest")
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    print(f"K={K} N={N} shift={shift}: max-abs {err:.3e}")
    assert err < 1e-3 * max(1.0, K / 32)


@pytest.mark.parametrize("K,N,shift,sbo", [(32, 32, 0, 10), (64, 224, 11, 10), (32, 64, 22, 10), (32, 32, 3, 12)])
def test_umma_selftest_row_group_stride(K, N, shift, sbo):
    """A-operand SBO other than 128 B: row m reads record shift + (m/8)*sbo + m%8 — the addressing of an 8-pixel-wide
    tile inside a wider halo tile (fused align kernel: sbo = 10 records = 160 B)."""
    from crfp_b200 import _lib as L
    rows = shift + 15 * sbo + 8
    g = torch.Generator().manual_seed(K * 1000 + N + shift + sbo)
    A = torch.randn(rows, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    idx = torch.tensor([shift + (m // 8) * sbo + m % 8 for m in range(128)])
    ref = A[idx].float() @ B.float().t()
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().crfp_selftest_umma_sbo(rows, K, N, shift, sbo, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), st), "selftest sbo")
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    print(f"K={K} N={N} shift={shift} sbo={sbo}: max-abs {err:.3e}")
    assert err < 1e-3 * max(1.0, K / 32)
