"""tcgen05 building blocks on the B200: descriptor encodings, TMEM operand layout, bulk copy + mbarrier protocol."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (32, 128), (128, 16), (144, 128), (176, 128)])
@pytest.mark.parametrize("npass", [1, 3])
def test_tc_gemm_selftest(K, N, npass):
    from pharmacoforge_b200 import _lib
    from pharmacoforge_b200.weights import split_bf16, umma_b_image
    lib = _lib.load()
    gen = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=gen)
    W = torch.randn(N, K, generator=gen) / K ** 0.5
    hi, lo = split_bf16(W)
    img_hi, img_lo = umma_b_image(hi).cuda(), umma_b_image(lo).cuda()
    Ad = A.cuda().contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pf_tc_selftest(C.c_void_p(Ad.data_ptr()), C.c_void_p(img_hi.data_ptr()),
                                  C.c_void_p(img_lo.data_ptr()), C.c_void_p(D.data_ptr()), K, N, npass, s),
               "pf_tc_selftest")
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t())
    err = (D.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if npass == 3 else 2e-2
    assert err <= tol * scale, f"K={K} N={N} npass={npass}: max abs err {err:.3e} (scale {scale:.2f})"
    if npass == 3:   # the compensated product must be far better than a single bf16 pass
        assert err <= 1e-3 * scale
