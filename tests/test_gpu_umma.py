"""tcgen05 descriptor / layout helpers (csrc/umma.cuh) against torch.matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('b_mn', [0, 1])
@pytest.mark.parametrize('N,K', [(32, 16), (32, 32), (96, 32), (48, 48), (128, 32), (32, 128),
                                 (192, 48), (80, 128), (256, 64),
                                 (208, 32), (32, 208), (48, 208), (256, 48), (48, 256)])   # window-14 core
def test_umma_selftest(built_lib, N, K, b_mn):
    from hrfuser_b200 import _lib
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).bfloat16().cuda()
    Bm = torch.randn(N, K, generator=g).bfloat16().cuda()          # logical B: (N, K)
    Bdev = Bm.t().contiguous() if b_mn else Bm
    D = torch.full((128, N), float('nan'), device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(built_lib.hrf_selftest_umma(A.data_ptr(), Bdev.data_ptr(), D.data_ptr(), N, K, b_mn, st))
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t()
    err = (D - ref).abs().max().item()
    assert torch.isfinite(D).all() and err < 1e-3 * max(1.0, ref.abs().max().item()), \
        f'N={N} K={K} b_mn={b_mn}: max abs err {err}'
