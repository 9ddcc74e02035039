"""GPU: tensor-core (tcgen05 / TMEM) building blocks."""
import pytest
import torch

import vfn_testutil as U  # noqa: F401
from vfnerf_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("N,K", [(256, 256), (16, 256), (256, 48), (224, 256), (128, 64), (64, 16)])
def test_umma_descriptor_conventions(built_lib, N, K):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    D = torch.zeros(128, N, device=DEV)
    _lib.check(built_lib.vfnerf_debug_umma_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, 0,
                                                torch.cuda.current_stream().cuda_stream), "debug_umma_gemm")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    err = (D - ref).abs().max().item()
    print(f"N={N} K={K}: max abs err {err:.3e} (ref magnitude {ref.abs().max().item():.1f})")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())      # fp32 accumulation-order noise only
