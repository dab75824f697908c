"""aum_gemm_wgrad (tcgen05, MN-major operands, split-K over tokens): dW += dY^T @ X against a float64 matmul of the
same (already rounded) operands.  Reference ops replaced: the einsums of selective_scan_interface.py:563,586,589 and
autograd's in_proj weight gradient.  B200 only (-m gpu)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (T tokens, No = dY channels, Ki = X channels, dY row pitch, X row pitch)
SHAPES = [
    (64, 128, 64, 128, 64),          # one tile, one k-block, BN = 64
    (300, 128, 128, 128, 128),       # token tail (300 = 4 x 64 + 44), BN = 128
    (1000, 80, 1536, 88, 1536),      # x_proj: 80 dY channels in an 88-wide buffer (second 64-channel box half empty)
    (513, 1536, 48, 1536, 56),       # dt_proj: X = first 48 columns of a 56-wide buffer, BN = 64, 12 row tiles
    (2000, 768, 1536, 768, 1536),    # out_proj shape, BN = 256, several token ranges
    (16416, 3072, 768, 3072, 768),   # in_proj at the config-3 size (32 sequences x 513 tokens): 72 tiles x split-K
    (130, 200, 330, 208, 336),       # every tail at once (No, Ki not multiples of 64, odd pitches)
    (5, 16, 8, 16, 8),               # tiny
]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("T,No,Ki,ldy,ldx", SHAPES)
def test_gemm_wgrad_tcgen05(dt, T, No, Ki, ldy, ldx):
    from aum_b200 import ops
    g = torch.Generator().manual_seed(T + No + Ki)
    dyb = torch.full((T, ldy), float("nan"), dtype=dt)       # pad columns hold NaN: they must never be read into the sum
    xb = torch.full((T, ldx), float("nan"), dtype=dt)
    dyb[:, :No] = torch.randn((T, No), generator=g).to(dt)
    xb[:, :Ki] = torch.randn((T, Ki), generator=g).to(dt)
    dy, x = dyb.to(DEV)[:, :No], xb.to(DEV)[:, :Ki]
    ref = (dyb[:, :No].double().t() @ xb[:, :Ki].double())
    init = torch.randn((No, Ki), generator=g)
    dw = init.clone().to(DEV)
    ops.gemm_wgrad(dy, x, dw)                                # accumulates
    torch.cuda.synchronize()
    tol = 8e-5 * (T ** 0.5) + 1e-5            # fp32 accumulation of T exact products whose sum is O(sqrt(T)); split-K order varies
    torch.testing.assert_close(dw.cpu().double(), ref + init.double(), rtol=2e-5, atol=tol)
    # a second call accumulates again
    ops.gemm_wgrad(dy, x, dw)
    torch.testing.assert_close(dw.cpu().double(), 2 * ref + init.double(), rtol=2e-5, atol=2 * tol)


def test_gemm_wgrad_into_a_view_of_a_flat_buffer_and_fp32_tier():
    """dW as a (32-byte aligned) view into a flat gradient buffer, as the trainer passes it; fp32 operands take the
    transpose + CUDA-core route and must agree."""
    from aum_b200 import ops
    g = torch.Generator().manual_seed(3)
    T, No, Ki = 700, 96, 192
    dy32, x32 = torch.randn((T, No), generator=g), torch.randn((T, Ki), generator=g)
    flat = torch.zeros(8 + No * Ki + 8, device=DEV)
    dw = flat[8:8 + No * Ki].view(No, Ki)
    ops.gemm_wgrad(dy32.to(DEV).bfloat16(), x32.to(DEV).bfloat16(), dw)
    ref = dy32.bfloat16().double().t() @ x32.bfloat16().double()
    torch.testing.assert_close(dw.cpu().double(), ref, rtol=1e-5, atol=1e-3)
    assert (flat[:8] == 0).all() and (flat[-8:] == 0).all()
    dw32 = torch.zeros((No, Ki), device=DEV)
    ops.gemm_wgrad(dy32.to(DEV), x32.to(DEV), dw32)
    torch.testing.assert_close(dw32.cpu().double(), dy32.double().t() @ x32.double(), rtol=1e-4, atol=1e-3)


def test_gemm_wgrad_rejects_bad_arguments():
    from aum_b200 import ops, _lib as L
    a = torch.zeros((64, 64), device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.AumError):
        ops.gemm_wgrad(a, a[:32], torch.zeros((64, 64), device=DEV))          # token counts differ
    with pytest.raises(L.AumError):
        ops.gemm_wgrad(a, a, torch.zeros((64, 64), device=DEV, dtype=torch.bfloat16))   # dW must be fp32
