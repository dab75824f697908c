"""Fused flat-buffer Adam (aum_adam_step) against torch.optim.Adam — the reference's optimiser recipe
(/root/reference/src/traintest.py:32-34: betas=(0.95, 0.999), weight_decay=5e-7).  B200 only (-m gpu)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("wd", [0.0, 5e-7, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    from aum_b200 import dist as D
    g = torch.Generator().manual_seed(11)
    shapes = [(7, 5), (1536, 16), (3,), (768, 33), (1,)]          # odd sizes: unaligned views, a 4-element tail
    ref_params = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    our_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    ref_opt = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.95, 0.999), weight_decay=wd)
    red = D.FlatGradReducer(our_params)
    opt = D.FlatAdam(red, lr=1e-3, betas=(0.95, 0.999), weight_decay=wd)
    for it in range(5):
        grads = [torch.randn(s, generator=g).to(DEV) * (10.0 ** (it - 2)) for s in shapes]
        for p, q, gr in zip(ref_params, our_params, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)                # gradients live in the reducer's flat buffer
        ref_opt.step()
        v0 = our_params[0]._version
        opt.step()
        assert our_params[0]._version > v0          # derived-weight caches see the in-place update
        for p, q in zip(ref_params, our_params):
            torch.testing.assert_close(q.detach(), p.detach(), rtol=2e-6, atol=1e-7)
    # parameters are views into the one flat buffer
    assert all(q.data_ptr() >= opt.flat_p.data_ptr() for q in our_params)


def test_adam_grad_scale_and_empty():
    from aum_b200 import ops
    p = torch.ones(1000, device=DEV)
    gr = torch.full((1000,), 4.0, device=DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ops.adam_step(p, gr, m, v, lr=0.1, betas=(0.9, 0.999), step=1, grad_scale=0.25)       # effective gradient 1
    torch.testing.assert_close(p, torch.full_like(p, 0.9), rtol=1e-6, atol=1e-6)
    e = torch.empty(0, device=DEV)
    ops.adam_step(e, e, e, e, lr=0.1)


def test_flat_adam_skips_a_step_with_non_finite_gradients():
    """The fp16 recipe's GradScaler behaviour (accelerate --mixed_precision=fp16, which the reference's scripts use):
    an inf / NaN anywhere in the flat gradient skips the update - p, m, v and the step count untouched."""
    from aum_b200 import dist as D
    ps = [torch.nn.Parameter(torch.ones(33, device=DEV)), torch.nn.Parameter(torch.ones(5, 7, device=DEV))]
    red = D.FlatGradReducer(ps)
    opt = D.FlatAdam(red, lr=0.1)
    ps[0].grad.fill_(1.0); ps[1].grad.fill_(1.0)
    ps[1].grad[2, 3] = float("inf")
    before = opt.flat_p.clone()
    assert opt.step(grad_scale=0.5, skip_nonfinite=True) is False
    assert torch.equal(opt.flat_p, before) and opt.t == 0 and float(opt.m.abs().sum()) == 0.0
    ps[1].grad[2, 3] = 1.0
    assert opt.step(grad_scale=0.5, skip_nonfinite=True) is True
    assert opt.t == 1 and not torch.equal(opt.flat_p, before)
