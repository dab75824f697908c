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


@pytest.mark.parametrize("sh", [torch.bfloat16, torch.float16])
def test_adam_step_dev_counts_on_the_device_and_writes_the_16bit_shadow(sh):
    """aum_adam_step_dev == aum_adam_step with the step number read from (and incremented in) device memory, plus the
    parameters' 16-bit copy written in the same pass (what a captured training step replays)."""
    from aum_b200 import ops
    g = torch.Generator().manual_seed(5)
    n = 4099                                                        # a 3-element tail
    p0, gr = torch.randn(n, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    pa, ma, va = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb, mb, vb = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step_dev = torch.zeros((), device=DEV, dtype=torch.int32)
    p16 = torch.zeros(n, device=DEV, dtype=sh)
    for t in range(1, 5):
        ops.adam_step(pa, gr, ma, va, lr=1e-2, betas=(0.95, 0.999), weight_decay=5e-7, step=t)
        ops.adam_step_dev(pb, gr, mb, vb, step_dev, lr=1e-2, betas=(0.95, 0.999), weight_decay=5e-7, p16=p16)
        assert int(step_dev) == t
        torch.testing.assert_close(pb, pa, rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(vb, va, rtol=1e-6, atol=1e-12)
        assert torch.equal(p16, pb.to(sh))
    with pytest.raises(Exception):
        ops.adam_step_dev(pb, gr, mb, vb, torch.zeros((), device=DEV), lr=1e-2)           # the counter must be int32


def test_shadow_weights_follow_the_optimiser_and_yield_to_manual_edits():
    """FlatAdam(shadow_dtype=...): mixer._w hands out the optimiser-maintained 16-bit copy (no cast kernel) while the
    parameter is untouched since the last step, and falls back to a fresh cast once somebody edits it in place."""
    from aum_b200 import dist as D, mixer
    w = torch.nn.Parameter(torch.randn(16, 24, device=DEV))
    red = D.FlatGradReducer([w])
    opt = D.FlatAdam(red, lr=0.1, shadow_dtype=torch.bfloat16)
    assert mixer._w(w, torch.bfloat16).data_ptr() == opt.flat16.data_ptr()
    w.grad.fill_(1.0)
    opt.step()
    sh = mixer._w(w, torch.bfloat16)
    assert sh.data_ptr() == opt.flat16.data_ptr() and torch.equal(sh, w.detach().to(torch.bfloat16))
    assert mixer._w(w, torch.float16).dtype == torch.float16                                # other dtype: ordinary cast
    with torch.no_grad():
        w.mul_(2.0)                                                                          # behind the optimiser's back
    fresh = mixer._w(w, torch.bfloat16)
    assert fresh.data_ptr() != opt.flat16.data_ptr() and torch.equal(fresh, w.detach().to(torch.bfloat16))
    opt.resync_shadow()
    assert mixer._w(w, torch.bfloat16).data_ptr() == opt.flat16.data_ptr()
    tw = mixer._wT(w, torch.bfloat16)
    assert torch.equal(tw, w.detach().to(torch.bfloat16).t().contiguous())


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16])
def test_graphed_training_step_equals_the_eager_one(act):
    """TrainStep(cuda_graph=True): the replayed CUDA graph of the whole step (zero, forward, loss, backward, fused Adam with
    the device-side step counter) leaves the same parameters and losses as the eagerly launched step, step after step,
    with a different batch every time; capturing does not consume a training step of its own."""
    from aum_b200.audio_mamba import AudioMamba
    from aum_b200.trainer import TrainStep
    kw = dict(embed_dim=64, depth=2, num_classes=11, bimamba_type="v1", spectrogram_size=(32, 64), act_dtype=act)
    torch.manual_seed(21)
    ma = AudioMamba(**kw).to(DEV)
    mb = AudioMamba(**kw).to(DEV)
    mb.load_state_dict(ma.state_dict())
    ta = TrainStep(ma, lr=1e-3)
    tb = TrainStep(mb, lr=1e-3, cuda_graph=True)
    g = torch.Generator().manual_seed(22)
    for it in range(4):
        x = torch.randn(3, 64, 32, generator=g).to(DEV)
        y = (torch.rand(3, 11, generator=g) > 0.7).float().to(DEV)
        la, lb = ta(x, y), tb(x, y)
        assert tb.opt.t == ta.opt.t == it + 1 and int(tb.opt.step_dev) == it + 1
        tol = dict(rtol=1e-5, atol=1e-6) if act == torch.float32 else dict(rtol=2e-2, atol=2e-3)
        torch.testing.assert_close(lb.detach().float(), la.detach().float(), **tol)
        # the atomics of the weight-gradient reductions make two runs differ in the last bits; Adam's first steps turn
        # that into a few 1e-3 * lr, so compare the updates at that scale
        torch.testing.assert_close(tb.opt.flat_p, ta.opt.flat_p, rtol=0, atol=(2e-4 if act == torch.float32 else 2e-3))
    # an eager inference forward after replays sees the updated weights (derived-weight caches were invalidated)
    with torch.no_grad():
        oa, ob = ma(x), mb(x)
    torch.testing.assert_close(ob, oa, rtol=2e-2, atol=2e-2)


def test_train_step_static_loss_scale_is_divided_out_and_overflow_skips_the_step():
    """fp16 recipe (the reference trains with --mixed_precision=fp16 and accelerate's GradScaler): TrainStep(loss_scale=s)
    scales the loss before backward and the fused Adam divides it out - same update as the unscaled step where nothing
    over- or underflows (fp32 activations here, so the comparison is exact up to rounding) - and an overflowing scale
    leaves parameters, moments and the step count untouched."""
    from aum_b200.audio_mamba import AudioMamba
    from aum_b200.trainer import TrainStep
    kw = dict(embed_dim=64, depth=2, num_classes=11, bimamba_type="v1", spectrogram_size=(32, 64), act_dtype=torch.float32)
    torch.manual_seed(31)
    ma, mb, mc = (AudioMamba(**kw).to(DEV) for _ in range(3))
    mb.load_state_dict(ma.state_dict()); mc.load_state_dict(ma.state_dict())
    ta, tb, tc = TrainStep(ma, lr=1e-3), TrainStep(mb, lr=1e-3, loss_scale=1024.0), TrainStep(mc, lr=1e-3, loss_scale=float("inf"))
    g = torch.Generator().manual_seed(32)
    x = torch.randn(3, 64, 32, generator=g).to(DEV)
    y = (torch.rand(3, 11, generator=g) > 0.7).float().to(DEV)
    la, lb = ta(x, y), tb(x, y)
    torch.testing.assert_close(lb.detach(), la.detach(), rtol=1e-6, atol=1e-7)          # the returned loss is the unscaled one
    torch.testing.assert_close(tb.opt.flat_p, ta.opt.flat_p, rtol=0, atol=2e-4)
    assert tb.last_step_applied and tb.opt.t == 1
    before = tc.opt.flat_p.clone()
    tc(x, y)
    assert tc.last_step_applied is False and tc.opt.t == 0 and int(tc.opt.step_dev) == 0
    assert torch.equal(tc.opt.flat_p, before) and float(tc.opt.m.abs().sum()) == 0.0
