"""CPU-only tests of the host-side logic around the engine: derived-weight cache invalidation, the flat-buffer
layout of the training helpers, and the JSON line of bench.py's reference arm (no GPU, no compute through the C ABI)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_derived_weight_cache_follows_parameter_versions():
    """16-bit / transposed / -exp(A_log) copies are cached per parameter and rebuilt after an in-place update
    (optimizer.step, load_state_dict) or a raw-pointer update announced with torch._C._increment_version."""
    from aum_b200 import mixer
    p = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
    calls = []

    def make(t):
        calls.append(1)
        return t.to(torch.float16)

    a = mixer._cache.get(p, "w:test", make)
    b = mixer._cache.get(p, "w:test", make)
    assert a is b and len(calls) == 1
    with torch.no_grad():
        p.add_(1.0)                                   # what an optimizer does
    c = mixer._cache.get(p, "w:test", make)
    assert len(calls) == 2 and torch.equal(c.float(), p.detach())
    torch._C._increment_version([p])                  # what FlatAdam does after its kernel wrote through a raw pointer
    mixer._cache.get(p, "w:test", make)
    assert len(calls) == 3
    neg = mixer._neg_exp(torch.nn.Parameter(torch.zeros(4, 16)))
    assert torch.equal(neg, -torch.ones(4, 16))


def test_shadow_weight_policy_and_16bit_training_switch():
    """Host logic of two round-2 additions: (1) mixer._w / _w2d hand out the optimiser-maintained 16-bit shadow of a parameter
    only while its version counter has not moved since the shadow was written and the dtype matches; (2) the 16-bit
    delta / gradient-term policy of the training path needs 16-bit activations, 128-channel CTAs and no generic-kernel
    override."""
    from aum_b200 import autograd as AG, mixer
    w = torch.nn.Parameter(torch.randn(8, 16))
    assert mixer.shadow16(w, torch.bfloat16) is None
    sh = w.detach().to(torch.bfloat16)
    w._aum_w16, w._aum_w16_ver = sh, w._version
    assert mixer._w(w, torch.bfloat16) is sh and mixer.shadow16(w, torch.float16) is None
    assert mixer._w2d(w, torch.bfloat16).data_ptr() == sh.data_ptr()
    assert mixer._w(w, torch.bfloat16, pad_cols=24).shape == (8, 24)          # padded copies never come from the shadow
    with torch.no_grad():
        w.mul_(2.0)                                                           # somebody edits the parameter in place
    fresh = mixer._w(w, torch.bfloat16)
    assert fresh is not sh and torch.equal(fresh, w.detach().to(torch.bfloat16))
    assert AG._train16(torch.bfloat16, 1536) and AG._train16(torch.float16, 768)
    assert not AG._train16(torch.float32, 1536) and not AG._train16(torch.bfloat16, 192)
    os.environ["AUM_SCAN_BWD_GENERIC"] = "1"
    try:
        assert not AG._train16(torch.bfloat16, 1536)
    finally:
        del os.environ["AUM_SCAN_BWD_GENERIC"]


def test_flat_gradient_buffer_layout():
    """Every tensor of the flat gradient buffer starts on a 32-byte boundary, gradients alias it, zero() clears it."""
    from aum_b200 import dist as D
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(3, 5), (7,), (309,), (768, 2)]]
    red = D.FlatGradReducer(ps)
    assert all(o % D.FlatGradReducer.ALIGN == 0 for o in red.offsets)
    assert red.numel >= sum(p.numel() for p in ps)
    for p, o in zip(ps, red.offsets):
        assert p.grad.data_ptr() == red.flat.data_ptr() + 4 * o
    sum((p * p).sum() for p in ps).backward()
    for p in ps:
        torch.testing.assert_close(p.grad, 2 * p.detach())
    red.zero()
    assert float(red.flat.abs().sum()) == 0.0
    assert D.shard_range(10, 0, 4) == (0, 3) and D.shard_range(10, 3, 4) == (8, 10)


def test_audio_mamba_mirror_accepts_the_reference_constructor_call():
    """The reference builds its model with src/run.py:248-274's keyword set; the mirror takes exactly that call (options
    that cannot change the default forward are accepted and ignored) and refuses options that would change it."""
    import pytest
    from aum_b200.audio_mamba import AudioMamba
    m = AudioMamba(spectrogram_size=(128, 128), patch_size=(16, 16), strides=(16, 16), depth=2, embed_dim=96, num_classes=35,
                   imagenet_pretrain=False, imagenet_pretrain_path=None, imagenet_pretrain_modelkey="model",
                   aum_pretrain=False, aum_pretrain_path=None, aum_pretrain_fstride=16, aum_pretrain_tstride=16,
                   pt_hw_seq_len=None, bilinear_rope=False, drop_path_rate=0.0, imagenet_load_double_cls_token=False,
                   imagenet_load_middle_cls_token=True, use_double_cls_token=False, use_middle_cls_token=True,
                   use_end_cls_token=False, bimamba_type="v1", transpose_token_sequence=False, if_cls_token=True,
                   flexible_patch_sizes=None)
    assert len(m.layers) == 2 and m.layers[0].mixer.bimamba_type == "v1"
    for bad in (dict(bilinear_rope=True), dict(use_end_cls_token=True), dict(aum_pretrain=True), dict(drop_path_rate=0.1),
                dict(some_unknown_option=1)):
        with pytest.raises(NotImplementedError):
            AudioMamba(depth=1, embed_dim=96, **bad)


def test_weight_generation_invalidates_derived_caches():
    """The fused Adam kernel writes parameters through raw pointers; FlatAdam bumps the engine's weight generation, which
    every derived-weight cache entry (and AudioMamba's CUDA-graph key) is keyed on."""
    import torch
    from aum_b200 import mixer
    p = torch.nn.Parameter(torch.ones(4, 4))
    a = mixer._cache.get(p, "t:neg", lambda t: -t)
    assert mixer._cache.get(p, "t:neg", lambda t: -t) is a
    with torch.no_grad():
        p.data.view(-1)[0] = 5.0                      # a raw write: no version bump
    g0 = mixer.generation()
    assert mixer.bump_generation() == g0 + 1
    b = mixer._cache.get(p, "t:neg", lambda t: -t)
    assert b is not a and float(b[0, 0]) == -5.0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU path from oracle/_ref, one whole clip per step; the oracle
    port only where no staged reference exists): one JSON line with the same
    metric / unit as the GPU arm, impl = reference, a cpu_baseline describing the run and an e2e object of its own."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("clips/sec AuM-Base")
    assert d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    staged = any(os.path.isfile(os.path.join(r_, "src", "models", "mamba_models.py"))
                 for r_ in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")))
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    if staged:      # a step is one whole clip: the line's own clock agrees with its value
        assert abs(d["ms_per_step"] / 1000.0 * d["value"] - 1.0) < 1e-6 and d["config"]["clips_per_step"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank of a torchrun launch does no work and prints nothing
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                        capture_output=True, text=True, timeout=120, env=env)
    assert r2.returncode == 0 and r2.stdout.strip() == ""
