"""Whole-model logits parity: the AudioMamba mirror on the B200 engine vs golden logits produced by the real
reference (tiny configs) and vs the CPU oracle at AuM-Base size.  B200 only (-m gpu)."""
import pytest
import torch

import aum_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("case", ["fobi_tiny", "bibi_tiny", "fofo_tiny"])
def test_audio_mamba_matches_reference_golden(case):
    from aum_b200.audio_mamba import AudioMamba
    c = load_golden("audio_mamba_tiny.pt")[case]
    m = AudioMamba(**c["kwargs"]).to(DEV).eval()
    m.load_state_dict(c["state"], strict=True)          # the reference model's own state dict
    with torch.no_grad():
        logits = m(c["x"].to(DEV))
        feats = m(c["x"].to(DEV), return_features=True)
    # north-star tolerance: logits within rtol 1e-3 of the reference (fp32 tier is far inside it)
    torch.testing.assert_close(logits.cpu(), c["logits"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(feats.cpu(), c["features"], rtol=1e-3, atol=1e-5)
    m.act_dtype = torch.float16
    with torch.no_grad():
        l16 = m(c["x"].to(DEV)).cpu()
    err = (l16 - c["logits"]).abs().max() / c["logits"].abs().max()
    assert err < 1e-2, err


def test_aum_base_logits_vs_oracle_full_depth():
    """BASELINE config 2 model (AuM-Base Fo-Bi, depth 24, 527 classes, 128x1024 mel) on 1 clip.
    fp32 tier: rtol 1e-3 (north star).  fp16 tier (the benchmarked dtype): error reported against the logit scale."""
    from aum_b200.audio_mamba import AudioMamba
    sd = O.make_audio_mamba_state(768, 24, num_classes=527, seed=21, perturb_A=0.1)
    x = O.make_spectrogram(1, (128, 1024), seed=22)
    ref = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v1")
    m = AudioMamba(embed_dim=768, depth=24, num_classes=527, bimamba_type="v1").to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = m(x.to(DEV)).cpu()
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item())
    scale = ref.abs().max().item()
    e32 = (out - ref).abs().max().item() / scale
    m.act_dtype = torch.float16
    with torch.no_grad():
        o16 = m(x.to(DEV)).cpu()
    e16 = (o16 - ref).abs().max().item() / scale
    m.act_dtype = torch.bfloat16
    with torch.no_grad():
        ob16 = m(x.to(DEV)).cpu()
    eb16 = (ob16 - ref).abs().max().item() / scale
    print(f"AuM-Base logits: max|ref|={scale:.4f} rel-to-scale err fp32={e32:.2e} fp16={e16:.2e} bf16={eb16:.2e}")
    assert e32 < 1e-4 and e16 < 1e-2 and eb16 < 8e-2
    assert (o16.argmax(-1) == ref.argmax(-1)).all()


def test_aum_small_bibi_logits_vs_oracle_full_depth():
    """BASELINE config 4 model (AuM-Small Bi-Bi: embed 384, depth 24, bimamba v2, if_devide_out) on 1 clip of
    128x1024 mel, fp32 tier against the oracle (rtol 1e-3), 16-bit tiers against the logit scale."""
    from aum_b200.audio_mamba import AudioMamba
    sd = O.make_audio_mamba_state(384, 24, num_classes=527, bimamba_type="v2", seed=31, perturb_A=0.1)
    x = O.make_spectrogram(1, (128, 1024), seed=32)
    ref = O.audio_mamba_forward_oracle(sd, x, depth=24, bimamba_type="v2")
    m = AudioMamba(embed_dim=384, depth=24, num_classes=527, bimamba_type="v2").to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = m(x.to(DEV)).cpu()
    scale = ref.abs().max().item()
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-3 * scale)
    m.act_dtype = torch.float16
    with torch.no_grad():
        o16 = m(x.to(DEV)).cpu()
    assert (o16 - ref).abs().max().item() / scale < 1e-2
    assert (o16.argmax(-1) == ref.argmax(-1)).all()


@pytest.mark.parametrize("bt", ["v1", "v2"])
def test_cuda_graph_and_sequence_groups_match_eager(bt):
    """The path bench.py times — CUDA-graph replay with the batch split into two sequence groups on two streams,
    input taken from a pinned HOST buffer — returns what the plain eager forward returns (same kernels, same sizes
    per group: bit-identical), also with the groups serialised on one stream (bench.py's per-kernel timing pass)."""
    from aum_b200.audio_mamba import AudioMamba
    torch.manual_seed(5)
    kw = dict(embed_dim=192, depth=3, num_classes=35, spectrogram_size=(128, 256), bimamba_type=bt, act_dtype=torch.float16)
    plain = AudioMamba(**kw).to(DEV).eval()
    fast = AudioMamba(**kw, use_cuda_graph=True, micro_batches=2).to(DEV).eval()
    fast.load_state_dict(plain.state_dict(), strict=True)
    x = 0.5 * torch.randn(6, 256, 128)
    with torch.no_grad():
        halves = torch.cat([plain(xc.to(DEV)) for xc in x.chunk(2, dim=0)], dim=0)     # eager, group by group
        whole = plain(x.to(DEV))                                                     # eager, one group of 6
        for _ in range(3):                                                           # capture, then two replays
            g_out = fast(x.pin_memory())
        fast.serialize_groups = True
        fast.use_cuda_graph = False
        s_out = fast(x.to(DEV))
    assert torch.equal(g_out, halves)
    assert torch.equal(s_out, halves)
    torch.testing.assert_close(g_out, whole, rtol=2e-3, atol=2e-3)      # other batch grouping: same values up to fp16 tiling
    # a weight update invalidates the captured graph (derived 16-bit weights are rebuilt)
    with torch.no_grad():
        fast.use_cuda_graph = True
        fast.serialize_groups = False
        fast.head.weight.mul_(2.0); fast.head.bias.mul_(2.0)
        g2 = fast(x.pin_memory())
    torch.testing.assert_close(g2, 2.0 * g_out, rtol=1e-5, atol=1e-5)


def test_cold_cache_multi_stream_forward_matches_single_stream():
    """the FIRST eager forward of a model that splits its batch over two streams (derived-weight caches cold:
    16-bit copies, -exp(A_log), padded dt_proj are built by that very forward) must equal the single-stream result, and
    so must the first forward after a weight update."""
    from aum_b200.audio_mamba import AudioMamba
    torch.manual_seed(9)
    kw = dict(embed_dim=192, depth=4, num_classes=35, spectrogram_size=(128, 256), bimamba_type="v1", act_dtype=torch.float16)
    plain = AudioMamba(**kw).to(DEV).eval()
    x = 0.5 * torch.randn(8, 256, 128, device=DEV)
    with torch.no_grad():
        want = torch.cat([plain(xc) for xc in x.chunk(2, dim=0)], dim=0)
    for _ in range(3):                      # fresh models: every first call is a cold-cache call
        two = AudioMamba(**kw, micro_batches=2).to(DEV).eval()
        two.load_state_dict(plain.state_dict(), strict=True)
        with torch.no_grad():
            got_cold = two(x)
            got_warm = two(x)               # second call forks onto the two streams
            assert torch.equal(got_cold, want) and torch.equal(got_warm, want)
            two.head.weight.mul_(2.0); two.head.bias.mul_(2.0)      # in-place update: caches cold again
            got_upd = two(x)
        torch.testing.assert_close(got_upd, 2.0 * want, rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_run_on_the_tensors_device_not_the_current_one():
    """a model on cuda:1 while cuda:0 is current (the reference's kernels take a CUDAGuard from their tensors)."""
    from aum_b200.audio_mamba import AudioMamba
    torch.manual_seed(10)
    kw = dict(embed_dim=192, depth=2, num_classes=35, spectrogram_size=(128, 128), bimamba_type="v1", act_dtype=torch.float16)
    m0 = AudioMamba(**kw).to("cuda:0").eval()
    m1 = AudioMamba(**kw).to("cuda:1").eval()
    m1.load_state_dict(m0.state_dict(), strict=True)
    x = 0.5 * torch.randn(2, 128, 128)
    torch.cuda.set_device(0)
    with torch.no_grad():
        a = m0(x.to("cuda:0"))
        b = m1(x.to("cuda:1"))              # current device is still 0
    assert b.device.index == 1 and torch.equal(a.cpu(), b.cpu())
