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
