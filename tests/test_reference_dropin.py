"""Drop-in check of SURVEY.md section 8(b): the reference's OWN ``src/models/mamba_models.py`` (AudioMamba / Block /
create_block, unmodified - from /root/reference in the build container, from the staged copy oracle/_ref on the GPU
box) runs on top of this repo's ``mamba_ssm`` package: ``from mamba_ssm.modules.mamba_simple import Mamba`` and
``from mamba_ssm.ops.triton.layernorm import RMSNorm, layer_norm_fn, rms_norm_fn`` (mamba_models.py:18,26) resolve to
the B200 engine.  Each case runs in a child process (the CPU oracle tests register the reference's own modules under
the same package name)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    return any(os.path.isfile(os.path.join(r, "src", "models", "mamba_models.py"))
               for r in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")))


_CHILD = r'''
import contextlib, io, os, sys, warnings
ROOT = sys.argv[1]; mode = sys.argv[2]
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "audio-mamba-aum_b200")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import torch
import ref_loader
ns = ref_loader.load_reference_model_over_shim()
gold = torch.load(os.path.join(ROOT, "tests", "golden", "audio_mamba_tiny.pt"), map_location="cpu", weights_only=False)
for name, c in gold.items():
    kw = dict(c["kwargs"])
    with contextlib.redirect_stdout(io.StringIO()):
        # the reference's own constructor call shape (src/run.py:248-274)
        m = ns.AudioMamba(patch_size=(16, 16), strides=(16, 16), **kw)
    assert type(m.layers[0].mixer).__module__ == "aum_b200.modules", type(m.layers[0].mixer)
    sd = m.state_dict()
    assert set(sd.keys()) == set(c["state"].keys()), (name, set(sd.keys()) ^ set(c["state"].keys()))
    for k, v in c["state"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), (name, k)
    m.load_state_dict(c["state"], strict=True)
    if mode == "gpu":
        # the reference's patch embedding is an F.conv2d (src/utilities/tokenization.py:306): torch's cuDNN convolutions
        # default to TF32, which alone moves fp32 logits by ~1e-2; the golden logits are true fp32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        m = m.cuda().eval()
        with torch.no_grad():
            logits = m(c["x"].cuda()).float().cpu()                       # the reference's forward, this repo's kernels
        torch.testing.assert_close(logits, c["logits"], rtol=1e-3, atol=1e-5)  # north-star tolerance, fp32 tier
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):  # the reference's --mixed_precision=fp16
            l16 = m(c["x"].cuda()).float().cpu()
        err = (l16 - c["logits"]).abs().max().item() / c["logits"].abs().max().item()
        assert err < 1e-2, (name, err)
        if name == "fobi_tiny":         # and it trains: loss.backward() through the reference's Block + this repo's Functions
            m.train()
            out = m(c["x"].cuda())
            out.square().mean().backward()
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
print("DROPIN-OK", mode)
'''


def _run(mode):
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, mode], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(not _have_reference(), reason="reference sources neither at /root/reference nor staged under oracle/_ref")
def test_reference_audio_mamba_constructs_over_the_shim_with_reference_state_dict_keys():
    _run("cpu")


@pytest.mark.gpu
@pytest.mark.skipif(not _have_reference(), reason="reference sources neither at /root/reference nor staged under oracle/_ref")
def test_reference_audio_mamba_forward_over_the_shim_matches_golden_logits():
    _run("gpu")
