"""Stage the reference's own Python sources of the hot path under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (kaistmm/Audio-Mamba-AuM) is pure Python; its implementation of
the path that can run without the un-vendored CUDA wheels is its own ``selective_scan_ref`` / ``bimamba_inner_ref`` /
``Mamba`` / ``AudioMamba`` code.  /root/reference exists only in the build container, so this recipe copies the
few files those need, byte for byte, from where they lie into ``oracle/_ref/`` (same relative paths), which then
travels to the GPU box with the snapshot like a built ``.so`` does.  Nothing is copied into tracked files; the
product (``audio-mamba-aum_b200/``) never imports it.  Consumers: ``bench.py --impl reference`` and the
``cpu_baseline`` leg (the reference's CPU path timed on the box's host cores, ``kind: "reference"``) and
``tests/test_reference_dropin_*.py`` (the reference's own ``src/models/mamba_models.py`` running on top of this
repo's ``mamba_ssm`` shim).

    python oracle/build_ref.py          # idempotent; called by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("AUM_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")

FILES = [
    "vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py",
    "vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py",
    "vim-mamba_ssm/mamba_ssm/ops/triton/selective_state_update.py",
    "vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py",
    "src/models/__init__.py",
    "src/models/ast_models.py",
    "src/models/mamba_models.py",
    "src/utilities/__init__.py",
    "src/utilities/rope.py",
    "src/utilities/stats.py",
    "src/utilities/tokenization.py",
    "src/utilities/util.py",
]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copy FILES from the reference tree; returns False (and does nothing) when the tree is absent."""
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        if verbose:
            print(f"build_ref: {SRC_ROOT} not present; keeping whatever is staged under {DST_ROOT}")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    json.dump({"source": SRC_ROOT, "sha256": manifest}, open(os.path.join(DST_ROOT, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"build_ref: staged {len(FILES)} reference files under {DST_ROOT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or os.path.isdir(DST_ROOT) else 1)
