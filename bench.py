#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric : clips/s, AuM-Base Fo-Bi (92.1 M params, depth 24, d_model 768) forward on 128-mel x 1024-frame
         spectrograms (-> L = 513 tokens), batch 64 PER GPU (weak scaling), fp16 activations / fp32 residual
         stream and scan state (the reference runs `--mixed_precision=fp16`, exps/*/**.sh).
value  : whole-job clips/s with inputs resident in HBM.
e2e    : the same through the public API with HOST buffers: pinned spectrogram batch -> H2D -> forward ->
         D2H logits, all inside the timed region.
roofline / cpu_baseline: see DESIGN.md sections 5-6.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "audio-mamba-aum_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "clips/sec AuM-Base 128x1024 mel fwd"
UNIT = "clips/s"
CFG = dict(embed_dim=768, depth=24, num_classes=527, spectrogram_size=(128, 1024), bimamba_type="v1")
BATCH_PER_GPU = 64
SEED = 3949


def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class _StdoutToStderr:
    """NCCL prints its version banner on fd 1 during initialisation; keep stdout clean for the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


SCAN_TRAFFIC_FILE = os.path.join("profiles", "r2_scan_traffic.json")


def scan_traffic(sequences_per_launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE scan launch of exactly this size, from the committed
    `ncu --set full` capture of the kernel (tools/ncu_scan_traffic.sh writes the file).  Only used when the capture
    was taken on the same launch size and on the kernel source that is in the tree now (sha256 of scan_fwd_tma.cu
    recorded next to the numbers); otherwise null - a DRAM byte count cannot be read outside a profiler."""
    try:
        import hashlib
        t = json.load(open(os.path.join(ROOT, SCAN_TRAFFIC_FILE)))
        src = os.path.join(ROOT, "audio-mamba-aum_b200", "csrc", "scan_fwd_tma.cu")
        if t.get("kernel_source_sha256") != hashlib.sha256(open(src, "rb").read()).hexdigest():
            return None, "null: committed ncu capture predates the current scan_fwd_tma.cu"
        if int(t.get("sequences_per_launch", -1)) != sequences_per_launch:
            return None, "null: committed ncu capture is of another launch size"
        return int(t["dram__bytes_read.sum"] + t["dram__bytes_write.sum"]), SCAN_TRAFFIC_FILE
    except Exception:
        return None, "null: no ncu capture committed"


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
_REF_MODEL = {}


def reference_clip_forward():
    """ONE whole clip through the reference's OWN code on the host CPU: the unmodified ``AudioMamba`` of
    src/models/mamba_models.py (config 2: AuM-Base Fo-Bi, depth 24) whose mixer runs the reference's own
    ``bimamba_inner_ref`` / ``selective_scan_ref`` (selective_scan_interface.py:673-709, 86-152), imported from the
    staged copy ``oracle/_ref`` (oracle/build_ref.py; the CUDA wheels it would otherwise call cannot run on sm_100).
    fp32, all the host threads torch will use.  Returns (seconds, threads)."""
    import contextlib
    import io
    import warnings
    import ref_loader
    if "m" not in _REF_MODEL:
        # the reference's per-token ops are small: beyond a few tens of threads intra-op oversubscription makes it SLOWER
        # (measured on the 128-core GPU box, profiles/r2_reference_threads.txt); AUM_REF_THREADS overrides
        torch.set_num_threads(int(os.environ.get("AUM_REF_THREADS", 0)) or max(1, min(os.cpu_count() or 1, 16)))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = ref_loader.load_reference_model()
            torch.manual_seed(SEED)
            with contextlib.redirect_stdout(io.StringIO()):
                m = ns.AudioMamba(spectrogram_size=CFG["spectrogram_size"], patch_size=(16, 16), strides=(16, 16),
                                  depth=CFG["depth"], embed_dim=CFG["embed_dim"], num_classes=CFG["num_classes"],
                                  bimamba_type=CFG["bimamba_type"]).eval()
        g = torch.Generator().manual_seed(SEED)
        _REF_MODEL["m"] = m
        _REF_MODEL["x"] = 0.5 * torch.randn(1, CFG["spectrogram_size"][1], CFG["spectrogram_size"][0], generator=g)
        _REF_MODEL["src"] = ref_loader.REF_KIND
    with torch.no_grad():
        t0 = time.perf_counter()
        out = _REF_MODEL["m"](_REF_MODEL["x"])
        dt_ = time.perf_counter() - t0
    assert tuple(out.shape) == (1, CFG["num_classes"]) and bool(torch.isfinite(out).all())
    return dt_, torch.get_num_threads()


def cpu_reference_sample(clips: int = 2):
    """cpu_baseline of the main arm: `clips` whole clips (one at a time) through reference_clip_forward()."""
    try:
        reference_clip_forward()                      # builds the model, first-touch
        ts = [reference_clip_forward() for _ in range(clips)]
    except Exception as e:                            # staged reference missing: fall back to the oracle port
        r = cpu_port_sample(blocks=2)
        r["note"] = f"oracle/_ref unavailable ({type(e).__name__}: {e}); oracle port timed instead"
        return r
    sec = statistics.median([t for t, _ in ts])
    return {"value": 1.0 / sec, "unit": UNIT, "cores": ts[0][1], "kind": "reference",
            "sample": f"{clips} whole clips (batch 1 each, median) through the reference's own AudioMamba + bimamba_inner_ref "
                      f"(+1 untimed warm-up clip), all 24 blocks, fp32, {_REF_MODEL['src']}",
            "seconds_per_clip": sec}


def cpu_port_sample(blocks: int = 2, repeats: int = 1):
    """Fallback only (no staged reference): the oracle port of selective_scan_ref / bimamba_inner_ref /
    AudioMamba.forward, fp32.  Bounded sample: ONE clip through the front end and `blocks` of
    the 24 blocks; whole-model time extrapolated linearly in the block count (every block is identical work)."""
    import aum_oracle as O
    # the oracle's per-token ops are small: beyond ~16 threads intra-op oversubscription makes it SLOWER
    # (measured on the 128-core GPU box: 28.6 s/block with 128 threads vs ~5 s with 8-16), so use min(cores, 16)
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    sd = O.make_audio_mamba_state(CFG["embed_dim"], blocks, num_classes=CFG["num_classes"],
                                  spectrogram_size=CFG["spectrogram_size"], bimamba_type=CFG["bimamba_type"],
                                  seed=SEED, perturb_A=0.1)
    x = O.make_spectrogram(1, CFG["spectrogram_size"], seed=SEED)
    best_full, best_front = None, None
    for _ in range(repeats):
        with torch.no_grad():
            t0 = time.perf_counter()
            O.audio_mamba_forward_oracle(sd, x, depth=blocks, bimamba_type=CFG["bimamba_type"], n_blocks=0)
            t1 = time.perf_counter()
            O.audio_mamba_forward_oracle(sd, x, depth=blocks, bimamba_type=CFG["bimamba_type"], n_blocks=blocks)
            t2 = time.perf_counter()
        front, full = t1 - t0, t2 - t1
        if best_full is None or full < best_full:
            best_full, best_front = full, front
    per_block = max(best_full - best_front, 1e-9) / blocks
    t_clip = best_front + CFG["depth"] * per_block
    return {"value": 1.0 / t_clip, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 clip x (front end + {blocks} of {CFG['depth']} blocks), extrapolated x{CFG['depth']}/{blocks} "
                      f"in blocks; oracle port of the reference's selective_scan_ref/bimamba_inner_ref/AudioMamba.forward, fp32",
            "seconds_per_block_per_clip": per_block}


REFERENCE_ARM_BUDGET_S = 330.0


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, timed on the box's host cores.
    One step = ONE WHOLE CLIP through the unmodified reference model (reference_clip_forward: all 24 blocks, nothing
    extrapolated): the bounded sample of the batch-64 workload is its batch size, 1 of 64 clips - every clip is the
    same work and the CPU path has no cross-clip reuse, so clips/s is what is reported.  W warm-up + K timed steps are
    run as asked, bounded by a wall-clock budget (REFERENCE_ARM_BUDGET_S): at ~5-10 s per clip the driver's 5 + 20
    steps fit; if the budget runs out first, `steps` reports the number actually timed.  Rank 0 only."""
    if rank != 0:
        return
    t_start = time.perf_counter()
    secs, threads, w_done = [], 1, 0
    try:
        for i in range(args.warmup + args.steps):
            if i >= 1 and time.perf_counter() - t_start > REFERENCE_ARM_BUDGET_S and (i < args.warmup or len(secs) >= 3):
                if i < args.warmup:
                    continue                  # out of budget during warm-up: go straight to the timed steps
                break
            t, threads = reference_clip_forward()
            if i >= args.warmup:
                secs.append(t)
            else:
                w_done += 1
        kind, src, note = "reference", _REF_MODEL.get("src", "?"), None
    except Exception as e:
        r = cpu_port_sample(blocks=1)
        secs, threads, kind, src = [1.0 / r["value"]], r["cores"], "port", "oracle port"
        note = f"oracle/_ref unavailable ({type(e).__name__}: {e}): oracle port, 1 clip x 1 of 24 blocks extrapolated"
    sec = statistics.median(secs)
    v = 1.0 / sec
    sample = (f"1 whole clip per step (1 of the 64 clips of a batch; all 24 blocks, nothing extrapolated) through the "
              f"reference's own AudioMamba + bimamba_inner_ref / selective_scan_ref, fp32, {src}")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(secs),
           "steps_requested": args.steps, "warmup": w_done, "ms_per_step": 1000.0 * sec,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "AuM-Base Fo-Bi forward (BASELINE configs[1]): 128x1024 mel -> 513 tokens, depth 24, "
                                  "d_model 768, d_state 16; CPU sample: batch 1 per step",
                      "clips_per_step": 1},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": note or sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t_start}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------
def train_leg(dev, world, dist, rank, steps=6, warmup=3, batch=32):
    """BASELINE configs[2]: AuM-Base Fo-Bi TRAINING step, bf16 activations, VGGSound-shape inputs (128x1024 mel, 309
    classes), batch 32 per GPU (256 at 8 GPUs), one process per GPU: forward + backward + the ONE gradient all-reduce
    (flat fp32 buffer, launched in 3 reverse-layer-order chunks from autograd hooks while backward runs) + fused Adam.
    Every rank runs it (the all-reduce is the path's only collective); device-timed, max over ranks."""
    from aum_b200 import _lib
    from aum_b200.audio_mamba import AudioMamba
    from aum_b200.trainer import TrainStep
    torch.manual_seed(SEED)
    model = AudioMamba(embed_dim=768, depth=24, num_classes=309, bimamba_type="v1", act_dtype=torch.bfloat16).to(dev)
    g = torch.Generator().manual_seed(SEED + 100 + rank)
    with torch.no_grad():
        for blk in model.layers:
            blk.mixer.A_log.add_(0.1 * torch.randn(blk.mixer.A_log.shape, generator=g).to(dev))
            blk.mixer.A_b_log.add_(0.1 * torch.randn(blk.mixer.A_b_log.shape, generator=g).to(dev))
    ts = TrainStep(model, lr=1e-5, n_chunks=3, cuda_graph=os.environ.get("AUM_TRAIN_GRAPH", "1") == "1")
    x = (0.5 * torch.randn(batch, 1024, 128, generator=g)).to(dev)
    y = (torch.rand(batch, 309, generator=g) > 0.97).float().to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # an eagerly launched pass first: counts the step's launches and all-reduce pieces and times the exposed part of the
    # all-reduce with CUDA events (which a captured graph cannot hold); it doubles as extra warm-up
    ts.timing = []
    for _ in range(2):              # (cold: allocator, NCCL buffers, kernel attributes)
        ts(x, y)
    barrier()
    ts.timing = []
    l0, a0 = _lib.launch_count(), ts.reducer.async_launches
    for _ in range(2):
        ts(x, y)
    barrier()
    launches = (_lib.launch_count() - l0) // 2
    pieces = (ts.reducer.async_launches - a0) // 2
    exposed_local = sum(a.elapsed_time(b) for a, b in ts.timing) / 2
    ts.timing = None
    for _ in range(warmup):
        ts(x, y)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = ts(x, y)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps, exposed_local], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, exposed = t[0].item(), t[1].item()
    # what the same all-reduce costs when nothing overlaps it (one piece, after backward)
    alone = None
    if dist is not None:
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(3):
            dist.all_reduce(ts.reducer.flat, op=dist.ReduceOp.AVG)
        s1.record()
        barrier()
        alone = s0.elapsed_time(s1) / 3
    out = {"metric": "clips/sec AuM-Base training step (fwd+bwd+grad all-reduce+Adam)", "value": world * batch / (ms / 1e3),
           "unit": "clips/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "dtype": "bf16",
           "config": {"workload": f"BASELINE configs[2]: AuM-Base Fo-Bi, depth 24, 309 classes, 128x1024 mel, batch {batch}/GPU "
                                  f"(global {world * batch}), BCE-with-logits, Adam betas (0.95, 0.999) wd 5e-7, fp32 master weights"},
           "allreduce": {"bytes": ts.reducer.numel * 4, "pieces_per_step": pieces, "launched_from_backward_hooks": max(pieces - 1, 0),
                         "exposed_ms": exposed, "alone_ms": alone,
                         "note": "exposed = device time between the end of backward and the gradient buffer being final"},
           "launch_mode": ("cuda-graph replay of the whole step (forward, loss, backward with the all-reduce pieces, Adam with a "
                           "device-side step counter)") if ts.cuda_graph else "eager launches",
           "gpu_launches": launches, "loss": loss.item(), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    del ts, model
    torch.cuda.empty_cache()
    return out


def small_bibi_leg(dev, world, dist, steps=10, warmup=3, batch=32):
    """BASELINE configs[3]: AuM-Small Bi-Bi (embed 384, depth 24, bimamba v2, if_devide_out) forward, 128x1024 mel,
    batch 32 per GPU (128 at 4 GPUs), fp16, CUDA-graph replay with two sequence groups - replicas, no collective."""
    from aum_b200.audio_mamba import AudioMamba
    torch.manual_seed(SEED)
    model = AudioMamba(embed_dim=384, depth=24, num_classes=527, bimamba_type="v2", act_dtype=torch.float16,
                       use_cuda_graph=True, micro_batches=2).to(dev).eval()
    g = torch.Generator().manual_seed(SEED + 200)
    with torch.no_grad():
        for blk in model.layers:
            blk.mixer.A_log.add_(0.1 * torch.randn(blk.mixer.A_log.shape, generator=g).to(dev))
            blk.mixer.A_b_log.add_(0.1 * torch.randn(blk.mixer.A_b_log.shape, generator=g).to(dev))
    x = (0.5 * torch.randn(batch, 1024, 128, generator=g)).to(dev)
    with torch.no_grad():
        for _ in range(warmup):
            model(x)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model(x)
        e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    n_params = sum(p.numel() for p in model.parameters())
    del model
    torch.cuda.empty_cache()
    return {"metric": "clips/sec AuM-Small Bi-Bi 128x1024 mel fwd", "value": world * batch / (ms / 1e3), "unit": "clips/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "dtype": "fp16",
            "config": {"workload": f"BASELINE configs[3]: AuM-Small Bi-Bi forward, depth 24, d_model 384, {n_params} params, "
                                   f"batch {batch}/GPU (global {world * batch})"}}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="aum_b200")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="clips per GPU (BASELINE config 2: 64)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--micro-batches", type=int, default=2, help="independent sequence groups on separate CUDA streams")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--e2e-per-call", action="store_true", help="e2e through one blocking model(x_host) call per step instead of infer_stream")
    ap.add_argument("--no-extra", action="store_true", help="skip the training-step (configs[2]) and AuM-Small Bi-Bi (configs[3]) legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        with _StdoutToStderr():
            try:        # NCCL kernels on a high-priority stream: the gradient all-reduce pieces launched from backward
                        # hooks then get SMs while the backward kernels of earlier layers are still queued
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                dist.init_process_group("nccl", device_id=dev, pg_options=opts)
            except Exception:
                dist.init_process_group("nccl", device_id=dev)
            dist.barrier()          # forces communicator creation (and NCCL's banner) now
            torch.cuda.synchronize()

    from aum_b200 import _lib, ops
    from aum_b200.audio_mamba import AudioMamba

    act = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    torch.manual_seed(SEED)
    model = AudioMamba(**CFG, act_dtype=act, use_cuda_graph=not args.no_graph,
                       micro_batches=args.micro_batches).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(SEED)
    with torch.no_grad():   # move A off its structured S4D-real init, as trained weights are
        for blk in model.layers:
            blk.mixer.A_log.add_(0.1 * torch.randn(blk.mixer.A_log.shape, generator=g).to(dev))
            blk.mixer.A_b_log.add_(0.1 * torch.randn(blk.mixer.A_b_log.shape, generator=g).to(dev))
    n_params = sum(p.numel() for p in model.parameters())
    B = args.batch
    F_, T_ = CFG["spectrogram_size"]
    x_host = (0.5 * torch.randn(B, T_, F_, generator=g)).pin_memory()
    x_dev = x_host.to(dev)
    logits_host = torch.empty((B, CFG["num_classes"]), dtype=torch.float32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            return model(x_dev)

    def step_e2e():
        with torch.no_grad():
            xd = x_host if model.use_cuda_graph else x_host.to(dev, non_blocking=True)   # graph path copies H2D itself
            out = model(xd)
            logits_host.copy_(out, non_blocking=True)

    # ---- warm-up (also builds the 16-bit weight copies once)
    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- timed region 1: HBM-resident inputs (per-forward working set ~27 GB >> 126 MB L2)
    sampler = ClockSampler(local_rank)
    launches_per_step = None
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- timed region 2: end to end from host buffers through the public streaming-inference call
    # (AudioMamba.infer_stream: pinned host batch -> H2D on a copy stream, double-buffered against the previous batch's
    # forward -> logits -> D2H into pinned memory; every step's copies are inside the timed region)
    def run_e2e(n):
        with torch.no_grad():
            for out in model.infer_stream(x_host for _ in range(n)):
                logits_host.copy_(out, non_blocking=True)
    if args.e2e_per_call:
        run_e2e = lambda n: [step_e2e() for _ in range(n)]      # one model(x_host) call per step, nothing overlapped
    run_e2e(2)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (t2.item() / args.steps / 1e3)

    # ---- roofline of the dominant kernel (bidirectional scan): CUDA-event pairs around every scan launch of an
    # instrumented, eagerly-launched repeat of the same K steps (kernels inside a replayed CUDA graph cannot be
    # bracketed by events); also counts this repo's kernel launches per step.
    # The pass launches the SAME kernels on the SAME sizes, but with the two sequence groups on one stream, so that each
    # launch runs alone: a launch duration measured while another stream's kernels share the SMs is not the kernel's.
    calls0 = _lib.launch_count()
    ops.PROFILE = []
    _lib.PROFILE_ALL = []
    model.serialize_groups = True
    barrier()
    for _ in range(args.steps):
        step_resident()
    barrier()
    model.serialize_groups = False
    prof, ops.PROFILE = ops.PROFILE, None
    prof_all, _lib.PROFILE_ALL = _lib.PROFILE_ALL, None
    launches = (_lib.launch_count() - calls0)
    # the scan's share of the step: its event-bracketed time over that of every kernel this repo launched in the pass
    # (the quantity the committed ncu launch list gives as well)
    all_ms = sum(s.elapsed_time(e) for (_, s, e) in prof_all) if prof_all else 0.0
    scan_all_ms = sum(s.elapsed_time(e) for (n, s, e) in prof_all if n == "aum_selective_scan_fwd") if prof_all else 0.0
    peak, peak_src = peaks()
    scan_ms = [s.elapsed_time(e) for (name, s, e) in prof if name == "selective_scan"] if prof else []
    Lq = (F_ // 16) * (T_ // 16) + 1
    Di, Nst = 2 * CFG["embed_dim"], 16
    mb = args.micro_batches if (args.micro_batches > 1 and B % args.micro_batches == 0 and B >= 2 * args.micro_batches) else 1
    M = (B // mb) * Lq            # rows one scan launch processes (one micro-batch)
    s_act = 2
    # ALGORITHMIC bytes of one fused bidirectional scan launch, SURVEY.md section 8(d):
    #   s*(4*B*D*L + 2*B*N*L) + 4*(2*D*N + 2*D)   (u, delta, z in, out written, B and C rows; A, A_b, D, delta_bias)
    # with every activation counted at the activation size s = 2.  The inference path stores delta in the activation dtype
    # (AUM_DELTA_16BIT, default on), so u, delta, z and out move exactly these bytes; the packed B|C rows stay fp32 (2 x the
    # formula's 2BNL term, 0.5 % of the launch): `bytes_this_build_moves` counts them (and delta, when kept fp32) as stored.
    alg_bytes = s_act * (4 * M * Di + 2 * M * Nst) + 4 * (2 * Di * Nst + 2 * Di)
    from aum_b200 import mixer as _mixer
    delta_sz = s_act if _mixer._DELTA_16BIT else 4
    build_bytes = M * Di * (3 * s_act + delta_sz) + M * 2 * Nst * 4 + 4 * (2 * Di * Nst + 2 * Di)
    roof = None
    if scan_ms:
        avg = sum(scan_ms) / len(scan_ms)
        ach = alg_bytes / (avg * 1e-3) / 1e9
        traffic, traffic_src = scan_traffic(B // mb)
        n_exp = M * Di * Nst * 2
        roof = {"kernel": "scan_fwd_tma_kernel (fused forward+reverse selective scan)", "bound": "hbm", "achieved": ach,
                "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "avg_launch_ms": avg, "launches_timed": len(scan_ms),
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_formula": "SURVEY 8(d): s*(4BDL+2BNL)+4*(2DN+2D), s=2",
                "bytes_this_build_moves": build_bytes, "delta_dtype": "activation dtype (16-bit)" if delta_sz == 2 else "fp32",
                "share_of_step": (scan_all_ms / all_ms) if all_ms > 0 else None, "sequences_per_launch": B // mb,
                # the ceiling that actually binds this kernel: one MUFU.EX2 per (token, channel, state, direction);
                # peak = 15.8 results/clk/SM measured (profiles/r1_microbench_pipe_rates.txt) x 148 SMs x max SM clock
                "xu_pipe": {"achieved_Texp_per_s": n_exp / (avg * 1e-3) / 1e12,
                            "peak_Texp_per_s": 15.8 * 148 * 1.965e9 / 1e12,
                            "frac": (n_exp / (avg * 1e-3) / 1e12) / (15.8 * 148 * 1.965e9 / 1e12)},
                "note": "16 ex2 per (token,channel,direction): MUFU-bound before HBM-bound (xu_pipe), see DESIGN.md section 5"}
    # whole-model roofline, SURVEY.md section 8(d): per block s*B*L*(2Dm + 7Di + 2(R+2N)) + s*P_block on the hot path and
    # B*L*Dm*(4+4+s+s) for the add+RMSNorm either side of it; once per forward the fp32 spectrogram, the patch-embed /
    # position / head parameters and the logits
    Dm, depth, R_ = CFG["embed_dim"], CFG["depth"], (CFG["embed_dim"] + 15) // 16
    p_block = 2 * Di * Dm + Di * 4 + Di + (R_ + 2 * Nst) * Di + Di * R_ + Di + 2 * Di * Nst + Di + Dm * Di + Dm
    rows = B * Lq
    bytes_block = s_act * rows * (2 * Dm + 7 * Di + 2 * (R_ + 2 * Nst)) + s_act * p_block
    bytes_norm = rows * Dm * (4 + 4 + s_act + s_act)
    bytes_once = B * F_ * T_ * 4 + s_act * (256 * Dm + CFG["num_classes"] * Dm) + 4 * Lq * Dm + B * CFG["num_classes"] * 4
    model_bytes = depth * (bytes_block + bytes_norm) + bytes_once
    model_flops = 2.0 * rows * depth * (Dm * 2 * Di + Di * (R_ + 2 * Nst) + R_ * Di + Di * Dm)
    roof_model = {"bound": "hbm", "algorithmic_bytes_per_step": model_bytes, "achieved": model_bytes / (ms_step * 1e-3) / 1e9,
                  "peak": peak, "unit": "GB/s", "frac": model_bytes / (ms_step * 1e-3) / 1e9 / peak,
                  "formula": "SURVEY 8(d): depth*(s*B*L*(2Dm+7Di+2(R+2N)) + s*P_block + B*L*Dm*(8+2s)) + input/embed/head",
                  "gemm_tflops": model_flops / (ms_step * 1e-3) / 1e12,
                  "note": "the path's floors are additive unless kernels overlap: HBM ~4.1 ms + tensor ~4.1 ms + MUFU (scan) ~8.9 ms per 64 clips"}

    # ---- the other GPU configurations of BASELINE.json, after the headline measurement (every rank: the training step
    # holds the path's one collective)
    train = extra = None
    if not args.no_extra:
        train = train_leg(dev, world, dist, rank)
        extra = [small_bibi_leg(dev, world, dist)]

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_sample(clips=2)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": f"AuM-Base Fo-Bi forward (BASELINE configs[1]): 128x1024 mel -> {Lq} tokens, "
                                      f"batch {B}/GPU, depth 24, d_model 768, d_state 16, {n_params} params, random init "
                                      "with perturbed A_log; fp32 residual stream + scan state",
                          "global_batch": world * B, "parallelism": f"dp{world} (replicas, no collective on the forward path)",
                          "l2": "inputs larger than L2: ~27 GB of activations per forward vs 126 MB L2"},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                       "d2h_bytes_per_step": logits_host.numel() * 4,
                       "api": "model(x_host) per step" if args.e2e_per_call else
                              "AudioMamba.infer_stream(host batches): H2D of batch i+1 overlaps the forward of batch i"},
               "gpu_launches": launches, "launch_mode": "eager" if args.no_graph else "cuda-graph replay of the same launches",
               "roofline": roof, "roofline_model": roof_model, "cpu_baseline": cpu, "clocks": clocks,
               "train": train, "extra": extra}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
