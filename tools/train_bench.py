"""BASELINE config 3: AuM-Base Fo-Bi TRAINING step (fwd + bwd + gradient all-reduce + Adam), bf16 autocast-style
activations, VGGSound-shape inputs (128x1024 mel, 309 classes), batch 32 per GPU, one process per GPU.

    python tools/train_bench.py [--steps 5] [--warmup 2] [--batch 32] [--depth 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py

Prints one JSON line (rank 0): clips/s for the whole job, ms/step (max over ranks), the share of the step spent
in the single flat NCCL gradient all-reduce.  Optimiser and loss follow the reference's recipe
(/root/reference/src/traintest.py:32-34: Adam betas=(0.95, 0.999), weight_decay 5e-7; BCEWithLogits)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--graph", type=int, default=0, help="1: replay a CUDA graph of the whole step")
    ap.add_argument("--shadow", type=int, default=1, help="0: no 16-bit shadow weights from the Adam kernel")
    args = ap.parse_args()
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from aum_b200 import dist as D
    from aum_b200.audio_mamba import AudioMamba
    world = D.init_from_env("nccl", dev)
    act = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    torch.manual_seed(3949)
    model = AudioMamba(embed_dim=768, depth=args.depth, num_classes=309, bimamba_type="v1", act_dtype=act).to(dev)
    g = torch.Generator().manual_seed(3949 + rank)
    with torch.no_grad():
        for blk in model.layers:
            blk.mixer.A_log.add_(0.1 * torch.randn(blk.mixer.A_log.shape, generator=g).to(dev))
            blk.mixer.A_b_log.add_(0.1 * torch.randn(blk.mixer.A_b_log.shape, generator=g).to(dev))
    from aum_b200.trainer import TrainStep
    ts = TrainStep(model, lr=1e-5, n_chunks=3, cuda_graph=bool(args.graph), shadow16=bool(args.shadow))
    red = ts.reducer
    x = (0.5 * torch.randn(args.batch, 1024, 128, generator=g)).to(dev)
    y = (torch.rand(args.batch, 309, generator=g) > 0.97).float().to(dev)
    ar_ms = []
    if not args.graph:
        ts.timing = ar_ms

    def step():
        return ts(x, y)

    for _ in range(args.warmup):
        step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ar_ms.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step()
    t1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = D.max_over_ranks(t0.elapsed_time(t1) / args.steps, dev)
    ar = sum(a.elapsed_time(b) for a, b in ar_ms) / len(ar_ms) if ar_ms else None
    if rank == 0:
        print(json.dumps({"metric": "clips/sec AuM-Base training step (fwd+bwd+allreduce+Adam)", "value": world * args.batch / (ms / 1e3),
                          "unit": "clips/s", "n_gpus": world, "ms_per_step": ms, "allreduce_ms": ar, "dtype": args.dtype, "graph": args.graph, "shadow": args.shadow,
                          "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "config": {"workload": f"AuM-Base Fo-Bi depth {args.depth}, 309 classes, batch {args.batch}/GPU, 128x1024 mel"}}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
