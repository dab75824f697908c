"""Which torch ops still launch kernels in a training step (2-block AuM-Base, bf16): torch.profiler op table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-mamba-aum_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from aum_b200.audio_mamba import AudioMamba
from aum_b200.trainer import TrainStep
dev = torch.device("cuda")
torch.manual_seed(0)
model = AudioMamba(embed_dim=768, depth=2, num_classes=309, bimamba_type="v1", act_dtype=torch.bfloat16).to(dev)
ts = TrainStep(model, n_chunks=1)
x = 0.5 * torch.randn(32, 1024, 128, device=dev)
y = (torch.rand(32, 309, device=dev) > 0.97).float()
for _ in range(2):
    ts(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ts(x, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
