"""Live per-kernel times of the config-2 step as bench.py runs it (tq.GraphedStep replays), from torch.profiler's CUDA
activity records (warm caches, back to back; ncu's per-launch times are cold and serialised)."""
import collections
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchquantum_b200 as tq  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = bench.VQEModel(tq).to(dev)
params = [p for p in model.parameters() if p.requires_grad]
x = (torch.rand(bench.BSZ, bench.N_WIRES) * np.pi).to(dev)
for _ in range(3):
    for p in params:
        p.grad = None
    model(x).backward()
step = tq.GraphedStep(lambda v: model(v), [x], params, warmup=3)
for _ in range(5):
    step(x)
torch.cuda.synchronize()
K = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(K):
        step(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name[:90]][0] += 1
        agg[ev.name[:90]][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"kernel time per step {tot / K:.1f} us over {sum(v[0] for v in agg.values()) / K:.0f} launches")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / K:8.1f} us/step {100 * t / tot:5.1f}%  x{c / K:5.1f}  {name}")
