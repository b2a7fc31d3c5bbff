"""QFT on one GPU through the planner (blocks + quadratic-phase passes) and the quadratic-phase kernel alone: timings at n = 28."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torchquantum_b200 as tq  # noqa: E402
from torchquantum_b200 import _native, engine, _codes as C  # noqa: E402
from dist_sharded import qft_ops, apply  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
# the kernel alone
st = torch.randn(1, 2 ** n, dtype=torch.complex64, device=dev)
coef = torch.randint(-2 ** 31, 2 ** 31 - 1, (C.QP_WORDS,), dtype=torch.int64).to(torch.int32).to(dev)
for _ in range(3):
    _native.call("tqb_quadphase", st, n, 1, coef, n, 0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for k in range(10):
    _native.call("tqb_quadphase", st, n, 1, coef, n, 0)
    ev[k + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(10))[5]
print(f"tqb_quadphase n={n}: {ms:.3f} ms = {16 * 2 ** n / ms / 1e6:.0f} GB/s")
del st
ops = qft_ops(n)
q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)
x = 0b1011001110001101011010011101 & ((1 << n) - 1)


def run():
    with torch.no_grad():
        q.reset_states(1)
        s = q.get_states_1d()
        s[0, 0] = 0
        s[0, x] = 1
        apply(q, ops, dev)
        return q.get_states_1d()


s = run()
torch.cuda.synchronize()
plan = next(reversed(engine._PLAN_CACHE.values()))
for stg in plan.stages:
    print("stage", stg.kind, "blocks", getattr(stg, "n_blocks", None), "quads", len(getattr(stg, "quads", [])),
          "passes", len(getattr(stg, "passes", [])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    s = run()
e1.record()
torch.cuda.synchronize()
k = torch.arange(2 ** n, dtype=torch.float64, device=dev)
ref = torch.exp(2j * np.pi * ((x * k) % 2 ** n) / 2 ** n) / np.sqrt(2 ** n)
rel = ((s[0].to(torch.complex128) - ref).norm() / ref.norm()).item()
print(f"QFT n={n}: {len(ops)} gates, {e0.elapsed_time(e1) / 3:.2f} ms (incl. host queueing), rel err vs analytic {rel:.2e}")
