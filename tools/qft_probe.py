"""QFT (H + cu1 ladders + swaps, layer/layers/qft_layer.py:11-115) on one GPU: time and analytic check |amp| = 2^(-n/2).
Usage: python tools/qft_probe.py [--n 24 28]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torchquantum_b200 as tq  # noqa: E402
from dist_sharded import apply, qft_ops  # noqa: E402
from torchquantum_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="*", default=[24, 28])
a = ap.parse_args()
dev = torch.device("cuda", 0)
for n in a.n:
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)
    ops = qft_ops(n)

    def run():
        with torch.no_grad():
            q.reset_states(1)
            apply(q, ops, dev)
            return q.get_states_1d()

    s = run()
    torch.cuda.synchronize()
    l0 = engine.STATS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        s = run()
    e1.record()
    torch.cuda.synchronize()
    amp = s[0, :2].abs().cpu().numpy() * 2 ** (n / 2)
    print(f"QFT n={n}: {len(ops)} gates, {e0.elapsed_time(e1) / 3:.2f} ms incl. host queueing, "
          f"{(engine.STATS['launches'] - l0) / 3:.0f} launches, |amp| * 2^(n/2) = {amp}")
