"""Quick device-side timing probe (not the bench): config-2-like on-chip circuit and streamed
large-n passes.  Usage: python tools/perf_probe.py [--big 26]"""
import argparse
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchquantum_b200 as tq
import torchquantum_b200.functional as tqf
from torchquantum_b200 import engine
from torchquantum_b200.measurement import expval_joint_analytical_multi


def ev_time(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def cfg2(bsz=4096, n=10, reps=20):
    dev = "cuda:0"
    x = torch.rand(bsz, n, device=dev) * np.pi
    thetas = [torch.nn.Parameter(torch.rand(1, 1, device=dev) * 6 - 3) for _ in range(2 * n * (reps + 1))]
    obs = ["I" * i + "ZZ" + "I" * (n - 2 - i) for i in range(n - 1)] + ["I" * i + "X" + "I" * (n - 1 - i) for i in range(n)]

    def step(backward=True):
        q = tq.QuantumDevice(n_wires=n, bsz=bsz, device=dev)
        for w in range(n):
            tqf.ry(q, wires=w, params=x[:, w])
        k = 0
        for r in range(reps + 1):
            for w in range(n):
                tqf.ry(q, wires=w, params=thetas[k]); k += 1
            for w in range(n):
                tqf.rz(q, wires=w, params=thetas[k]); k += 1
            if r < reps:
                for w in range(n - 1):
                    tqf.cnot(q, wires=[w, w + 1])
        ev = expval_joint_analytical_multi(q, obs)
        loss = ev.sum(-1).mean()
        if backward:
            loss.backward()
        return loss

    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"cfg2 first step (plan build) {1e3*(t1-t0):.1f} ms")
    tf = ev_time(lambda: step(False))
    tfb = ev_time(step)
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    print(f"cfg2 n={n} bsz={bsz}: fwd {tf:.3f} ms, fwd+bwd {tfb:.3f} ms (events), wall {1e3*wall:.3f} ms -> {bsz/wall:.0f} circuits/s")


def big(n, layers=4):
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    ops = []
    for layer in range(layers):
        for w in range(n):
            ops.append((["rx", "ry", "rz"][rng.integers(3)], [w], torch.tensor([[rng.uniform(-3, 3)]], dtype=torch.float32, device=dev)))
        for w in range(layer % 2, n - 1, 2):
            ops.append(("cnot", [w, w + 1], None))
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)

    def run():
        q.reset_states(1)
        for name, wires, p in ops:
            getattr(tqf, name)(q, wires=wires, params=p)
        return q.get_states_1d()

    with torch.no_grad():
        run()
        plan = list(engine._PLAN_CACHE.values())[-1]
        t = ev_time(run, warm=2, iters=5)
    npass = plan.n_passes
    bytes_pass = 16 * 2 ** n
    print(f"n={n}: {len(ops)} gates in {npass} passes, {t:.3f} ms total, {t/npass:.3f} ms/pass, "
          f"{npass*bytes_pass/t/1e6:.0f} GB/s algorithmic (incl. reset+launch overheads)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, nargs="*", default=[24, 26, 28])
    ap.add_argument("--skip-cfg2", action="store_true")
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0))
    if not a.skip_cfg2:
        cfg2()
    for n in a.big:
        big(n)
