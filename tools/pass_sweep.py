"""Per-pass timing sweep of the fused gate-pass kernel: one pass over a 2^n state with k gates on
the tile's qubits, k = 0 (pure copy through shared memory) ... 96.  Prints achieved algorithmic
GB/s (16 B per amplitude per pass) so the HBM-bound and the issue-bound regimes can be told apart.
Usage: python tools/pass_sweep.py [--n 28] [--bwd]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchquantum_b200 import _native, engine  # noqa: E402
from torchquantum_b200.engine import _build_mats, _desc, _dev  # noqa: E402


def one_pass_key(n, k, rng, kinds=("rx", "ry", "rz"), cnot_every=3, hi_bits=7):
    """k gates whose targets all lie in one tile: the 5 low bits plus `hi_bits` high ones."""
    bits = list(range(5)) + list(range(n - hi_bits, n))
    key, thetas = [], []
    for i in range(k):
        if cnot_every and i % cnot_every == cnot_every - 1:
            a, b = rng.choice(len(bits), 2, replace=False)
            key.append(("cnot", (n - 1 - bits[a], n - 1 - bits[b]), False, None, False, 0))
        else:
            b = bits[rng.integers(len(bits))]
            key.append((kinds[rng.integers(len(kinds))], (n - 1 - b,), False, "sh", False, 0))
            thetas.append(rng.uniform(-np.pi, np.pi))
    return key, thetas


def time_passes(n, key, thetas, dev, bwd=False, reps=5):
    plan = engine.get_plan(n, tuple(key), bwd)
    if not plan.stages:
        return None
    stage = plan.stages[0]
    psh = torch.tensor(thetas or [0.0], dtype=torch.float32, device=dev)
    pps = torch.zeros(1, 1, dtype=torch.float32, device=dev)
    state = torch.zeros(1, 2 ** n, dtype=torch.complex64, device=dev)
    state[0, 0] = 1
    lam = torch.zeros_like(state) if bwd else None
    gr = torch.zeros(1, stage.G, dtype=torch.float32, device=dev)
    msh, mps = _build_mats(stage, 1, psh, pps, [], dev)
    evs = []
    for rep in range(reps + 1):
        for pp in stage.passes:
            prog = _dev(pp, "prog", dev)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if bwd:
                _native.call("tqb_pass_backward", state, lam, state, lam, n, 1, prog,
                             _desc(pp), msh, mps, stage.M_ps, gr, stage.G, 0)
            else:
                _native.call("tqb_pass_forward", state, state, n, 1, prog, _desc(pp), msh,
                             mps, stage.M_ps, 0)
            e.record()
            if rep > 0:
                evs.append((s, e))
    torch.cuda.synchronize()
    ms = np.array([s.elapsed_time(e) for s, e in evs])
    return ms, len(stage.passes), sum(p.n_groups for p in stage.passes)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--ks", type=int, nargs="*", default=[1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96])
    ap.add_argument("--mix", default="both", choices=["both", "rot", "gen"])
    a = ap.parse_args()
    dev = "cuda:0"
    print(torch.cuda.get_device_name(0), "n =", a.n, "bwd" if a.bwd else "fwd")
    bytes_pass = (32.0 if a.bwd else 16.0) * 2 ** a.n
    mixes = [(("rx", "ry", "rz"), "rx/ry/rz+cnot"), (("u3like",), "general 2x2")]
    mixes = [m for m in mixes if a.mix == "both" or (a.mix == "rot") == (m[0][0] == "rx")]
    for kinds, label in mixes:
        for k in a.ks:
            rng = np.random.default_rng(k)
            if kinds[0] == "u3like":
                key, thetas = one_pass_key(a.n, k, rng, kinds=("sx",), cnot_every=0)
                key = [(nm, w, inv, None, g, e) for nm, w, inv, pm, g, e in key]
                thetas = []
            else:
                key, thetas = one_pass_key(a.n, k, rng, kinds=kinds)
            r = time_passes(a.n, key, thetas, dev, bwd=a.bwd)
            if r is None:
                continue
            ms, npass, ngroups = r
            per_pass = ms.mean()
            print(f"{label:>14} k={k:3d}: {npass} pass(es), {ngroups} groups, {per_pass:.3f} ms/pass, "
                  f"{bytes_pass / per_pass / 1e6:7.0f} GB/s algorithmic, {k * 2 ** a.n / (ms.sum() / (len(ms) / npass)) / 1e6:8.0f} M gate-amp/ms")
