"""Kernel-only timing of the config-2 circuit (10-qubit EfficientSU2 reps=20 behind a per-sample ry encoder,
bsz=4096): CUDA events around tqb_pass_forward / tqb_pass_backward of the step's own plan, no Python in between.
Usage: python tools/cfg2_kernels.py [--bsz 4096] [--n 10] [--reps 20]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchquantum_b200 import _native, engine  # noqa: E402
from torchquantum_b200.engine import _build_mats, _desc, _dev  # noqa: E402


def cfg2_key(n, reps, grad):
    key = [("ry", (w,), False, "ps", False, 0) for w in range(n)]
    for r in range(reps + 1):
        for name in ("ry", "rz"):
            key += [(name, (w,), False, "sh", grad, 0) for w in range(n)]
        if r < reps:
            key += [("cnot", (w, w + 1), False, None, False, 0) for w in range(n - 1)]
    return tuple(key)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bsz", type=int, default=4096)
    ap.add_argument("--n", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = "cuda:0"
    n, bsz = a.n, a.bsz
    out = {}
    for grad in (True, False):
        key = cfg2_key(n, a.reps, grad)
        plan = engine.get_plan(n, key, grad)
        stage = plan.stages[0]
        g = torch.Generator().manual_seed(0)
        psh = (torch.rand(stage.P_sh, generator=g) * 6 - 3).to(dev)
        pps = (torch.rand(bsz, max(stage.P_ps, 1), generator=g) * np.pi).to(dev)
        state = torch.zeros(bsz, 2 ** n, dtype=torch.complex64, device=dev)
        state[:, 0] = 1
        dst = torch.empty_like(state)
        lam = torch.randn(bsz, 2 ** n, dtype=torch.complex64, device=dev)
        psi2, lam2 = torch.empty_like(state), torch.empty_like(state)
        gr = torch.zeros(bsz, stage.G, dtype=torch.float32, device=dev)
        msh, mps = _build_mats(stage, bsz, psh, pps, [], dev)
        assert len(stage.passes) == 1
        pp = stage.passes[0]
        prog = _dev(pp, "prog", dev)

        def fwd():
            _native.call("tqb_pass_forward", state, dst, n, bsz, prog, _desc(pp), msh, mps,
                         stage.M_ps, 0)

        def bwd():
            _native.call("tqb_pass_backward", dst, lam, psi2, lam2, n, bsz, prog, _desc(pp),
                         msh, mps, stage.M_ps, gr, stage.G, 0)

        for name, fn in (("fwd", fwd),) + ((("bwd", bwd),) if grad else ()):
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(a.iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / a.iters
            gau = len(key) * bsz * 2.0 ** n / (ms * 1e-3)
            label = f"{name} ({'grad' if grad else 'forward-only'} program, {pp.n_seg} segments, {pp.n_groups} groups)"
            print(f"cfg2 n={n} bsz={bsz} {label}: {ms:.3f} ms, {gau / 1e12:.2f}e12 gate-amp updates/s")
            out[f"{name}_{'grad' if grad else 'fast'}_ms"] = ms
        if grad:
            err = abs(float((psi2.abs() ** 2).sum(1).mean()) - 1.0)
            print(f"  un-applied state norm error {err:.2e}")
    return out


if __name__ == "__main__":
    main()
