"""Per-block timing of the planned config-4 family circuit: which block placements (kpos, transposed write-back) are slow.
python tools/block_times.py [--n 28] [--layers 16] [--extra]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torchquantum_b200 import _codes as C, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--layers", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    key, thetas = bench._cfg4_family(a.n, a.layers)
    plan = engine.get_plan(a.n, tuple(key), False)
    state = torch.zeros(1, 2 ** a.n, dtype=torch.complex64, device=dev)
    state[0, 0] = 1
    other = torch.empty_like(state)
    for stage in plan.stages:
        if stage.kind != "blocks":
            print("stage", stage.kind)
            continue
        psh = torch.tensor([thetas[r] for r in stage.sh_recs] or [0.0], dtype=torch.float32, device=dev)
        acc = np.zeros(stage.n_blocks)
        for rep in range(4):
            t = []
            out = engine.run_blocks(state, stage, a.n, 0, psh, True, timing=t, scratch=other)
            if out is not state:
                state, other = out, state
            torch.cuda.synchronize()
            if rep:
                acc += np.array([s.elapsed_time(e) for s, e in t])
        acc /= 3
        for b in range(stage.n_blocks):
            hdr = stage.blocks[b]
            kpos = [int(hdr[C.KB_KPOS + i]) for i in range(6)]
            print(f"block {b:3d} kpos {kpos} T={int(hdr[C.KB_FLAGS]) & C.KBF_TRANSPOSE} R={(int(hdr[C.KB_FLAGS]) & C.KBF_ROTATE) >> 1} gates={int(hdr[C.KB_NOPS]):3d}  {acc[b]:.3f} ms")
        print("mean", acc.mean(), "blocks", stage.n_blocks, "total ms", acc.sum())


if __name__ == "__main__":
    main()
