"""Full-depth parity of the config-4 family on the GPU against the fp64 twin of the oracle (oracle/tq_oracle64.py, run
with torch's CUDA ops): brickwork rx/ry/rz + CNOT on |0...0>, forward-only (tensor-core blocks / register-tile passes)
and the grad-mode program; rel-L2 and |norm - 1| (bars: 1e-5).  Also times config 4 itself.

  python tools/depth_parity.py [--n 20 22 24] [--depth 100] [--full 28] [--out gpurun_out/depth_parity.json]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="*", default=[20, 22, 24])
    ap.add_argument("--depth", type=int, default=100)
    ap.add_argument("--full", type=int, default=0, help="also run this n forward-only (blocks on and off)")
    ap.add_argument("--out", default="gpurun_out/depth_parity.json")
    args = ap.parse_args()
    dev = "cuda:0"
    rows = []
    for n in args.n:
        for mode, blocks in (("fwd", True), ("fwd", False), ("grad", None)):
            t = time.time()
            rel, nerr, kinds = cases.depth_parity(dev, n, args.depth, mode, blocks=blocks)
            rows.append({"n": n, "depth": args.depth, "mode": mode, "blocks": "blocks" in kinds, "rel_l2": rel,
                         "norm_err": nerr, "wall_s": time.time() - t})
            print(rows[-1], flush=True)
    if args.full:
        for blocks in (True, False):
            t = time.time()
            rel, nerr, kinds = cases.depth_parity(dev, args.full, args.depth, "fwd", blocks=blocks)
            rows.append({"n": args.full, "depth": args.depth, "mode": "fwd", "blocks": "blocks" in kinds, "rel_l2": rel,
                         "norm_err": nerr, "wall_s": time.time() - t})
            print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)
    bad = [r for r in rows if r["rel_l2"] > 1e-5 or r["norm_err"] > 1e-5]
    print("FAILED" if bad else "all within 1e-5", bad)


if __name__ == "__main__":
    main()
