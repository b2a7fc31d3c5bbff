"""K4 probe: a dense k-qubit unitary on an n-qubit state through (a) the tensor-core kernel (mma.sync, 3xTF32),
(b) the CUDA-core dense kernel, and (c) the same unitary's worth of brickwork gates through the fused interpreter.
Usage: python tools/dense_probe.py [--n 26]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchquantum_b200 import _native  # noqa: E402


def ev(fn, iters=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=26)
    a = ap.parse_args()
    dev, n = "cuda:0", a.n
    st = torch.randn(1, 2 ** n, dtype=torch.complex64, device=dev)
    st /= st.norm()
    out = torch.empty_like(st)
    for k in (4, 5, 6, 7):
        u, _ = torch.linalg.qr(torch.randn(2 ** k, 2 ** k, dtype=torch.complex64))
        u = u.to(dev).contiguous()
        for label, bits in (("low bits", list(range(k))[::-1]), ("high bits", list(range(n - k, n))[::-1])):
            host = [int(b) for b in bits]
            bp = torch.tensor(bits, dtype=torch.int32, device=dev)
            t_tc = ev(lambda: _native.call("tqb_apply_dense_tc", st, out, n, 1, u, k, host, 0))
            ref = torch.empty_like(st)
            t_cc = ev(lambda: _native.call("tqb_apply_dense", st, ref, n, 1, u, 0, k, bp, 0), iters=2)
            err = float((out - ref).abs().max() / ref.abs().max())
            flops = 8.0 * 2 ** k * 3 * 2 ** n
            print(f"n={n} k={k} {label:9s}: tensor-core {t_tc:7.3f} ms ({16 * 2 ** n / t_tc / 1e6:6.0f} GB/s alg, "
                  f"{flops / t_tc / 1e9:6.1f} TFLOP/s incl. 3x split), CUDA-core dense {t_cc:8.3f} ms, max rel diff {err:.1e}")


if __name__ == "__main__":
    main()
