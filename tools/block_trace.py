"""Per-tile pipeline trace of tqb_block6_kernel (CTA 0, first 64 tiles): clock64 stamps of every hand-off.
python tools/block_trace.py [--n 28] [--kpos 6 7 8 9 10 11]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ["load_issue", "full_seen", "lo_free", "split_done", "mma_ready", "mma_issue", "acc_full", "epi_done", "out_ready", "stored"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--kpos", type=int, nargs=6, default=[6, 7, 8, 9, 10, 11])
    ap.add_argument("--rotate", type=int, default=0, help="1: rotating write-back into a second buffer")
    ap.add_argument("--warm", type=int, default=3, help="untraced passes first (1500: the board reaches its power cap)")
    a = ap.parse_args()
    lib = ctypes.CDLL(os.path.join(ROOT, "tools", "_ab", "libtqb200_debug.so"))   # tools/build_debug_lib.sh
    lib.tqb_last_error.restype = ctypes.c_char_p
    fn = lib.tqb_apply_blocks
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.tqb_debug_block_trace.argtypes = [ctypes.c_void_p]
    dev = "cuda:0"
    from tools.umma_probe import a_planes
    g = torch.Generator().manual_seed(1)
    M = torch.randn(64, 64, dtype=torch.float64, generator=g) + 1j * torch.randn(64, 64, dtype=torch.float64, generator=g)
    U, _ = torch.linalg.qr(M)
    planes = a_planes(U).to(dev)
    st = torch.randn(1, 2 ** a.n, 2, device=dev, dtype=torch.float32)
    st /= st.norm()
    trace = torch.zeros(64 * 10, dtype=torch.int64, device=dev)
    hdr = (ctypes.c_int32 * 12)(*([0, 0, 2 * a.rotate, 0] + list(a.kpos) + [0, 0]))
    bufs = [st, torch.empty_like(st) if a.rotate else st]

    def run():
        rc = fn(bufs[0].data_ptr(), bufs[1].data_ptr(), None, a.n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1,
                torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.tqb_last_error().decode()
        bufs.reverse()

    for _ in range(a.warm):
        run()
    torch.cuda.synchronize()
    lib.tqb_debug_block_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.tqb_debug_block_trace(None)
    t = trace.cpu().numpy().reshape(64, 10).astype(np.float64)
    t0 = t[0, 0]
    t = (t - t0) / 1000.0   # kilo-cycles
    print("kilocycles since the first load issue; columns:", " ".join(NAMES))
    for i in range(0, 40):
        print(f"{i:3d} " + " ".join(f"{v:8.2f}" for v in t[i]))
    steady = slice(16, 60)
    per = np.diff(t[steady, 9]).mean()
    print(f"\nsteady-state period {per:.3f} kcycles/tile")
    for a_, b_, what in ((0, 1, "load issue -> landed"), (1, 2, "wait for lo plane"), (2, 3, "split"), (3, 4, "ready -> MMA warp sees it"),
                         (4, 5, "MMA waits for a free accumulator"), (5, 6, "MMA issue -> accumulator complete"),
                         (6, 7, "epilogue"), (7, 8, "epilogue done -> store warp"), (8, 9, "store issue -> smem read out")):
        d = t[steady, b_] - t[steady, a_]
        print(f"  {what:42s} mean {d.mean():7.3f}  min {d.min():7.3f}  max {d.max():7.3f}")


if __name__ == "__main__":
    main()
