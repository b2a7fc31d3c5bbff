"""Timing of one block pass for several block-bit placements at n = 28 (which placements are slow: distinct 2 MiB pages
touched per tile = 2^(number of block bits >= 18))."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.umma_probe import a_planes  # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "torchquantum_b200", "lib", "libtqb200.so"))
fn = lib.tqb_apply_blocks
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
               ctypes.c_int, ctypes.c_void_p]
n = 28
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
M = torch.randn(64, 64, dtype=torch.float64, generator=g) + 1j * torch.randn(64, 64, dtype=torch.float64, generator=g)
U, _ = torch.linalg.qr(M)
planes = a_planes(U).to(dev)
st = torch.randn(1, 2 ** n, 2, device=dev, dtype=torch.float32)
st /= st.norm()
cases = [list(range(b, b + 6)) for b in (6, 12, 14, 15, 16, 17, 18, 19, 20, 22)] + \
        [[6, 7, 20, 21, 22, 23], [6, 19, 20, 21, 22, 23], [6, 7, 8, 25, 26, 27], [12, 13, 24, 25, 26, 27], [14, 15, 16, 25, 26, 27],
         [16, 17, 18, 25, 26, 27], [10, 23, 24, 25, 26, 27]]
for kpos in cases:
    for tr in (0, 1):
        hdr = (ctypes.c_int32 * 12)(*([0, 0, tr, 0] + kpos + [0, 0]))
        for _ in range(3):
            fn(st.data_ptr(), st.data_ptr(), None, n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        ev[0].record()
        for k in range(8):
            fn(st.data_ptr(), st.data_ptr(), None, n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, torch.cuda.current_stream().cuda_stream)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(8))[4]
        pages = 2 ** sum(1 for b in kpos if b >= 18)
        print(f"kpos {kpos} T={tr} pages/tile {pages:3d}  {ms:.3f} ms  {16 * 2 ** n / ms / 1e6:7.0f} GB/s", flush=True)
