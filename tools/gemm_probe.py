"""GPU probe of tqb_gemm_tf32x3 through engine.complex_gemm: accuracy vs a float64 product and timing at the config-2 shapes.
python tools/gemm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchquantum_b200 import engine  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
ok = True
for M, K, N in ((4096, 1024, 1024), (1024, 4096, 1024), (300, 128, 128), (128, 256, 512), (4096, 128, 128)):
    a = torch.randn(M, K, dtype=torch.complex64, device=dev) / K ** 0.5
    b = torch.randn(K, N, dtype=torch.complex64, device=dev)
    got = engine.complex_gemm(a, b)
    want = a.to(torch.complex128) @ b.to(torch.complex128)
    rel = ((got - want).norm() / want.norm()).item()
    mx = ((got - want).abs().max() / want.abs().max()).item()
    ref32 = ((a @ b - want).norm() / want.norm()).item()
    for _ in range(3):
        engine.complex_gemm(a, b)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        engine.complex_gemm(a, b)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
    # kernel alone
    ar = torch.view_as_real(a).reshape(M, 2 * K)
    Mp = -(-M // 128) * 128
    if Mp != M:
        ar = torch.nn.functional.pad(ar, (0, 0, 0, Mp - M))
    bt = b.transpose(0, 1).conj()
    btr = torch.view_as_real(torch.stack([bt, 1j * bt], dim=1).contiguous()).reshape(2 * N, 2 * K)
    ah, al = engine._tf32_planes(ar)
    bh, bl = engine._tf32_planes(btr)
    c = torch.empty(Mp, 2 * N, dtype=torch.float32, device=dev)
    from torchquantum_b200 import _native
    for _ in range(3):
        _native.call("tqb_gemm_tf32x3", ah, al, bh, bl, c)
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(10):
        _native.call("tqb_gemm_tf32x3", ah, al, bh, bl, c)
        ev[i + 1].record()
    torch.cuda.synchronize()
    kms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
    flops = 3 * 2.0 * Mp * (2 * N) * (2 * K)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        a @ b
    t1.record(); torch.cuda.synchronize()
    print(f"M={M} K={K} N={N}: rel_l2 {rel:.2e} max_rel {mx:.2e} (torch c64 matmul: {ref32:.2e}) | complex_gemm {ms:.3f} ms, kernel "
          f"{kms:.3f} ms = {flops / kms / 1e9:.0f} TFLOP/s tf32 (incl. the 3x split), torch cgemm {t0.elapsed_time(t1) / 10:.3f} ms", flush=True)
    ok &= rel < 2e-6
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
