"""Sustained vs burst: one rotating block pass (n = 28) launched back to back for several seconds while nvidia-smi samples
the SM clock and the power draw; the same for a plain device copy.  Shows what the power cap does to each."""
import ctypes
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.umma_probe import a_planes  # noqa: E402

n = 28
dev = "cuda:0"
lib = ctypes.CDLL(os.path.join(ROOT, "torchquantum_b200", "lib", "libtqb200.so"))
fn = lib.tqb_apply_blocks
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p]
g = torch.Generator().manual_seed(1)
M = torch.randn(64, 64, dtype=torch.float64, generator=g) + 1j * torch.randn(64, 64, dtype=torch.float64, generator=g)
U, _ = torch.linalg.qr(M)
planes = a_planes(U).to(dev)
a = torch.randn(1, 2 ** n, 2, device=dev)
a /= a.norm()
b = torch.empty_like(a)
stream = torch.cuda.current_stream().cuda_stream
hdr = (ctypes.c_int32 * 12)(*([0, 0, 2, 0] + list(range(20, 26)) + [0, 0]))

samples = []
stop = False


def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append((time.time(), out))
        time.sleep(0.05)


def run(name, launch, seconds=3.0, chunk=100):
    global stop, samples
    samples, stop = [], False
    th = threading.Thread(target=sampler)
    th.start()
    torch.cuda.synchronize()
    t0 = time.time()
    times = []
    while time.time() - t0 < seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(chunk):
            launch()
        e1.record()
        torch.cuda.synchronize()
        times.append((time.time() - t0, e0.elapsed_time(e1) / chunk))
    stop = True
    th.join()
    print(f"== {name}")
    for t, ms in times[:3] + times[len(times) // 2:len(times) // 2 + 2] + times[-3:]:
        print(f"   t={t:5.2f}s  {ms:.4f} ms per launch  ({16 * 2 ** n / ms / 1e6:.0f} GB/s)")
    for t, s in samples[::max(len(samples) // 12, 1)]:
        print(f"   t={t - t0:5.2f}s  sm MHz, mem MHz, W, C, sw_power_cap: {s}")


state = [a, b]


def block():
    fn(state[0].data_ptr(), state[1].data_ptr(), None, n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, stream)
    state.reverse()


run("rotating block pass [20..25]", block)
time.sleep(2)
run("device copy 2 GiB -> 2 GiB", lambda: b.copy_(a))
time.sleep(2)
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
run("cuBLAS bf16 8192^3", lambda: torch.matmul(x, x), chunk=20)
