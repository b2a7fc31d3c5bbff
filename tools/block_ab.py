"""A/B timing of one block pass (n = 28): a previous build of the library (tools/_ab/libtqb200_old.so, 9-argument
tqb_apply_blocks) against the current one, in place / out of place / rotating, plus the box's copy bandwidth."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.umma_probe import a_planes  # noqa: E402

n = 28
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
M = torch.randn(64, 64, dtype=torch.float64, generator=g) + 1j * torch.randn(64, 64, dtype=torch.float64, generator=g)
U, _ = torch.linalg.qr(M)
planes = a_planes(U).to(dev)
st = torch.randn(1, 2 ** n, 2, device=dev, dtype=torch.float32)
st /= st.norm()
other = torch.empty_like(st)
stream = torch.cuda.current_stream().cuda_stream


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        f()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    return ms[len(ms) // 2], ms[0]


med, mn = timeit(lambda: other.copy_(st))
print(f"copy_ 2 GiB -> 2 GiB: {med:.3f} ms median = {2 * st.numel() * 4 / med / 1e6:.0f} GB/s", flush=True)

places = {"low": list(range(6, 12)), "mid": list(range(14, 20)), "19_24": list(range(19, 25)), "top": list(range(22, 28))}
old_tag = next((a for a in sys.argv[1:] if a.startswith("o")), "old" if len(sys.argv) < 2 else None)
old_path = os.path.join(ROOT, "tools", "_ab", f"libtqb200_{old_tag}.so")
if old_tag and os.path.exists(old_path):
    old = ctypes.CDLL(old_path)
    fo = old.tqb_apply_blocks
    fo.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    for name, kpos in places.items():
        hdr = (ctypes.c_int32 * 12)(*([0, 0, 0, 0] + kpos + [0, 0]))
        med, mn = timeit(lambda: fo(st.data_ptr(), st.data_ptr(), n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, stream))
        print(f"{old_tag:4s} in place      {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)
        med, mn = timeit(lambda: fo(st.data_ptr(), other.data_ptr(), n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, stream))
        print(f"old  out of place  {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)

variants = [("new", os.path.join(ROOT, "torchquantum_b200", "lib", "libtqb200.so"))]
for v in sorted(os.listdir(os.path.join(ROOT, "tools", "_ab"))):
    if v.startswith("libtqb200_v"):
        variants.append((v[len("libtqb200_"):-3], os.path.join(ROOT, "tools", "_ab", v)))
if len(sys.argv) > 1:
    variants = [v for v in variants if v[0] in sys.argv[1:]]
assert len(variants) == 1 or True   # NOTE: one variant per process (static locals of inline functions are process-global)
for tag, path in variants:
  new = ctypes.CDLL(path)
  new.tqb_last_error.restype = ctypes.c_char_p
  fn = new.tqb_apply_blocks
  fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                 ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]


  def call(src, dst, hdr):
      rc = fn(src.data_ptr(), dst.data_ptr(), None, n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, stream)
      if rc:
          raise RuntimeError(new.tqb_last_error().decode())


  for name, kpos in places.items():
      hdr = (ctypes.c_int32 * 12)(*([0, 0, 0, 0] + kpos + [0, 0]))
      med, mn = timeit(lambda: call(st, st, hdr))
      print(f"{tag:4s} in place      {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)
      med, mn = timeit(lambda: call(st, other, hdr))
      print(f"{tag:4s} out of place  {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)
      hdr_r = (ctypes.c_int32 * 12)(*([0, 0, 2, 0] + kpos + [0, 0]))
      med, mn = timeit(lambda: call(st, other, hdr_r))
      print(f"{tag:4s} rotating      {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)
      bufs = [st, other]

      def pingpong():
          call(bufs[0], bufs[1], hdr_r)
          bufs.reverse()
      med, mn = timeit(pingpong)
      print(f"{tag:4s} rotating a<->b {name:6s} {med:.3f} ms (min {mn:.3f})", flush=True)
