import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.umma_probe import a_planes
which = sys.argv[1]
n = 28
planes = torch.zeros(2, 128, 64, device="cuda")
st = torch.zeros(1, 2 ** n, 2, device="cuda")
print("state ptr %x" % st.data_ptr())
kpos = list(range(14, 20))
hdr = (ctypes.c_int32 * 12)(*([0, 0, 0, 0] + kpos + [0, 0]))
if which == "old":
    lib = ctypes.CDLL(os.path.join(ROOT, "tools", "_ab", "libtqb200_old.so"))
    f = lib.tqb_apply_blocks
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    f(st.data_ptr(), st.data_ptr(), n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, None)
else:
    lib = ctypes.CDLL(os.path.join(ROOT, "torchquantum_b200", "lib", "libtqb200.so"))
    f = lib.tqb_apply_blocks
    f.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    f(st.data_ptr(), st.data_ptr(), None, n, 1, planes.data_ptr(), ctypes.addressof(hdr), 1, 1, None)
torch.cuda.synchronize()
