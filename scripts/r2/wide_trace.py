"""clock64 timeline of k_wide_backward (CTA 0, tiles 2..5 of the CTA); needs S3D_NVCC_EXTRA=-DS3D_WTRACE.  Development tool."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from seal3d_b200 import _lib
dev = torch.device("cuda", 0)
B = 1356004
cin, hid, cout, nl = 160, 128, 16, 2
x = torch.randn(B, cin, device=dev).half()
w = (torch.randn(hid * cin + hid * hid * (nl - 1) + cout * hid, device=dev) * 0.05).half()
fb = torch.empty(nl, B, hid, device=dev, dtype=torch.float16)
out = torch.empty(B, cout, device=dev, dtype=torch.float16)
_lib.call("s3d_ffmlp_forward", x, w, B, cin, cout, hid, nl, 0, 6, fb, out)
g = torch.randn(B, cout, device=dev).half()
bb = torch.empty_like(fb); gi = torch.empty_like(x); gw = torch.empty_like(w)
for _ in range(3):
    _lib.call("s3d_ffmlp_backward", g, x, w, fb, B, cin, cout, hid, nl, 0, 6, 1, bb, gi, gw)
torch.cuda.synchronize()
lib = _lib.lib()
n = 4 * 2 * 4 * 8
buf = (ctypes.c_longlong * n)()
lib.s3d_debug_wtrace.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.s3d_debug_wtrace(buf, n))
t = np.array(buf, dtype=np.int64).reshape(4, 2, 4, 8)
ev_names = ["pre-sync", "post-sync", "mma-issued", "loads-issued", "mma-done", "epilogue-done"]
for it in range(4):
    t0 = t[it, 0, 3, 0]
    print("== tile %d of CTA 0: start 0, operand in TMEM at %d" % (it + 2, t[it, 0, 3, 1] - t0))
    for s in range(3):
        for who in range(2):
            print("   step %d thread %3d: " % (s, who * 128) + "  ".join("%s %d" % (ev_names[k], t[it, who, s, k] - t0) for k in range(6) if t[it, who, s, k] > 0))
    if it + 1 < 4:
        print("   next tile starts at %d" % (t[it + 1, 0, 3, 0] - t0))
