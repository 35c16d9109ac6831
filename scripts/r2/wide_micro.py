"""micro-timings of the wide FFMLP / linear kernels at the TensoRF head's shapes (development tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from seal3d_b200 import _lib
dev = torch.device("cuda", 0)
B = 1356004
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
SHAPES = ((160, 128, 16, 2), (32, 64, 16, 6), (32, 256, 16, 2))
if os.environ.get('SHAPE'):
    SHAPES = (SHAPES[int(os.environ['SHAPE'])],)
for (cin, hid, cout, nl) in SHAPES:
    x = torch.randn(B, cin, device=dev).half()
    w = (torch.randn(hid * cin + hid * hid * (nl - 1) + cout * hid, device=dev) * 0.05).half()
    fb = torch.empty(nl, B, hid, device=dev, dtype=torch.float16)
    out = torch.empty(B, cout, device=dev, dtype=torch.float16)
    a = t(lambda: _lib.call("s3d_ffmlp_forward", x, w, B, cin, cout, hid, nl, 0, 6, fb, out))
    b = t(lambda: _lib.call("s3d_ffmlp_inference", x, w, B, cin, cout, hid, nl, 0, 6, None, out))
    g = torch.randn(B, cout, device=dev).half()
    bb = torch.empty_like(fb); gi = torch.empty_like(x); gw = torch.empty_like(w)
    c = t(lambda: _lib.call("s3d_ffmlp_backward", g, x, w, fb, B, cin, cout, hid, nl, 0, 6, 1, bb, gi, gw))
    print("ffmlp %d-%dx%d-%d  B=%d: forward %.3f ms  inference (no buffer) %.3f ms  backward %.3f ms" % (cin, hid, nl, cout, B, a, b, c), flush=True)
x = torch.randn(B, 144, device=dev).half(); w = (torch.randn(32, 144, device=dev) * 0.05).half(); y = torch.empty(B, 32, device=dev, dtype=torch.float16)
print("linear 144->32: forward %.3f ms" % t(lambda: _lib.call("s3d_linear_forward", x, w, B, 144, 32, y)))
gy = torch.randn(B, 32, device=dev).half(); gx = torch.empty_like(x); gw = torch.empty_like(w)
print("linear 144->32: backward %.3f ms" % t(lambda: _lib.call("s3d_linear_backward", gy, x, w, B, 144, 32, gx, gw)))
