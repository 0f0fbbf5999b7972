"""Time the fused Heun step kernels (adb_edm_heun_mid / adb_edm_heun_post) on a state much larger than L2:
python tools/time_step.py [elements] [reps]   — prints achieved GB/s against the algorithmic 32 B per element per Heun step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodiffuser_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
lib, st = N.lib(), N.stream_ptr(dev)
x, f1, f2, d, x1, out = (torch.randn(n, device=dev) for _ in range(6))


def step():
    N.check(lib.adb_edm_heun_mid(N.ptr(x), N.ptr(f1), 2.0, 0.2, -0.5, N.ptr(d), N.ptr(x1), n, st))      # r x,F  w d,x1 : 16 B
    N.check(lib.adb_edm_heun_post(N.ptr(x), N.ptr(d), N.ptr(f2), 1.5, 0.2, -0.5, N.ptr(out), n, st))   # r x,d,F w x   : 16 B


for _ in range(2):
    step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(reps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"heun mid+post on {n} elements: {ms:.3f} ms per step pair -> {32.0 * n / ms / 1e6:.1f} GB/s algorithmic")
