import os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from audiodiffuser_b200 import _native as N
dev = torch.device("cuda:0"); lib = N.lib(); st = N.stream_ptr(dev)
for (B, L, C) in [(32, 32, 512), (32, 256, 512), (32, 4096, 256)]:
    x = torch.randn(B, L, C, device=dev).to(torch.bfloat16); o = torch.empty_like(x)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    sums = torch.empty(B * 8 * 2, dtype=torch.float64, device=dev)
    for _ in range(3):
        N.check(lib.adb_cl_groupnorm(N.ptr(x), N.ptr(g), N.ptr(b), ctypes.c_void_p(0), 0, N.ptr(o), N.ptr(sums), B, L, C, 8, 1e-5, 2, 1, st))
torch.cuda.synchronize()
print("ok")
