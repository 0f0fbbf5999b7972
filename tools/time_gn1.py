import os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from audiodiffuser_b200 import _native as N
dev = torch.device("cuda:0"); lib = N.lib(); st = N.stream_ptr(dev)
B, L, C = 16, 4096, 512
for dt, adt in ((1, torch.bfloat16), (0, torch.float32)):
    x = torch.randn(B, L, C, device=dev).to(adt); o = torch.empty_like(x)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    sums = torch.empty(B * 8 * 2, dtype=torch.float64, device=dev)
    for _ in range(3):
        N.check(lib.adb_cl_groupnorm(N.ptr(x), N.ptr(g), N.ptr(b), ctypes.c_void_p(0), 0, N.ptr(o), N.ptr(sums), B, L, C, 8, 1e-5, 2, dt, st))
torch.cuda.synchronize()
print("ok")
