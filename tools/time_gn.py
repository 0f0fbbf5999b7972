import os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from audiodiffuser_b200 import _native as N
dev = torch.device("cuda:0"); lib = N.lib(); st = N.stream_ptr(dev)
def run(B, L, C, use_ss, dt, reps=200):
    adt = torch.bfloat16 if dt else torch.float32
    x = torch.randn(B, L, C, device=dev).to(adt); o = torch.empty_like(x)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    ss = torch.randn(B, 4 * C, device=dev) * 0.1
    sums = torch.empty(B * 8 * 2, dtype=torch.float64, device=dev)
    def f():
        N.check(lib.adb_cl_groupnorm(N.ptr(x), N.ptr(g), N.ptr(b), N.ptr(ss) if use_ss else ctypes.c_void_p(0), 4 * C, N.ptr(o), N.ptr(sums), B, L, C, 8, 1e-5, 2, dt, st))
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"GN B={B} L={L} C={C} ss={use_ss} dtype={'bf16' if dt else 'f32'}: {us:.1f} us per call ({x.numel()*x.element_size()*3/us/1e3:.0f} GB/s eff)")
for shape in [(8, 256, 512), (8, 4096, 256), (8, 4096, 512), (16, 4096, 512)]:
    for ss in (False, True):
        run(*shape, ss, 1)
run(8, 4096, 256, True, 0)
