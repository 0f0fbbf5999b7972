"""Time the full-size bf16 network forward (B from argv) — quick perf probe."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodiffuser_b200 import WaveNetNoise, _native
from oracle.weights import make_wavenet_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 36
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
net = WaveNetNoise(256, layers, 12, precision="bf16")
net.load_state_dict(make_wavenet_state_dict(256, layers, 0), strict=True)
net = net.to(dev)
x = torch.randn(B, 16000, device=dev); t = torch.zeros(B, device=dev)
net(x, t); _native.check_async()
net.set_timing(True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps): net(x, t)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
tm = net.timers()
fl = (606.093e9 - (36 - layers) * 16.777e9) * B
print(f"B={B} layers={layers} flags={os.environ.get('ADB_DEBUG_FLAGS')}: {dt*1e3:.2f} ms/eval {fl/dt/1e12:.1f} TFLOP/s; conv {tm['conv'][0]/reps:.2f} ms ({fl/(tm['conv'][0]/reps*1e-3)/1e12:.1f} TF/s) aux {tm['aux'][0]/reps:.2f} ms")
if os.environ.get("ADB_DEBUG_FLAGS") and int(os.environ["ADB_DEBUG_FLAGS"]) & 2:
    import ctypes
    lib = ctypes.CDLL(_native.LIB_PATH)
    buf = (ctypes.c_ulonglong * 16)()
    lib.adb_debug_tc_cycles(buf, 1)
    net(x, t)
    lib.adb_debug_tc_cycles(buf, 1)
    names = ["mma:wait tempty(G1)", "mma:wait tempty(G2)", "mma:wait zready", "mma:wait full", "mma:total", "prod:wait empty",
             "prod:total", "epi:wait tfull(G1)", "epi:E1 work", "epi:wait tfull(G2)", "epi:E2 work", "epi:total"]
    ntiles = B * 125 * layers
    for i, n in enumerate(names):
        print(f"  {n:24s} {buf[i] / ntiles:10.0f} cycles/tile")
