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
import subprocess, threading
samples, halt = [], threading.Event()
def poll():
    while not halt.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            samples.append((float(o[0]), float(o[1])))
        except Exception:
            pass
        halt.wait(0.1)
th = threading.Thread(target=poll, daemon=True); th.start()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps): net(x, t)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
halt.set(); th.join(timeout=6)
tm = net.timers()
fl = (606.093e9 - (36 - layers) * 16.777e9) * B
cs = (tm['conv'][0] + tm['skip'][0]) / reps
print(f"B={B} layers={layers} kernel={os.environ.get('ADB_BLOCK_KERNEL', '3')} flags={os.environ.get('ADB_DEBUG_FLAGS')}: {dt*1e3:.2f} ms/eval "
      f"{fl/dt/1e12:.1f} TFLOP/s; block {tm['conv'][0]/reps:.2f} ms + skip gemm {tm['skip'][0]/reps:.2f} ms = {cs:.2f} ms "
      f"({fl/(cs*1e-3)/1e12:.1f} TF/s) aux {tm['aux'][0]/reps:.2f} ms tail {tm['tail'][0]/reps:.2f} ms; "
      f"clocks/power samples (MHz, W): {sorted(samples)[len(samples)//2:][:1]} n={len(samples)} max_w={max([s[1] for s in samples], default=0):.0f}")
if os.environ.get("ADB_DEBUG_FLAGS") and int(os.environ["ADB_DEBUG_FLAGS"]) & 2:
    import ctypes
    lib = ctypes.CDLL(_native.LIB_PATH)
    buf = (ctypes.c_ulonglong * 16)()
    lib.adb_debug_tc_cycles(buf, 1)
    net(x, t)
    lib.adb_debug_tc_cycles(buf, 1)
    names = ["mma:wait tempty(G1)", "mma:wait tempty(G2)", "mma:wait zready", "mma:wait full", "mma:total", "prod:wait empty",
             "prod:total", "epi:wait tfull(G1)", "epi:E1 work", "epi:wait tfull(G2)", "epi:E2 work", "epi:total", "prod:wait flags (ML)"]
    ntiles = B * 125 * layers
    for i, n in enumerate(names):
        print(f"  {n:24s} {buf[i] / ntiles:10.0f} cycles/tile")
