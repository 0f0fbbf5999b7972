"""Time the UNet1d (BASELINE config 4 architecture) forward: python tools/time_unet.py [B] [L] [precision] [reps]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodiffuser_b200 import UNet1dBase, _native
from oracle.weights import make_unet1d_state_dict, UNET1D_CONFIG4
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
net = UNet1dBase(precision=prec, **UNET1D_CONFIG4)
net.load_state_dict(make_unet1d_state_dict(UNET1D_CONFIG4, 0), strict=True)
net = net.to(dev)
net.use_cuda_graph = os.environ.get("ADB_NO_GRAPH") is None
net.fuse_groupnorm = os.environ.get("ADB_UNET_NOFUSE") is None      # tools only: time the separate GroupNorm passes
x = torch.randn(B, 2, L, device=dev); t = torch.zeros(B, device=dev)
for _ in range(2): y = net(x, t)
_native.check_async()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
for _ in range(reps): net(x, t)
e1.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / reps
ms = e0.elapsed_time(e1) / reps
fl = 61.69e9 * B * L / 262144
print(f"unet1d cfg4 B={B} L={L} {prec} graph={net.use_cuda_graph} fuse_gn={net.fuse_groupnorm}: {ms:.3f} ms/eval device, {wall*1e3:.3f} ms wall -> {fl/ms/1e9:.1f} TFLOP/s")
