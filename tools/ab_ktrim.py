import os, sys, torch
sys.path.insert(0, os.getcwd())
from audiodiffuser_b200 import UNet1dBase, _native
from oracle.weights import make_unet1d_state_dict, UNET1D_CONFIG4
dev = torch.device("cuda:0")
net = UNet1dBase(precision="bf16", **UNET1D_CONFIG4)
net.load_state_dict(make_unet1d_state_dict(UNET1D_CONFIG4, 0), strict=True)
net = net.to(dev)
x = torch.randn(128, 2, 262144, device=dev); t = torch.zeros(128, device=dev)
def timeit(tag):
    for _ in range(2): net(x, t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): net(x, t)
    e1.record(); torch.cuda.synchronize()
    print(tag, e0.elapsed_time(e1) / 10, "ms/eval")
timeit("ktrim on ")
P = net._pack()
n = 0
for k, v in P.items():
    if isinstance(v, dict) and "ktrim" in v:
        del v["ktrim"]; n += 1
net._graphs = {}
print("removed ktrim from", n, "convs")
timeit("ktrim off")
