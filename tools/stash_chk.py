"""Pair kernel: outputs with the fp16 skip stash (default) against ADB_NO_STASH=1 and against the reference goldens."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2
from audiodiffuser_b200 import WaveNetNoise, _native as N
from oracle.weights import make_wavenet_state_dict
dev = torch.device("cuda:0")
def make(C, layers, cycle, seed):
    net = WaveNetNoise(C, layers, cycle, precision="bf16")
    net.load_state_dict(make_wavenet_state_dict(C, layers, seed), strict=True)
    return net.to(dev)
for name in ["wavenet_c256_l3", "wavenet_c256_l13_dil2048", "wavenet_c256_l2_short"]:
    g = load_golden(name)
    C, layers, cycle, B, L, seed = (int(v) for v in g["cfg"])
    audio, t = torch.from_numpy(g["audio"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    os.environ["ADB_NO_STASH"] = "1"
    plain = make(C, layers, cycle, seed)(audio, t)
    del os.environ["ADB_NO_STASH"]
    out = make(C, layers, cycle, seed)(audio, t)
    N.check_async()
    print(name, layers, "stash vs plain", rel_l2(out, plain), "stash vs golden", rel_l2(out, g["out"]), "plain vs golden", rel_l2(plain, g["out"]))
