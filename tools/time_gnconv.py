"""Time adb_cl_gn_conv3 (fused GroupNorm-apply convolution) against the unfused GroupNorm + adb_cl_conv on U-Net shapes:
python tools/time_gnconv.py [B] ; with ADB_LIB=debug ADB_DEBUG_FLAGS=2 also prints the in-kernel cycle accounting."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodiffuser_b200 import _native as N
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
lib, st = N.lib(), N.stream_ptr(dev)
dbg = os.environ.get("ADB_DEBUG_FLAGS") and int(os.environ["ADB_DEBUG_FLAGS"]) & 2
raw = ctypes.CDLL(N.LIB_PATH)
names = ["mma:wait tempty", "mma:wait aready", "mma:wait bfull", "mma:total", "prod:wait aempty", "prod:wait bempty", "prod:total",
         "xf:wait araw", "xf:work", "xf:total", "epi:wait tfull", "epi:work", "epi:total"]
G, s2 = 8, 2 ** -0.5
for (L, C1, C2, Nn, res) in [(4096, 256, 0, 256, False), (4096, 256, 0, 256, True), (4096, 256, 256, 256, False), (1024, 512, 0, 512, True),
                             (1024, 512, 512, 512, False)]:
    Cin = C1 + C2
    h = torch.randn(B, L, C1, device=dev).to(torch.bfloat16)
    sk = torch.randn(B, L, C2, device=dev).to(torch.bfloat16) if C2 else None
    gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
    ss = 0.1 * torch.randn(B, 2 * Cin, device=dev)
    w = torch.randn(3, Cin, Nn, device=dev) / math.sqrt(3 * Cin)
    bias = torch.zeros(Nn, device=dev)
    r = torch.randn(B, L, Nn, device=dev).to(torch.bfloat16) if res else None
    packed = torch.empty(lib.adb_cl_conv_packed_elems(Cin, Nn, 3), dtype=torch.bfloat16, device=dev)
    N.check(lib.adb_cl_pack_conv_weights(N.ptr(w), N.ptr(packed), Cin, Nn, 3, st))
    packed16 = torch.empty(lib.adb_cl_conv_packed_elems(Cin, Nn, 3), dtype=torch.float16, device=dev)
    N.check(lib.adb_cl_pack_conv_weights_f16(N.ptr(w), N.ptr(packed16), Cin, Nn, 3, st))
    sums = torch.zeros(B * G * 2, dtype=torch.float64, device=dev)
    tickets = torch.zeros(B, dtype=torch.int32, device=dev)
    coef = torch.empty(2, B, Cin, dtype=torch.float32, device=dev)
    g1 = C1 // (Cin // G)
    def stats():
        N.check(lib.adb_cl_gn_coef(N.ptr(h), N.ptr(sums), N.ptr(tickets), N.ptr(coef), B, L, C1, g1, G, 0, 0, Cin, N.ptr(gamma), N.ptr(beta),
                                   N.ptr(ss), 2 * Cin, 1e-5, 1.0, st))
        if C2:
            N.check(lib.adb_cl_gn_coef(N.ptr(sk), N.ptr(sums), N.ptr(tickets), N.ptr(coef), B, L, C2, G - g1, G, g1, C1, Cin, N.ptr(gamma),
                                       N.ptr(beta), N.ptr(ss), 2 * Cin, 1e-5, s2, st))
    stats()
    out = torch.empty(B, L, Nn, dtype=torch.bfloat16, device=dev)
    def fused():
        N.check(lib.adb_cl_gn_conv3(N.ptr(h), C1, N.ptr(sk), C2, N.ptr(coef), N.ptr(packed16), N.ptr(bias), N.ptr(r), N.ptr(out), B, L, Nn, st))
    cat = torch.cat([h, sk], dim=2).contiguous() if C2 else h
    def plain():
        N.check(lib.adb_cl_conv(N.ptr(cat), N.ptr(packed), N.ptr(bias), N.ptr(r), N.ptr(out), B, L, L, Cin, Nn, 3, -1, 1, 0, 0, 0, 0, 1, st))
    res_t = {}
    for name, fn in (("fused", fused), ("plain conv", plain), ("stats+coef", stats)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res_t[name] = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * B * L * Cin * Nn * 3
    print(f"B={B} L={L} Cin={C1}+{C2} N={Nn} res={res}: fused {res_t['fused']:.1f} us ({fl/res_t['fused']/1e6:.0f} TF/s)  "
          f"plain conv alone {res_t['plain conv']:.1f} us ({fl/res_t['plain conv']/1e6:.0f} TF/s)  stats+coef {res_t['stats+coef']:.1f} us")
    if dbg:
        buf = (ctypes.c_ulonglong * 16)()
        raw.adb_debug_tc_cycles(buf, 1)
        fused()
        raw.adb_debug_tc_cycles(buf, 1)
        items = ((B * ((L + 127) // 128) + 1) // 2) * (Nn // 256 if Nn % 256 == 0 else 1) * (Cin // 64)     # leader CTAs only
        print("   cycles per (tile pair, K-block), summed over the pair leaders / items:")
        for i, n in enumerate(names):
            print(f"   {n:20s} {buf[i] / items:9.0f}")
