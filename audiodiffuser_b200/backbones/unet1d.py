"""1-D U-Net backbone behind the reference's class API, executed by the sm_100a kernels of libadb200.

Mirrors `src/models/backbones/unet1d.py:818-893` (`UNet1dBase`, wrapping `UNet1d` :624-816) for the
unconditional configuration: same constructor kwargs, same parameter names and shapes (so a reference
checkpoint loads with `load_state_dict(strict=True)`), same `forward(x, t, ..., cond_drop_prob=None)`
protocol that `EluDiffusion.denoise_fn` drives (diffusion.py:50).

Internally activations are channels-last `[B][L][C]` (fp32, or bf16 on the tensor-core path) and every
convolution / linear layer is one call of the generic GEMM-convolution `adb_cl_conv` (include/adb200.h); this
file only sequences the C-ABI calls in the order of `UNet1d.forward` (unet1d.py:769-816) and re-lays the
weights once per parameter version. There is no PyTorch compute on the path and no CPU fallback.
"""
import math
from collections import OrderedDict
from ctypes import c_void_p
from typing import Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from .. import _native as N

ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU = 0, 1, 2, 3


def _param_shapes(cfg) -> "OrderedDict[str, tuple]":
    """Parameter names / shapes of UNet1d in registration order (unet1d.py:650-767)."""
    ch, m = cfg["channels"], list(cfg["multipliers"])
    nf, W, cin = cfg["num_filters"], cfg["window_length"], cfg["in_channels"]
    cout = cfg.get("out_channels") or cin
    T, am, km = ch * 4, cfg["attention_multiplier"], cfg["kernel_multiplier_downsample"]
    cdim = ch * 4 if cfg.get("class_cond") else 0                    # classes_channels, unet1d.py:843-850
    s = OrderedDict()

    def resnet(p, ci, co):
        s[p + ".to_cond_embedding.1.weight"] = (2 * co, T + cdim)
        s[p + ".to_cond_embedding.1.bias"] = (2 * co,)
        for blk, c_in in (("block1", ci), ("block2", co)):
            s[f"{p}.{blk}.groupnorm.weight"] = (c_in,)
            s[f"{p}.{blk}.groupnorm.bias"] = (c_in,)
            s[f"{p}.{blk}.project.weight"] = (co, c_in, 3)
            s[f"{p}.{blk}.project.bias"] = (co,)
        if ci != co:
            s[p + ".to_out.weight"] = (co, ci, 1)
            s[p + ".to_out.bias"] = (co,)

    def transformer(p, c):
        s[p + ".norm.weight"] = (c,)
        s[p + ".norm.bias"] = (c,)
        s[p + ".attention.to_q.weight"] = (c, c)
        s[p + ".attention.to_kv.weight"] = (2 * c, c)
        s[p + ".attention.to_out.weight"] = (c, c)
        s[p + ".feed_forward.0.g"] = (1, c, 1)
        s[p + ".feed_forward.1.weight"] = (c * am, c, 1)
        s[p + ".feed_forward.3.g"] = (1, c * am, 1)
        s[p + ".feed_forward.4.weight"] = (c, c * am, 1)

    s["to_in.to_in.weight"] = (nf, cin, W)
    s["to_out.to_out.weight"] = (nf, cout, W)
    s["to_time.0.0.weights"] = (ch // 2,)
    s["to_time.0.1.weight"] = (T, ch + 1)
    s["to_time.0.1.bias"] = (T,)
    s["to_time.2.weight"] = (T, T)
    s["to_time.2.bias"] = (T,)
    n = len(m) - 1
    for i in range(n):
        ci, co, f = ch * m[i], ch * m[i + 1], cfg["factors"][i]
        p = f"downsamples.{i}"
        s[p + ".downsample.weight"] = (co, ci, f * km + 1)
        s[p + ".downsample.bias"] = (co,)
        for j in range(cfg["num_blocks"][i]):
            resnet(f"{p}.blocks.{j}", co, co)
        if cfg["attentions"][i]:
            transformer(p + ".transformer", co)
    cb = ch * m[-1]
    resnet("bottleneck.pre_block", cb, cb)
    if cfg["use_attention_bottleneck"]:
        transformer("bottleneck.transformer", cb)
    resnet("bottleneck.post_block", cb, cb)
    for u, i in enumerate(reversed(range(n))):
        ci, co, f = ch * m[i + 1], ch * m[i], cfg["factors"][i]
        p = f"upsamples.{u}"
        for j in range(cfg["num_blocks"][i] + (1 if cfg["attentions"][i] else 0)):
            resnet(f"{p}.blocks.{j}", 2 * ci, ci)
        if cfg["attentions"][i]:
            transformer(p + ".transformer", ci)
        s[p + ".upsample.weight"] = (co, ci, 3) if f == 1 else (ci, co, 2 * f)   # ConvTranspose1d: [Cin][Cout][k]
        s[p + ".upsample.bias"] = (co,)
    return s


def _init_param(name: str, shape) -> Tensor:
    """torch's default initialisers for the layer types the reference uses."""
    if name.endswith("to_time.0.0.weights"):
        return torch.randn(shape)                                    # unet1d.py:135
    if name == "to_out.to_out.weight":
        return torch.zeros(shape)                                    # unet1d.py:619
    if name.endswith((".g", "groupnorm.weight", "norm.weight")):
        return torch.ones(shape)
    if name.endswith(("groupnorm.bias", "norm.bias")):
        return torch.zeros(shape)
    if name.endswith(".bias"):
        return None                                                  # filled from the matching weight's fan-in
    fan_in = shape[1] * (shape[2] if len(shape) == 3 else 1)
    return (torch.rand(shape) * 2 - 1) / math.sqrt(fan_in)


class _Node(nn.Module):
    """Name-only container: reproduces the reference's module tree so state_dict keys match."""


def _register(root: nn.Module, dotted: str, value: Tensor):
    parts = dotted.split(".")
    node = root
    for part in parts[:-1]:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    node.register_parameter(parts[-1], nn.Parameter(value))


class UNet1dBase(nn.Module):
    """`UNet1dBase(channels, cond_drop_prob, ..., **unet_kwargs)` — unet1d.py:818-854.

    Extra optional kwarg `precision`: "bf16" (tcgen05 tensor-core path) or "fp32" (CUDA-core path).
    Conditioning branches (class_cond / text_cond / use_condition_block / inj_*) and
    `use_nearest_upsample` are outside the fused path and raise NotImplementedError.
    """

    def __init__(self, channels: int, cond_drop_prob: float, num_classes: int = None, class_embed_dim: int = None,
                 class_cond: bool = False, text_cond: bool = False, max_text_len: int = None, text_embed_dim=768,
                 text_cond_multiplier: int = None, use_self_text_cond: bool = False, use_condition_block: bool = False,
                 precision: str = "bf16", **kwargs):
        super().__init__()
        if text_cond or use_condition_block or (class_cond and num_classes is None):
            raise NotImplementedError("the fused UNet1d covers the unconditional and the label-conditioned (num_classes) "
                                      "configurations; text conditioning, class embeddings and the condition block are not built")
        if kwargs.get("use_nearest_upsample", False):
            raise NotImplementedError("use_nearest_upsample=True (nn.Upsample + ReflectionPad1d, unet1d.py:234-245) is not built")
        if precision not in N.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(N.PRECISIONS)}")
        required = ("num_filters", "window_length", "stride", "in_channels", "multipliers", "factors", "num_blocks",
                    "attentions", "attention_heads", "attention_multiplier", "resnet_groups", "kernel_multiplier_downsample",
                    "use_skip_scale", "use_attention_bottleneck")
        missing = [k for k in required if k not in kwargs]
        if missing:
            raise TypeError(f"UNet1d missing required arguments: {missing}")       # like the reference's signature
        cfg = dict(channels=channels, class_cond=bool(class_cond), num_classes=num_classes, **kwargs)
        n = len(cfg["multipliers"]) - 1
        assert len(cfg["factors"]) == n and len(cfg["attentions"]) == n and len(cfg["num_blocks"]) == n   # unet1d.py:672-675
        if cfg["num_filters"] != channels * cfg["multipliers"][0]:
            raise ValueError("num_filters must equal channels * multipliers[0] (to_in feeds the first down block)")
        if cfg["kernel_multiplier_downsample"] % 2 != 0:
            raise AssertionError("Kernel multiplier must be even")                  # unet1d.py:217
        self.cfg = cfg
        self.cond_drop_prob = cond_drop_prob
        self.precision = precision
        if class_cond:                                            # LabelEmbedder, conditioner.py:59-92 (registered before the unet)
            cd = channels * 4
            self.label_conditioner = _Node()
            for name, v in (("null_classes_emb", torch.randn(1, channels)), ("label_emb.weight", torch.randn(num_classes, channels)),
                            ("class_to_cond.0.weight", torch.ones(channels)), ("class_to_cond.0.bias", torch.zeros(channels)),
                            ("class_to_cond.1.weight", (torch.rand(cd, channels) * 2 - 1) / math.sqrt(channels)),
                            ("class_to_cond.1.bias", (torch.rand(cd) * 2 - 1) / math.sqrt(channels)),
                            ("class_to_cond.3.weight", (torch.rand(cd, cd) * 2 - 1) / math.sqrt(cd)),
                            ("class_to_cond.3.bias", (torch.rand(cd) * 2 - 1) / math.sqrt(cd))):
                _register(self.label_conditioner, name, v)
        self.unet = _Node()
        shapes = _param_shapes(cfg)
        for name, shape in shapes.items():
            v = _init_param(name, shape)
            if v is None:                                                            # bias: U(+-1/sqrt(fan_in)) of its weight
                wshape = shapes[name[:-4] + "weight"]
                fan_in = wshape[1] * (wshape[2] if len(wshape) == 3 else 1)
                v = (torch.rand(shape) * 2 - 1) / math.sqrt(fan_in)
            _register(self.unet, name, v)
        self._packed = None
        self._packed_key = None
        self._graphs = {}
        self.use_cuda_graph = True
        self.fuse_groupnorm = True      # bf16: GroupNorm apply inside the consumer convolution (adb_cl_gn_conv3); False = separate passes
        self.graph_launches = 0          # kernel launches replayed through CUDA graphs (invisible to adb_launch_count)

    # ---- weight re-layout (once per parameter version) ----------------------------------------------
    def _param_key(self):
        return (self.precision,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def __getstate__(self):
        """copy.deepcopy / pickle carry the parameters only (EMA snapshots, diffunet_complex_module.py:162-167): packed
        weights, workspaces and captured CUDA graphs are rebuilt lazily by the copy."""
        state = self.__dict__.copy()
        state.update(_packed=None, _packed_key=None, _graphs={})
        return state

    def _pack(self):
        key = self._param_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        sd = {k[len("unet."):]: v.detach().to(torch.float32) for k, v in self.state_dict().items() if k.startswith("unet.")}
        lc = {k[len("label_conditioner."):]: v.detach().to(torch.float32).contiguous() for k, v in self.state_dict().items()
              if k.startswith("label_conditioner.")}
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.AdbError("UNet1dBase parameters are on the CPU: adb200 has no CPU path; call .cuda() on a B200")
        N.ensure_device(dev)
        cfg, bf16 = self.cfg, self.precision in ("bf16", "bfloat16")
        lib, st = N.lib(), N.stream_ptr(dev)
        P = {}

        def gemm_weight(w3: Tensor, bias: Optional[Tensor], meta: dict):
            """w3: fp32 [taps][Cin][N] on device."""
            w3 = w3.contiguous()
            taps, cin, n = w3.shape
            ent = dict(meta, taps=taps, cin=cin, n=n, bias=bias.contiguous() if bias is not None else None)
            if bf16:
                packed = torch.empty(lib.adb_cl_conv_packed_elems(cin, n, taps), dtype=torch.bfloat16, device=dev)
                N.check(lib.adb_cl_pack_conv_weights(N.ptr(w3), N.ptr(packed), cin, n, taps, st))
                ent["w"], ent["_keep"] = packed, w3
            else:
                ent["w"] = w3
            return ent

        def conv_k(name, bias_name=None):          # nn.Conv1d weight [Cout][Cin][k], stride 1, "same"
            w = sd[name]
            k = w.shape[2]
            return gemm_weight(w.permute(2, 1, 0), sd.get(bias_name), dict(off0=-(k // 2), dil=1, ups=0))

        def linear(name):                          # bias-free nn.Linear weight [N][K]
            return gemm_weight(sd[name].t().unsqueeze(0), None, dict(off0=0, dil=1, ups=0))

        def down(name, bias_name, f):              # Conv1d k = f*km + 1, stride f, pad f*km/2 on the [L/f][f*Cin] view
            w = sd[name]
            co, ci, k = w.shape
            km = cfg["kernel_multiplier_downsample"]
            wc = torch.zeros(km + 1, f * ci, co, device=dev)
            for j in range(km + 1):
                for ph in range(f):
                    kk = j * f + ph
                    if kk < k:
                        wc[j, ph * ci:(ph + 1) * ci, :] = w[:, :, kk].t()
            meta = dict(off0=-(km // 2), dil=1, ups=0, f=f)
            if bf16 and ci % 64 == 0 and k == f * km + 1:
                meta["ktrim"] = ci // 64            # the last coarse tap holds ONE fine tap: only its first ci channels are non-zero
            return gemm_weight(wc, sd[bias_name], meta)

        def up(name, bias_name, f):                # ConvTranspose1d weight [Cin][Cout][2f], stride f
            w = sd[name]
            ci, co, k = w.shape
            wt = torch.empty(2, ci, f * co, device=dev)
            for ph in range(f):
                wt[0, :, ph * co:(ph + 1) * co] = w[:, :, ph]
                wt[1, :, ph * co:(ph + 1) * co] = w[:, :, ph + f]
            return gemm_weight(wt, sd[bias_name].repeat(f), dict(off0=0, dil=-1, ups=f, shift=f // 2 + f % 2, cout=co,
                                                                  out_pad=f % 2))

        cond_w, cond_b, cond_off = [], [], {}

        def resnet(p, cat=False):
            co = sd[p + ".block1.project.weight"].shape[0]
            cond_off[p] = sum(t.shape[0] for t in cond_w)
            cond_w.append(sd[p + ".to_cond_embedding.1.weight"])
            cond_b.append(sd[p + ".to_cond_embedding.1.bias"])
            P[p + ".block1"] = conv_k(p + ".block1.project.weight", p + ".block1.project.bias")
            P[p + ".block2"] = conv_k(p + ".block2.project.weight", p + ".block2.project.bias")
            if bf16:
                # fp16 copy of the packed weights for the fused GroupNorm convolution (its activation operand is produced in fp16)
                for blk in ("block1", "block2"):
                    ent = P[f"{p}.{blk}"]
                    if ent["cin"] % 64 == 0 and ent["n"] % 64 == 0:
                        w16 = torch.empty(lib.adb_cl_conv_packed_elems(ent["cin"], ent["n"], ent["taps"]), dtype=torch.float16, device=dev)
                        N.check(lib.adb_cl_pack_conv_weights_f16(N.ptr(ent["_keep"]), N.ptr(w16), ent["cin"], ent["n"], ent["taps"], st))
                        ent["w16"] = w16
            for blk in ("block1", "block2"):
                P[f"{p}.{blk}.gn"] = (sd[f"{p}.{blk}.groupnorm.weight"].contiguous(), sd[f"{p}.{blk}.groupnorm.bias"].contiguous())
            if p + ".to_out.weight" in sd:
                P[p + ".to_out"] = conv_k(p + ".to_out.weight", p + ".to_out.bias")
                if cat and bf16:
                    # the block reads cat(x, skip * skip_scale) (unet1d.py:552-556) as TWO inputs: the constant goes into the
                    # weight rows of the skip channels (the first `co` input channels are x, UpsampleBlock1d :507-509)
                    w = sd[p + ".to_out.weight"].permute(2, 1, 0).clone()
                    w[:, co:, :] *= 2 ** -0.5 if cfg["use_skip_scale"] else 1.0
                    P[p + ".to_out_cat"] = gemm_weight(w, sd[p + ".to_out.bias"], dict(off0=0, dil=1, ups=0))
            P[p + ".co"] = co

        def transformer(p):
            P[p + ".norm"] = (sd[p + ".norm.weight"].contiguous(), sd[p + ".norm.bias"].contiguous())
            for nm in ("to_q", "to_kv", "to_out"):
                P[f"{p}.{nm}"] = linear(f"{p}.attention.{nm}.weight")
            P[p + ".ff0"] = sd[p + ".feed_forward.0.g"].reshape(-1).contiguous()
            P[p + ".ff1"] = conv_k(p + ".feed_forward.1.weight")
            P[p + ".ff3"] = sd[p + ".feed_forward.3.g"].reshape(-1).contiguous()
            P[p + ".ff4"] = conv_k(p + ".feed_forward.4.weight")

        for k in ("to_in.to_in.weight", "to_out.to_out.weight", "to_time.0.0.weights", "to_time.0.1.weight", "to_time.0.1.bias",
                  "to_time.2.weight", "to_time.2.bias"):
            P[k] = sd[k].contiguous()
        n = len(cfg["multipliers"]) - 1
        for i in range(n):
            p = f"downsamples.{i}"
            P[p + ".downsample"] = down(p + ".downsample.weight", p + ".downsample.bias", cfg["factors"][i])
            for j in range(cfg["num_blocks"][i]):
                resnet(f"{p}.blocks.{j}")
            if cfg["attentions"][i]:
                transformer(p + ".transformer")
        resnet("bottleneck.pre_block")
        if cfg["use_attention_bottleneck"]:
            transformer("bottleneck.transformer")
        resnet("bottleneck.post_block")
        for u, i in enumerate(reversed(range(n))):
            p, f = f"upsamples.{u}", cfg["factors"][i]
            for j in range(cfg["num_blocks"][i] + (1 if cfg["attentions"][i] else 0)):
                resnet(f"{p}.blocks.{j}", cat=True)
            if cfg["attentions"][i]:
                transformer(p + ".transformer")
            P[p + ".upsample"] = conv_k(p + ".upsample.weight", p + ".upsample.bias") if f == 1 else \
                up(p + ".upsample.weight", p + ".upsample.bias", f)
        nf, W, S = cfg["num_filters"], cfg["window_length"], cfg["stride"]
        cout_w = cfg.get("out_channels") or cfg["in_channels"]
        if bf16 and nf % 64 == 0 and W == 2 * S and S * cout_w <= 64:       # WAVdec on the tensor cores
            pk = torch.empty(lib.adb_cl_wavdec_packed_elems(nf), dtype=torch.bfloat16, device=dev)
            N.check(lib.adb_cl_wavdec_pack(N.ptr(P["to_out.to_out.weight"]), N.ptr(pk), nf, cout_w, W, S, st))
            P["to_out.packed"] = pk
        cin_w = cfg["in_channels"]
        if bf16 and W == 2 * S and (W * cin_w) % 64 == 0 and (2 * nf) % 64 == 0:   # WAVenc on the tensor cores (see adb_cl_wavenc_prep)
            wk = sd["to_in.to_in.weight"].permute(2, 1, 0).reshape(W * cin_w, nf)   # row k*Cin + c, column f
            wc = torch.zeros(2, W * cin_w, 2 * nf, device=dev)
            wc[0, :, :nf] = wk                                    # even frame 2m: buffer row m
            wc[0, S * cin_w:, nf:] = wk[:S * cin_w]               # odd frame 2m+1: second half of row m ...
            wc[1, :S * cin_w, nf:] = wk[S * cin_w:]               # ... and first half of row m+1
            P["to_in.tc"] = gemm_weight(wc, None, dict(off0=0, dil=1, ups=0))
        P["cond_w"] = torch.cat(cond_w, dim=0).contiguous()
        if bf16 and P["cond_w"].shape[0] % 64 == 0 and P["cond_w"].shape[1] % 64 == 0:
            # all (scale, shift) projections as one tensor-core GEMM [B][T] x [T][sum 2C] (bias in the epilogue)
            P["cond_tc"] = gemm_weight(P["cond_w"].t().unsqueeze(0), torch.cat(cond_b, dim=0), dict(off0=0, dil=1, ups=0))
        P["cond_b"] = torch.cat(cond_b, dim=0).contiguous()
        P["cond_off"] = cond_off
        P["lc"] = lc
        torch.cuda.current_stream(dev).synchronize()
        self._packed, self._packed_key = P, key
        self._graphs = {}
        return P

    # ---- one forward as a sequence of C-ABI calls -------------------------------------------------
    def _run(self, x: Tensor, t: Tensor, out: Tensor, classes: Optional[Tensor] = None, drop: Optional[Tensor] = None,
             in_scale: Optional[Tensor] = None):
        """in_scale (optional, fp32 [B]): per-sample factor on x folded into the WAVenc re-layout (the EDM input scale c_in of
        the device-resident trajectory); only honoured on the tensor-core WAVenc path (`_wavenc_tc_ok`)."""
        cfg, P = self.cfg, self._pack()
        lib, dev = N.lib(), x.device
        st = N.stream_ptr(dev)
        bf16 = self.precision in ("bf16", "bfloat16")
        dt, adt = (1, torch.bfloat16) if bf16 else (0, torch.float32)
        B, cin, L = x.shape
        groups, heads = cfg["resnet_groups"], cfg["attention_heads"]
        W, S = cfg["window_length"], cfg["stride"]
        sums = torch.zeros(B * groups * 2, dtype=torch.float64, device=dev)     # adb_cl_groupnorm clears it itself; adb_cl_gn_coef leaves it zero
        tickets = torch.zeros(B, dtype=torch.int32, device=dev)

        def conv(h, ent, act=ACT_NONE, res=None):
            Bh, Lh, Ch = h.shape
            if "f" in ent:                                      # strided conv on the [L/f][f*C] view
                f = ent["f"]
                if Lh % f:
                    raise N.AdbError(f"length {Lh} is not a multiple of the down-sampling factor {f}")
                Lh, Ch = Lh // f, Ch * f
            assert Ch == ent["cin"], (Ch, ent["cin"])
            ups = ent["ups"]
            if ups:
                rows, cout = Lh + 1, ent["cout"]
                L_out = (Lh - 1) * ups - 2 * ent["shift"] + 2 * ups + ent["out_pad"]
                o = torch.empty(Bh, L_out, cout, dtype=adt, device=dev)
                shift = ent["shift"]
            else:
                rows, L_out, shift = Lh, 0, 0
                o = torch.empty(Bh, Lh, ent["n"], dtype=adt, device=dev)
            if "ktrim" in ent and not ups:
                N.check(lib.adb_cl_conv_ktrim(N.ptr(h), N.ptr(ent["w"]), N.ptr(ent["bias"]), N.ptr(res), N.ptr(o), Bh, Lh, rows, Ch,
                                              ent["n"], ent["taps"], ent["off0"], ent["dil"], act, ent["ktrim"], dt, st))
            else:
                N.check(lib.adb_cl_conv(N.ptr(h), N.ptr(ent["w"]), N.ptr(ent["bias"]), N.ptr(res), N.ptr(o), Bh, Lh, rows, Ch,
                                        ent["n"], ent["taps"], ent["off0"], ent["dil"], act, ups, shift, L_out, dt, st))
            return o

        def gn(h, gb, ss_ptr=None, ss_ld=0):
            o = torch.empty_like(h)
            N.check(lib.adb_cl_groupnorm(N.ptr(h), N.ptr(gb[0]), N.ptr(gb[1]), ss_ptr if ss_ptr is not None else c_void_p(0), ss_ld,
                                         N.ptr(o), N.ptr(sums), h.shape[0], h.shape[1], h.shape[2], groups, 1e-5, ACT_SILU, dt, st))
            return o

        def ln(h, g, b):
            o = torch.empty_like(h)
            N.check(lib.adb_cl_layernorm(N.ptr(h), N.ptr(g), N.ptr(b), N.ptr(o), h.shape[0] * h.shape[1], h.shape[2], 1e-5, dt, st))
            return o

        def stats_ok(c, g):
            # shapes the vectorised statistics kernel takes (adb_cl_gn_coef): 16-byte vectors inside a group, 256 threads = whole rows
            return g >= 1 and c % g == 0 and (c // g) % 8 == 0 and g <= 64 and 256 % g == 0 and c <= 2048 and 256 % (c // 8) == 0

        def fusable(c1, c2, ent1, ent2):
            # adb_cl_gn_conv3: bf16, k = 3 "same" convolutions, 64-channel K-blocks, groups that do not straddle the two inputs
            cin, co = c1 + c2, ent1["n"]
            if not (bf16 and self.fuse_groupnorm and c1 % 64 == 0 and c2 % 64 == 0 and co % 64 == 0 and cin % groups == 0
                    and c1 % (cin // groups) == 0):
                return False
            g1 = c1 // (cin // groups)
            return (stats_ok(c1, g1) and (c2 == 0 or stats_ok(c2, groups - g1)) and stats_ok(co, groups)
                    and all(e["taps"] == 3 and e["off0"] == -1 and e["dil"] == 1 and not e["ups"] and "f" not in e and "w16" in e
                            for e in (ent1, ent2)))

        def gn_conv(h, sk, gb, ss_ptr, ss_ld, ent, res):
            # statistics + per-channel coefficients of the raw input(s), then GroupNorm apply + SiLU inside the convolution's operand path
            Bh, Lh, c1 = h.shape
            c2 = sk.shape[2] if sk is not None else 0
            cin = c1 + c2
            g1 = c1 // (cin // groups)
            coef = torch.empty(2, Bh, cin, dtype=torch.float32, device=dev)
            ssp = ss_ptr if ss_ptr is not None else c_void_p(0)
            N.check(lib.adb_cl_gn_coef(N.ptr(h), N.ptr(sums), N.ptr(tickets), N.ptr(coef), Bh, Lh, c1, g1, groups, 0, 0, cin, N.ptr(gb[0]),
                                       N.ptr(gb[1]), ssp, ss_ld, 1e-5, 1.0, st))
            if sk is not None:
                N.check(lib.adb_cl_gn_coef(N.ptr(sk), N.ptr(sums), N.ptr(tickets), N.ptr(coef), Bh, Lh, c2, groups - g1, groups, g1, c1, cin,
                                           N.ptr(gb[0]), N.ptr(gb[1]), ssp, ss_ld, 1e-5, skip_scale, st))
            o = torch.empty(Bh, Lh, ent["n"], dtype=adt, device=dev)
            N.check(lib.adb_cl_gn_conv3(N.ptr(h), c1, N.ptr(sk), c2, N.ptr(coef), N.ptr(ent["w16"]), N.ptr(ent["bias"]), N.ptr(res), N.ptr(o),
                                        Bh, Lh, ent["n"], st))
            return o

        def resnet(p, h, sk=None):                               # ResnetBlock1d.forward, unet1d.py:297-316
            co = P[p + ".co"]
            ss_ptr = c_void_p(ss_all.data_ptr() + 4 * P["cond_off"][p])
            if fusable(h.shape[2], sk.shape[2] if sk is not None else 0, P[p + ".block1"], P[p + ".block2"]) and \
                    (sk is None or (p + ".to_out_cat") in P):
                if sk is not None:                               # residual branch on the never-materialised concatenation
                    ent = P[p + ".to_out_cat"]
                    r = torch.empty(h.shape[0], h.shape[1], ent["n"], dtype=adt, device=dev)
                    N.check(lib.adb_cl_conv_cat(N.ptr(h), h.shape[2], N.ptr(sk), sk.shape[2], N.ptr(ent["w"]), N.ptr(ent["bias"]),
                                                N.ptr(None), N.ptr(r), h.shape[0], h.shape[1], ent["n"], 1, 0, 1, ACT_NONE, st))
                else:
                    r = conv(h, P[p + ".to_out"]) if (p + ".to_out") in P else h
                a = gn_conv(h, sk, P[p + ".block1.gn"], None, 0, P[p + ".block1"], None)
                return gn_conv(a, None, P[p + ".block2.gn"], ss_ptr, ss_all.shape[1], P[p + ".block2"], r)
            if sk is not None:
                cat = torch.empty(h.shape[0], h.shape[1], h.shape[2] + sk.shape[2], dtype=adt, device=dev)
                N.check(lib.adb_cl_concat(N.ptr(h), N.ptr(sk), skip_scale, N.ptr(cat), h.shape[0] * h.shape[1], h.shape[2],
                                          sk.shape[2], dt, st))
                h = cat
            r = conv(h, P[p + ".to_out"]) if (p + ".to_out") in P else h
            a = conv(gn(h, P[p + ".block1.gn"]), P[p + ".block1"])
            assert a.shape[2] == co
            return conv(gn(a, P[p + ".block2.gn"], ss_ptr, ss_all.shape[1]), P[p + ".block2"], res=r)

        def transformer(p, h):                                   # TransformerBlock1d.forward, unet1d.py:106-122
            Bh, Lh, Ch = h.shape
            nrm = ln(h, *P[p + ".norm"])
            q, kv = conv(nrm, P[p + ".to_q"]), conv(nrm, P[p + ".to_kv"])
            a = torch.empty_like(q)
            N.check(lib.adb_cl_attention(N.ptr(q), N.ptr(kv), N.ptr(a), Bh, Lh, Ch, heads, dt, st))
            h = conv(a, P[p + ".to_out"], res=h)
            f = conv(ln(h, P[p + ".ff0"], None), P[p + ".ff1"], act=ACT_GELU)
            return conv(ln(f, P[p + ".ff3"], None), P[p + ".ff4"], res=h)

        # time embedding + every block's (scale, shift) projection in one batched linear (unet1d.py:678-684, :304-310)
        half = P["to_time.0.0.weights"].numel()
        T = P["to_time.2.bias"].numel()
        feat = torch.empty(B, 2 * half + 1, dtype=torch.float32, device=dev)
        N.check(lib.adb_cl_time_features(N.ptr(t), N.ptr(P["to_time.0.0.weights"]), N.ptr(feat), B, half, st))
        e1 = torch.empty(B, T, dtype=torch.float32, device=dev)
        N.check(lib.adb_cl_linear(N.ptr(feat), N.ptr(P["to_time.0.1.weight"]), N.ptr(P["to_time.0.1.bias"]), N.ptr(e1), B,
                                  2 * half + 1, T, 0, ACT_SILU, st))
        temb = torch.empty(B, T, dtype=torch.float32, device=dev)
        N.check(lib.adb_cl_linear(N.ptr(e1), N.ptr(P["to_time.2.weight"]), N.ptr(P["to_time.2.bias"]), N.ptr(temb), B, T, T, 0,
                                  ACT_NONE, st))
        cond = temb
        if classes is not None:                                   # LabelEmbedder.forward, conditioner.py:94-111 + unet1d.py:304-306
            lc, ch = P["lc"], cfg["channels"]
            cd = lc["class_to_cond.3.bias"].numel()
            e0 = torch.empty(B, ch, dtype=torch.float32, device=dev)
            N.check(lib.adb_cl_label_embed(N.ptr(lc["label_emb.weight"]), N.ptr(lc["null_classes_emb"]), N.ptr(classes), N.ptr(drop),
                                           N.ptr(e0), B, ch, cfg["num_classes"], st))
            e1 = torch.empty_like(e0)
            N.check(lib.adb_cl_layernorm(N.ptr(e0), N.ptr(lc["class_to_cond.0.weight"]), N.ptr(lc["class_to_cond.0.bias"]), N.ptr(e1), B,
                                         ch, 1e-5, 0, st))
            e2 = torch.empty(B, cd, dtype=torch.float32, device=dev)
            N.check(lib.adb_cl_linear(N.ptr(e1), N.ptr(lc["class_to_cond.1.weight"]), N.ptr(lc["class_to_cond.1.bias"]), N.ptr(e2), B, ch,
                                      cd, 0, ACT_SILU, st))
            e3 = torch.empty(B, cd, dtype=torch.float32, device=dev)
            N.check(lib.adb_cl_linear(N.ptr(e2), N.ptr(lc["class_to_cond.3.weight"]), N.ptr(lc["class_to_cond.3.bias"]), N.ptr(e3), B, cd,
                                      cd, 0, ACT_NONE, st))
            cond = torch.empty(B, T + cd, dtype=torch.float32, device=dev)
            N.check(lib.adb_cl_concat(N.ptr(temb), N.ptr(e3), 1.0, N.ptr(cond), B, T, cd, 0, st))
        ss_all = torch.empty(B, P["cond_w"].shape[0], dtype=torch.float32, device=dev)
        if "cond_tc" in P:
            ent, Tc, Nc = P["cond_tc"], cond.shape[1], ss_all.shape[1]
            cond_b16 = torch.empty(B, Tc, dtype=torch.bfloat16, device=dev)
            N.check(lib.adb_cl_cast(N.ptr(cond), N.ptr(cond_b16), B * Tc, 1, 1, st))
            ss_b16 = torch.empty(B, Nc, dtype=torch.bfloat16, device=dev)
            N.check(lib.adb_cl_conv(N.ptr(cond_b16), N.ptr(ent["w"]), N.ptr(ent["bias"]), N.ptr(None), N.ptr(ss_b16), 1, B, B, Tc, Nc,
                                    1, 0, 1, ACT_NONE, 0, 0, 0, dt, st))
            N.check(lib.adb_cl_cast(N.ptr(ss_b16), N.ptr(ss_all), B * Nc, 0, 0, st))
        else:
            N.check(lib.adb_cl_linear(N.ptr(cond), N.ptr(P["cond_w"]), N.ptr(P["cond_b"]), N.ptr(ss_all), B, cond.shape[1],
                                      ss_all.shape[1], 1, ACT_NONE, st))

        # input transform (WAVenc1d, unet1d.py:572-594)
        nf = cfg["num_filters"]
        pad = W // 2 - S // 2
        Lc = (L + 2 * pad - W) // S + 1
        if "to_in.tc" in P and L % W == 0:
            ent, rows = P["to_in.tc"], L // W
            xb = torch.empty(B, rows + 1, W * cin, dtype=adt, device=dev)
            N.check(lib.adb_cl_wavenc_prep(N.ptr(x), N.ptr(xb), B, cin, L, W, S, N.ptr(in_scale), st))
            h = torch.empty(B, rows, 2 * nf, dtype=adt, device=dev)
            N.check(lib.adb_cl_conv(N.ptr(xb), N.ptr(ent["w"]), N.ptr(None), N.ptr(None), N.ptr(h), B, rows + 1, rows, W * cin, 2 * nf,
                                    2, 0, 1, ACT_NONE, 0, 0, 0, dt, st))
            h = h.view(B, Lc, nf)                                # row m holds frames 2m and 2m+1
        else:
            if in_scale is not None:
                raise N.AdbError("internal: in_scale needs the tensor-core WAVenc path")
            h = torch.empty(B, Lc, nf, dtype=adt, device=dev)
            N.check(lib.adb_cl_wavenc(N.ptr(x), N.ptr(P["to_in.to_in.weight"]), N.ptr(h), B, cin, L, nf, W, S, dt, st))

        n = len(cfg["multipliers"]) - 1
        skip_scale = 2 ** -0.5 if cfg["use_skip_scale"] else 1.0
        skips_list = []
        for i in range(n):                                       # DownsampleBlock1d.forward, unet1d.py:432-457
            p = f"downsamples.{i}"
            h = conv(h, P[p + ".downsample"])
            skips = []
            for j in range(cfg["num_blocks"][i]):
                h = resnet(f"{p}.blocks.{j}", h)
                skips.append(h)
            if cfg["attentions"][i]:
                h = transformer(p + ".transformer", h)
                skips.append(h)
            skips_list.append(skips)
        h = resnet("bottleneck.pre_block", h)                    # BottleneckBlock1d.forward, unet1d.py:368-383
        if cfg["use_attention_bottleneck"]:
            h = transformer("bottleneck.transformer", h)
        h = resnet("bottleneck.post_block", h)
        for u, i in enumerate(reversed(range(n))):               # UpsampleBlock1d.forward, unet1d.py:539-566
            p = f"upsamples.{u}"
            skips = skips_list.pop()
            for j in range(cfg["num_blocks"][i] + (1 if cfg["attentions"][i] else 0)):
                h = resnet(f"{p}.blocks.{j}", h, skips.pop())     # cat(h, skip * skip_scale), unet1d.py:552-556
            if cfg["attentions"][i]:
                h = transformer(p + ".transformer", h)
            h = conv(h, P[p + ".upsample"])
        # output transform (WAVdec1d, unet1d.py:596-622)
        cout = cfg.get("out_channels") or cin
        if out.shape != (B, cout, (h.shape[1] - 1) * S - 2 * pad + W):
            raise N.AdbError(f"output buffer shape {tuple(out.shape)} does not match the network output")
        if "to_out.packed" in P:
            N.check(lib.adb_cl_wavdec_tc(N.ptr(h), N.ptr(P["to_out.packed"]), N.ptr(out), B, h.shape[1], nf, cout, W, S, st))
        else:
            N.check(lib.adb_cl_wavdec(N.ptr(h), N.ptr(P["to_out.to_out.weight"]), N.ptr(out), B, h.shape[1], nf, cout, W, S, dt, st))
        return out

    @torch.no_grad()
    def forward(self, x: Tensor, t: Tensor, classes: Optional[Tensor] = None, text_embeds: Optional[Tensor] = None,
                text_mask: Optional[Tensor] = None, inj_embeddings: Optional[Tensor] = None,
                inj_channels: Optional[Tensor] = None, cond_drop_prob=None, **kwargs) -> Tensor:
        """(x [B, in_channels, L], t [B]) -> [B, out_channels, L]   (unet1d.py:856-893)."""
        if any(v is not None for v in (text_embeds, inj_embeddings, inj_channels)):
            raise NotImplementedError("text / injection conditioning inputs are outside the fused UNet1d")
        if classes is not None and not self.cfg["class_cond"]:
            raise NotImplementedError("this UNet1d was built without class conditioning (class_cond=False)")
        x = N.require_cuda_f32(x, "x")
        if x.ndim != 3 or x.shape[1] != self.cfg["in_channels"]:
            raise N.AdbError(f"UNet1d expects x [B, {self.cfg['in_channels']}, L]; got {tuple(x.shape)}")
        B, _, L = x.shape
        t = N.require_cuda_f32(t, "t").reshape(B)
        drop = None
        if classes is not None:
            p = self.cond_drop_prob if cond_drop_prob is None else cond_drop_prob
            if isinstance(p, Tensor):                                # per-sample flags (used by the batched CFG pair)
                drop = p.to(device=x.device, dtype=torch.int32).reshape(B).contiguous()
            elif p <= 0:
                drop = torch.zeros(B, dtype=torch.int32, device=x.device)
            elif p >= 1:
                drop = torch.ones(B, dtype=torch.int32, device=x.device)
            else:
                raise NotImplementedError("random label dropout (0 < cond_drop_prob < 1, training) is not built on the fused path")
            classes = classes.to(device=x.device, dtype=torch.int64).reshape(B).contiguous()
        self._pack()
        cout = self.cfg.get("out_channels") or self.cfg["in_channels"]
        if not self.use_cuda_graph:
            return self._run(x, t, torch.empty(B, cout, L, dtype=torch.float32, device=x.device), classes, drop)
        # CUDA graph per (B, L, conditioned): ~400 small launches per evaluation would otherwise be host-bound
        key = (B, L, x.device.index, classes is not None)
        g = self._graphs.get(key)
        if g is None:
            sx, stt = torch.empty_like(x), torch.empty_like(t)
            so = torch.empty(B, cout, L, dtype=torch.float32, device=x.device)
            sc = torch.zeros(B, dtype=torch.int64, device=x.device) if classes is not None else None
            sdp = torch.zeros(B, dtype=torch.int32, device=x.device) if classes is not None else None
            sx.copy_(x)
            stt.copy_(t)
            if sc is not None:
                sc.copy_(classes)
                sdp.copy_(drop)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                self._run(sx, stt, so, sc, sdp)                 # warm-up outside capture (tensor-map cache, attributes)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            before = N.lib().adb_launch_count(0)
            with torch.cuda.graph(graph):
                self._run(sx, stt, so, sc, sdp)
            per_replay = N.lib().adb_launch_count(0) - before
            g = self._graphs[key] = (graph, sx, stt, so, per_replay, sc, sdp)
        graph, sx, stt, so, per_replay, sc, sdp = g
        sx.copy_(x)
        stt.copy_(t)
        if sc is not None:
            sc.copy_(classes)
            sdp.copy_(drop)
        graph.replay()
        self.graph_launches += per_replay
        return so.clone()

    # ---- device-resident EDM trajectory (EDMSampler.forward, sampler_edm.py:371-397) ---------------------------------------
    def _trajectory_graphs(self, B: int, L: int, dev, sigma_data: float):
        """Static state buffers + two CUDA graphs of one network evaluation each, F = net(c_in(sigma) x, c_noise(sigma)):
        graph 0 reads the state x, graph 1 the Heun midpoint x1; sigma lives in a 1-element device buffer that the loop
        rewrites before each replay. Nothing is copied in or out per evaluation (the plain forward() replay copies x in
        and clones F out: two 2 x 268 MB passes per evaluation at config 4)."""
        key = ("traj", B, L, dev.index, float(sigma_data))
        g = self._graphs.get(key)
        if g is not None:
            return g
        cin = self.cfg["in_channels"]
        lib = N.lib()
        P = self._pack()
        fold_scale = "to_in.tc" in P and L % self.cfg["window_length"] == 0      # c_in folded into the WAVenc re-layout
        buf = {k: torch.empty(B, cin, L, dtype=torch.float32, device=dev) for k in ("x", "x1", "d", "F") + (() if fold_scale else ("net_in",))}
        buf["c_noise"] = torch.empty(B, dtype=torch.float32, device=dev)
        buf["sigma"] = torch.ones(1, dtype=torch.float32, device=dev)
        buf["one"] = torch.ones(1, dtype=torch.float32, device=dev)
        buf["x"].zero_()
        buf["x1"].zero_()

        buf["c_in"] = torch.empty(B, dtype=torch.float32, device=dev)

        def evaluate(src):
            if fold_scale:
                # sigma is one value for the whole batch: broadcast it through a stride-0 read into per-sample c_in / c_noise
                N.check(lib.adb_edm_precond_coef(N.ptr(buf["sigma"]), 0, float(sigma_data), N.ptr(buf["c_in"]), N.ptr(buf["c_noise"]), B,
                                                 N.stream_ptr(dev)))
                self._run(src, buf["c_noise"], buf["F"], in_scale=buf["c_in"])
            else:
                N.check(lib.adb_edm_precond_in(N.ptr(src), N.ptr(buf["sigma"]), 0, float(sigma_data), N.ptr(buf["net_in"]),
                                               N.ptr(buf["c_noise"]), B, cin * L, N.stream_ptr(dev)))
                self._run(buf["net_in"], buf["c_noise"], buf["F"])

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            evaluate(buf["x"])                                   # warm-up outside capture (tensor-map cache, attributes)
        torch.cuda.current_stream(dev).wait_stream(side)
        graphs, per_replay = [], 0
        for src in ("x", "x1"):
            graph = torch.cuda.CUDAGraph()
            before = lib.adb_launch_count(0)
            with torch.cuda.graph(graph, pool=graphs[0].pool() if graphs else None):
                evaluate(buf[src])
            per_replay = lib.adb_launch_count(0) - before
            graphs.append(graph)
        g = self._graphs[key] = (buf, graphs, per_replay)
        return g

    @torch.no_grad()
    def _adb_fused_sample(self, noise, sig, num_steps, sigma_data, s_tmin, s_tmax, s_churn, s_noise, use_heun, alpha, eps,
                          churn_seed=0, sample_offset=0):
        """EDM Heun / Euler + churn trajectory for the unconditional U-Net with the state resident in static buffers: per step
        [churn kernel] -> graph replay -> fused mid kernel -> graph replay -> fused post kernel (48 B of HBM traffic per state
        element and Heun step around the network). Returns None when this mode does not apply (the caller then runs the generic
        Python loop): class-conditioned nets, the alpha sampler, eager mode, differing in/out channel counts."""
        cfg = self.cfg
        if alpha > 0 or cfg["class_cond"] or not self.use_cuda_graph or (cfg.get("out_channels") or cfg["in_channels"]) != cfg["in_channels"]:
            return None
        from math import sqrt
        from ..components.sampler_edm import _f32
        B, cin, L = noise.shape
        dev = noise.device
        self._pack()
        buf, graphs, per_replay = self._trajectory_graphs(B, L, dev, sigma_data)
        lib, st, n = N.lib(), N.stream_ptr(dev), noise.numel()
        x, x1, d, F = buf["x"], buf["x1"], buf["d"], buf["F"]

        def evaluate(which, sigma):
            N.check(lib.adb_edm_scale(N.ptr(buf["one"]), float(sigma), N.ptr(buf["sigma"]), 1, st))
            graphs[which].replay()
            self.graph_launches += per_replay

        N.check(lib.adb_edm_scale(N.ptr(noise), sig[0], N.ptr(x), n, st))
        gamma_on = _f32(min(s_churn / num_steps, sqrt(2) - 1))
        nfe = 0
        for i in range(num_steps):
            sigma = _f32(sig[i])
            sigma_next = _f32(sig[i + 1]) if i + 1 < len(sig) else 0.0
            gamma = gamma_on if (s_tmin <= sigma <= s_tmax) else 0.0
            sigma_hat = sigma
            if gamma > 0:
                sigma_hat = _f32(sigma + _f32(gamma * sigma))
                a = _f32(sqrt(_f32(_f32(sigma_hat * sigma_hat) - _f32(sigma * sigma))))
                if eps is not None:
                    tmp = torch.empty_like(x)
                    N.check(lib.adb_edm_scale(N.ptr(eps[i]), float(s_noise), N.ptr(tmp), n, st))
                    N.check(lib.adb_edm_axpy(N.ptr(x), N.ptr(tmp), a, N.ptr(x), n, st))
                else:
                    N.check(lib.adb_edm_churn_rng(N.ptr(x), N.ptr(x), a, float(s_noise), int(churn_seed), i, int(sample_offset), B,
                                                  cin * L, st))
            evaluate(0, sigma_hat)
            nfe += 1
            h = _f32(sigma_next - sigma_hat)
            if sigma_next != 0 and use_heun:
                N.check(lib.adb_edm_heun_mid(N.ptr(x), N.ptr(F), sigma_hat, float(sigma_data), h, N.ptr(d), N.ptr(x1), n, st))
                evaluate(1, sigma_next)
                nfe += 1
                N.check(lib.adb_edm_heun_post(N.ptr(x), N.ptr(d), N.ptr(F), sigma_next, float(sigma_data), h, N.ptr(x), n, st))
            else:
                N.check(lib.adb_edm_euler_raw(N.ptr(x), N.ptr(F), sigma_hat, float(sigma_data), h, N.ptr(x), n, st))
        return x.clone(), nfe

    @torch.no_grad()
    def _adb_cfg_pair(self, x: Tensor, t: Tensor, classes: Tensor):
        """Classifier-free guidance pair in ONE network evaluation of batch 2B (diffusion.py:50-53 calls the net twice):
        returns (F(x | classes), F(x | null))."""
        B = x.shape[0]
        flags = torch.cat([torch.zeros(B, dtype=torch.int32, device=x.device), torch.ones(B, dtype=torch.int32, device=x.device)])
        out = self.forward(torch.cat([x, x]), torch.cat([t, t]), classes=torch.cat([classes, classes]), cond_drop_prob=flags)
        return out[:B], out[B:]
