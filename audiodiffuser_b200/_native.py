"""ctypes binding of the C-ABI library `libadb200.so` (include/adb200.h).

There is no CPU or PyTorch fallback: if the library has not been built, or the device is not a
B200 (sm_100), every call raises. PyTorch only provides device memory (`tensor.data_ptr()`) and the
current CUDA stream.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# tools/ may select the -DADB_DEBUG build (cycle accounting) with ADB_LIB=debug; everything else loads the product library
LIB_PATH = os.path.join(_HERE, "libadb200_dbg.so" if os.environ.get("ADB_LIB") == "debug" else "libadb200.so")

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISIONS = {"fp32": PRECISION_FP32, "float32": PRECISION_FP32, "bf16": PRECISION_BF16, "bfloat16": PRECISION_BF16}
TIMER_NAMES = ("conv", "step", "aux", "tail", "skip")

_lib = None

_F = POINTER(c_float)
_SIGS = {
    "adb_last_error": (c_char_p, []),
    "adb_version": (c_int, []),
    "adb_device_check": (c_int, [c_int]),
    "adb_check_async": (c_int, []),
    "adb_debug_zs_job_order": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int]),
    "adb_debug_ml_order": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(ctypes.c_longlong), POINTER(c_int), POINTER(c_int)]),
    "adb_launch_count": (ctypes.c_longlong, [c_int]),
    "adb_edm_precond_in": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "adb_edm_precond_out": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_float, c_void_p, c_int,
                                    c_int64, c_void_p]),
    "adb_edm_scale": (c_int, [c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "adb_edm_axpy": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "adb_edm_churn_rng": (c_int, [c_void_p, c_void_p, c_float, c_float, c_uint64, c_int, c_int64, c_int, c_int64, c_void_p]),
    "adb_edm_euler": (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    "adb_edm_rk2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p,
                            c_int64, c_void_p]),
    "adb_edm_lincomb": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_int64, c_void_p]),
    "adb_edm_clamp": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "adb_ema_lerp": (c_int, [c_void_p, c_void_p, c_float, c_int64, c_void_p]),
    "adb_edm_lincomb_n": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "adb_pcm16_encode": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "adb_edm_heun_mid": (c_int, [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    "adb_edm_heun_post": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_int64, c_void_p]),
    "adb_edm_euler_raw": (c_int, [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_int64, c_void_p]),
    "adb_edm_noise_in": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                 c_void_p]),
    "adb_edm_dsm_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int64, c_void_p]),
    "adb_edm_dsm_loss_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "adb_edm_dsm_loss_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                      c_void_p]),
    "adb_wavenet_param_count": (c_int64, [c_int, c_int]),
    "adb_wavenet_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_int64, c_int]),
    "adb_wavenet_load_params": (c_int, [c_void_p, c_void_p, c_int64, c_int]),
    "adb_wavenet_destroy": (None, [c_void_p]),
    "adb_wavenet_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int]),
    "adb_wavenet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_int64, c_void_p]),
    "adb_wavenet_forward_debug": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "adb_wavenet_denoise": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_int64, c_void_p]),
    "adb_wavenet_sample_edm": (c_int, [c_void_p, c_void_p, _F, c_int, c_int, c_float, c_float, c_float, c_float,
                                       c_float, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_int64, POINTER(c_int), c_void_p]),
    "adb_wavenet_sample_edm_seeded": (c_int, [c_void_p, c_void_p, _F, c_int, c_int, c_float, c_float, c_float, c_float,
                                              c_float, c_int, c_float, c_void_p, c_uint64, c_int64, c_void_p, c_int, c_int, c_int,
                                              c_void_p, c_int64, POINTER(c_int), c_void_p]),
    "adb_cl_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_conv_ktrim": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_conv_packed_elems": (c_int64, [c_int, c_int, c_int]),
    "adb_cl_pack_conv_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "adb_cl_pack_conv_weights_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "adb_cl_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "adb_cl_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_cast": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "adb_cl_label_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "adb_cl_time_features": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "adb_cl_groupnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_float, c_int, c_int, c_void_p]),
    "adb_cl_gn_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_gn_coef": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "adb_cl_gn_conv3": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p]),
    "adb_cl_conv_cat": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_void_p]),
    "adb_cl_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "adb_cl_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_concat": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "adb_cl_wavenc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_wavenc_prep": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "adb_edm_precond_coef": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "adb_cl_wavdec": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_wavenet_train_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int]),
    "adb_wavenet_dsm_forward_train": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int,
                                              c_void_p, c_int64, c_void_p]),
    "adb_wavenet_dsm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_int64, c_void_p]),
    "adb_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_float,
                               c_int, c_float, c_void_p]),
    "adb_cl_wavdec_packed_elems": (c_int64, [c_int]),
    "adb_cl_wavdec_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_cl_wavdec_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "adb_wavenet_set_timing": (c_int, [c_void_p, c_int]),
    "adb_wavenet_timers": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64)]),
}


class AdbError(RuntimeError):
    pass


def lib():
    """Load libadb200.so (once). Raises loudly if it is missing — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AdbError(
                f"{LIB_PATH} not found: build the sm_100a CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()' or python -m audiodiffuser_b200.build)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise AdbError(f"adb200 error {rc}: {lib().adb_last_error().decode()}")


def check_async():
    """Synchronise and surface asynchronous kernel / pipeline errors."""
    check(lib().adb_check_async())


def stream_ptr(device=None):
    """Current stream of `device` as a void*. The kernels launch on the CUDA device that is current in this thread, so the
    tensors' device must be that device (one process per GPU is the intended deployment): a mismatch fails here, loudly,
    instead of launching on the wrong GPU."""
    if device is not None:
        idx = torch.device(device).index
        cur = torch.cuda.current_device()
        if idx is not None and idx != cur:
            raise AdbError(f"tensors live on cuda:{idx} but the current CUDA device is cuda:{cur}: wrap the call in "
                           f"`with torch.cuda.device({idx}):` (or call torch.cuda.set_device) — adb200 launches on the current device")
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise AdbError(f"{name} is on {t.device}: adb200 has no CPU path; move it to a B200")
    if t.dtype != torch.float32:
        raise AdbError(f"{name} must be float32 (state and preconditioning are fp32, SURVEY.md §8(b)); got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def alloc_workspace(nbytes, device):
    """uint8 scratch tensor plus a 1024-byte aligned pointer into it (TMA tiles need the alignment; the caching
    allocator only guarantees 512)."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    base = (buf.data_ptr() + 1023) // 1024 * 1024
    return buf, c_void_p(base)


_device_ok = set()


def ensure_device(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        check(lib().adb_device_check(idx))
        _device_ok.add(idx)
    return idx
