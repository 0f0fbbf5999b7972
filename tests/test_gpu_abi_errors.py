"""The C ABI's error behaviour (include/adb200.h: non-zero code + adb_last_error, never a silent fallback), called
directly through ctypes like a foreign host would."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from audiodiffuser_b200 import _native as N
    assert torch.cuda.is_available()
    return N, N.lib(), torch.device("cuda:0")


def test_null_and_shape_errors_return_codes(env):
    N, lib, dev = env
    x = torch.zeros(16, device=dev)
    st = N.stream_ptr(dev)
    assert lib.adb_edm_scale(ctypes.c_void_p(0), 1.0, N.ptr(x), 16, st) == 1                 # ADB_ERR_INVALID
    assert b"adb_edm_scale" in lib.adb_last_error()
    assert lib.adb_edm_precond_in(N.ptr(x), N.ptr(x), 2, 0.2, N.ptr(x), ctypes.c_void_p(0), 1, 16, st) == 1   # bad sigma_stride
    assert lib.adb_cl_conv(N.ptr(x), N.ptr(x), ctypes.c_void_p(0), ctypes.c_void_p(0), N.ptr(x), 1, 4, 4, 48, 64, 1, 0, 1, 0, 0, 0, 0,
                           1, st) == 1                                                          # Cin % 64 != 0 on the tensor-core path
    assert b"Cin" in lib.adb_last_error()
    assert lib.adb_cl_groupnorm(N.ptr(x), N.ptr(x), N.ptr(x), ctypes.c_void_p(0), 0, N.ptr(x), N.ptr(x), 1, 2, 10, 4, 1e-5, 0, 0,
                                st) == 1                                                        # C % G != 0
    assert lib.adb_cl_attention(N.ptr(x), N.ptr(x), N.ptr(x), 1, 100000, 64, 1, 0, st) == 3    # ADB_ERR_UNSUPPORTED: too many keys
    N.check_async()


def test_wavenet_handle_errors(env):
    N, lib, dev = env
    h = ctypes.c_void_p()
    flat = torch.zeros(10, device=dev)
    assert lib.adb_wavenet_create(ctypes.byref(h), 256, 2, 12, N.ptr(flat), 10, 1) == 1          # wrong parameter count
    assert b"parameter vector" in lib.adb_last_error()
    assert lib.adb_wavenet_create(ctypes.byref(h), 100, 2, 12, N.ptr(flat), 10, 1) == 1          # channels not a multiple of 64
    n = lib.adb_wavenet_param_count(64, 2)
    flat = torch.randn(n, device=dev) * 0.05
    assert lib.adb_wavenet_create(ctypes.byref(h), 64, 2, 2, N.ptr(flat), n, 1) == 0
    try:
        B, L = 1, 100
        x, t, out = torch.zeros(B, L, device=dev), torch.zeros(B, device=dev), torch.zeros(B, L, device=dev)
        need = lib.adb_wavenet_workspace_bytes(h, B, L, 0)
        buf, ws = N.alloc_workspace(need, dev)
        st = N.stream_ptr(dev)
        # workspace too small / misaligned / bf16 path on a width it is not built for: refused, nothing runs
        assert lib.adb_wavenet_forward(h, N.ptr(x), N.ptr(t), ctypes.c_void_p(0), 0, N.ptr(out), B, L, 0, ws, need - 1, st) == 1
        assert lib.adb_wavenet_forward(h, N.ptr(x), N.ptr(t), ctypes.c_void_p(0), 0, N.ptr(out), B, L, 0, ctypes.c_void_p(ws.value + 8),
                                       need, st) == 1
        assert lib.adb_wavenet_forward(h, N.ptr(x), N.ptr(t), ctypes.c_void_p(0), 0, N.ptr(out), B, L, 1, ws, need, st) == 3
        assert b"256" in lib.adb_last_error()
        assert lib.adb_wavenet_forward(h, N.ptr(x), N.ptr(t), ctypes.c_void_p(0), 0, N.ptr(out), B, L, 0, ws, need, st) == 0
        assert lib.adb_wavenet_load_params(h, N.ptr(flat), n - 1, 1) == 1
        N.check_async()
    finally:
        lib.adb_wavenet_destroy(h)


def test_launch_counter(env):
    N, lib, dev = env
    x = torch.ones(64, device=dev)
    lib.adb_launch_count(1)
    N.check(lib.adb_edm_scale(N.ptr(x), 2.0, N.ptr(x), 64, N.stream_ptr(dev)))
    N.check(lib.adb_edm_clamp(N.ptr(x), N.ptr(x), 64, N.stream_ptr(dev)))
    assert lib.adb_launch_count(0) == 2
    assert float(x.max()) == 1.0
