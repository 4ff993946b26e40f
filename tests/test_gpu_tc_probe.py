"""Pins the tcgen05 operand conventions of the tensor-core path on real hardware: one-tile BF16 GEMM
through the debug entry point opz_debug_umma (tests/libopz_probe.so, built from tests/csrc/opz_tc_probe.cu: test
infrastructure, not part of libopz.so) against a float64 product of the same BF16 operands."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)


def bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("N,K,mode", [(80, 96, 0), (208, 96, 0), (32, 400, 0), (80, 400, 0), (256, 64, 0), (16, 16, 0),
                                      (80, 96, 1), (32, 400, 1), (208, 96, 1)])
def test_umma_tile(opz, N, K, mode):
    opz.default_context()
    rng = np.random.default_rng(N * 1000 + K + mode)
    A = bf16_bits(rng.normal(size=(128, K)))
    B = bf16_bits(rng.normal(size=(N, K)))
    D = np.zeros((128, N), dtype=np.float32)
    import os
    probe = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libopz_probe.so"))
    fn = probe.opz_debug_umma
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = C.c_int
    opz.check(fn(N, K, mode, A.ctypes.data, B.ctypes.data, D.ctypes.data))
    ref = bits_to_f32(A).astype(np.float64) @ bits_to_f32(B).astype(np.float64).T
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 1e-5, f"N={N} K={K} mode={mode}: rel err {err}"
