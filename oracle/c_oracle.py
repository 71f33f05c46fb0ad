"""ctypes binding of oracle/libaqlm_oracle.so (the C restatement of the CPU oracle).

TEST INFRASTRUCTURE, NOT PRODUCT -- see the header of aqlm_oracle.c.  Used by tests (cross-check of the
numpy oracle), and by bench.py's `cpu_baseline` / `--impl reference` legs as the timed CPU port.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaqlm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "aqlm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaqlm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        L.aqlm_oracle_num_threads.restype = i32
        L.aqlm_oracle_dequantize_weight.argtypes = [vp, i32, vp, vp, vp, i64, i64, i32, i32, i32, i32]
        L.aqlm_oracle_dequantize_gemm.argtypes = [vp, i64, vp, i32, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32]
        L.aqlm_oracle_lut_gemv.argtypes = [vp, vp, vp, vp, vp, i64, i64, i32, i32, i32]
        _lib = L
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().aqlm_oracle_num_threads())


def dequantize_weight(codes, codebooks, scales, nthreads: int = 0) -> np.ndarray:
    """codes: PACKED ints [out, in_groups, K]; codebooks [K, 2^nbits, 1, g]."""
    codes = np.ascontiguousarray(codes)
    K, cb_size, og, g = codebooks.shape
    assert og == 1
    nbits = int(cb_size).bit_length() - 1
    out, in_groups, _ = codes.shape
    cb, sc = _f32(codebooks), _f32(None if scales is None else np.asarray(scales).reshape(-1))
    W = np.empty((out, in_groups * g), dtype=np.float32)
    rc = lib().aqlm_oracle_dequantize_weight(_ptr(codes), codes.itemsize, _ptr(cb), _ptr(sc), _ptr(W), out, in_groups,
                                             K, nbits, g, nthreads)
    assert rc == 0, rc
    return W


def dequantize_gemm(x, codes, codebooks, scales, bias, nthreads: int = 0) -> np.ndarray:
    codes = np.ascontiguousarray(codes)
    K, cb_size, og, g = codebooks.shape
    assert og == 1
    nbits = int(cb_size).bit_length() - 1
    out, in_groups, _ = codes.shape
    x2 = _f32(x).reshape(-1, in_groups * g)
    cb, sc, b = _f32(codebooks), _f32(np.asarray(scales).reshape(-1)), _f32(bias)
    y = np.empty((x2.shape[0], out), dtype=np.float32)
    rc = lib().aqlm_oracle_dequantize_gemm(_ptr(x2), x2.shape[0], _ptr(codes), codes.itemsize, _ptr(cb), _ptr(sc),
                                           _ptr(b), _ptr(y), out, in_groups, K, nbits, g, nthreads)
    assert rc == 0, rc
    return y.reshape(tuple(np.asarray(x).shape[:-1]) + (out,))


def lut_gemv(x, codes_alt, codebooks, scales, nthreads: int = 0) -> np.ndarray:
    """codes_alt: uint8 [in_groups, out, K] (the permuted CPU layout)."""
    codes_alt = np.ascontiguousarray(np.asarray(codes_alt).view(np.uint8))
    K, cb_size, og, g = codebooks.shape
    assert og == 1 and cb_size == 256
    in_groups, out, _ = codes_alt.shape
    xx, cb, sc = _f32(x).reshape(-1), _f32(codebooks), _f32(np.asarray(scales).reshape(-1))
    y = np.empty(out, dtype=np.float32)
    rc = lib().aqlm_oracle_lut_gemv(_ptr(xx), _ptr(codes_alt), _ptr(cb), _ptr(sc), _ptr(y), out, in_groups, K, g,
                                    nthreads)
    assert rc == 0, rc
    return y
