"""CPU ORACLE for the AQLM quantized-linear hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This module is a plain-numpy restatement of the reference's own definition of "correct" for
`aqlm.QuantizedLinear.forward`.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
`aqlm_b200` never does (it fails loudly when the CUDA extension is missing).

Parity pinning: the reference ships no golden vectors or known-answer tests for this path
(SURVEY.md §4, §8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the
build container: `tests/golden/make_golden.py` imports `/root/reference/inference_lib/src/aqlm`,
runs `dequantize_gemm`, `QuantizedLinear.forward` (CPU) and the Numba LUT kernel on seeded inputs and
commits the results under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
here against them.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "get_int_dtype",
    "pack_int_data",
    "unpack_int_data",
    "dequantize_weight",
    "dequantize_gemm",
    "lut_gemv",
    "relative_error",
    "code_bytes",
    "make_case",
]


def get_int_dtype(nbits: int):
    """inference_lib/src/aqlm/utils.py:11-20 -- smallest signed int type holding `nbits`."""
    if nbits <= 8:
        return np.int8
    if nbits <= 16:
        return np.int16
    if nbits <= 32:
        return np.int32
    if nbits <= 64:
        return np.int64
    raise ValueError(f"No dtype available for {nbits}-bit codebooks")


def pack_int_data(data: np.ndarray, nbits: int) -> np.ndarray:
    """inference_lib/src/aqlm/utils.py:23-26 -- values >= 2^(nbits-1) wrap to negative, then cast.

    (The reference mutates its argument in place; the oracle works on a copy.)
    """
    data = np.array(data, dtype=np.int64, copy=True)
    data[data >= 2 ** (nbits - 1)] -= 2**nbits
    return data.astype(get_int_dtype(nbits))


def unpack_int_data(data: np.ndarray, nbits: int) -> np.ndarray:
    """inference_lib/src/aqlm/utils.py:29-31 -- `.to(int64) % 2**nbits` (python modulo: non-negative)."""
    return np.asarray(data).astype(np.int64) % (2**nbits)


def dequantize_weight(codes: np.ndarray, codebooks: np.ndarray, scales: np.ndarray | None = None,
                      dtype=np.float32) -> np.ndarray:
    """inference_lib/src/aqlm/utils.py:43-70 (`_dequantize_weight`).

    codes      [num_out_groups, num_in_groups, num_codebooks]  UNSIGNED code values (already unpacked)
    codebooks  [num_codebooks, codebook_size, out_group_size, in_group_size]
    scales     broadcastable with [num_out_groups, num_in_groups, out_group_size, in_group_size]
    returns    [out_features, in_features] = scales * sum_c codebooks[c, codes[..., c]]

    The reference gathers with `F.embedding_bag(mode="sum")` (utils.py:60-62), views the result as
    [og_n, ig_n, og, ig] (64-66), multiplies by scales (67-68) and swaps axes -3,-2 (70).
    """
    codes = np.asarray(codes)
    num_out_groups, num_in_groups, num_codebooks = codes.shape
    nc, codebook_size, out_group_size, in_group_size = codebooks.shape
    assert nc == num_codebooks
    cb = np.asarray(codebooks, dtype=dtype)
    acc = np.zeros((num_out_groups, num_in_groups, out_group_size, in_group_size), dtype=dtype)
    for c in range(num_codebooks):  # embedding_bag(mode="sum") over the codebook axis
        acc += cb[c][codes[:, :, c]]
    if scales is not None:
        acc = acc * np.asarray(scales, dtype=dtype).reshape(num_out_groups, 1, 1, 1)
    return acc.swapaxes(-3, -2).reshape(num_out_groups * out_group_size, num_in_groups * in_group_size)


def dequantize_gemm(x: np.ndarray, codes: np.ndarray, codebooks: np.ndarray, scales: np.ndarray,
                    bias: np.ndarray | None, dtype=np.float32) -> np.ndarray:
    """inference_lib/src/aqlm/inference_kernels/dequantization.py:9-21 -- THE oracle.

    `codes` are the PACKED (signed-storage) codes as stored in the module; nbits is recovered as
    `codebooks.shape[1].bit_length() - 1` (dequantization.py:17).
    """
    nbits = int(codebooks.shape[1]).bit_length() - 1
    w = dequantize_weight(unpack_int_data(codes, nbits), codebooks, scales, dtype=dtype)
    y = np.asarray(x, dtype=dtype) @ w.T
    if bias is not None:
        y = y + np.asarray(bias, dtype=dtype)
    return y


def lut_gemv(x: np.ndarray, codes_alt: np.ndarray, codebooks: np.ndarray, scales: np.ndarray,
             dtype=np.float32) -> np.ndarray:
    """inference_lib/src/aqlm/inference_kernels/numba_kernel.py:37-48 (`numba_gemv_lut_`), race-free.

    x          [in_features]
    codes_alt  [num_in_groups, out_features, num_codebooks] uint8 view (the permuted CPU layout,
               inference.py:78-83)
    codebooks  [num_codebooks, 256, 1, in_group_size]
    lut[j, c, k] = x_j . codebooks[c, k]           (numba_kernel.py:39-40)
    y[i]       = scales[i] * sum_j sum_c lut[j, c, codes_alt[j, i, c]]   (42-47)
    """
    num_codebooks, codebook_size, out_group_size, in_group_size = codebooks.shape
    assert out_group_size == 1
    xg = np.asarray(x, dtype=dtype).reshape(-1, in_group_size)
    cb = np.asarray(codebooks, dtype=dtype).reshape(-1, in_group_size)
    lut = (xg @ cb.T).reshape(-1, num_codebooks, codebook_size)
    codes_alt = np.asarray(codes_alt).view(np.uint8) if codes_alt.dtype == np.int8 else np.asarray(codes_alt)
    num_in_groups, out_features, _ = codes_alt.shape
    y = np.zeros(out_features, dtype=dtype)
    jj = np.arange(num_in_groups)[:, None]
    for c in range(num_codebooks):
        y += lut[jj, c, codes_alt[:, :, c]].sum(axis=0, dtype=dtype)
    return y * np.asarray(scales, dtype=dtype).reshape(-1)


def relative_error(y: np.ndarray, y_ref: np.ndarray) -> float:
    """benchmark/matmul_benchmark.py:108 -- mean|y - y_ref| / mean|y_ref| (the reference's own metric)."""
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    return float(np.mean(np.abs(y - y_ref)) / np.mean(np.abs(y_ref)))


def code_bytes(out_features: int, in_features: int, num_codebooks: int, nbits: int, in_group_size: int = 8) -> int:
    """algorithmic bytes per matvec = out * (in/g) * K * ceil(nbits/8)."""
    return out_features * (in_features // in_group_size) * num_codebooks * ((nbits + 7) // 8)


def make_case(seed: int, in_features: int, out_features: int, num_codebooks: int, nbits: int,
              in_group_size: int = 8, batch: int = 1, bias: bool = False, float_dtype=np.float16):
    """Seeded synthetic inputs mirroring benchmark/matmul_benchmark.py:83-97 (randn x / codebooks,
    randint codes) with non-trivial scales (matmul_benchmark_cpu.py:123 uses randn scales; we use
    0.75 + 0.5*U so no output is scaled to ~0).  Values are rounded to `float_dtype` (fp16 by default)
    so that the oracle and the CUDA path see bit-identical inputs.

    Returns dict(x [batch,in], codes [out, in/g, K] packed ints, codebooks [K, 2^nbits, 1, g],
                 scales [out,1,1,1], bias [out] | None).
    """
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, in_features), dtype=np.float32).astype(float_dtype)
    raw = rng.integers(0, 2**nbits, size=(out_features, in_features // in_group_size, num_codebooks), dtype=np.int64)
    codes = pack_int_data(raw, nbits)
    codebooks = rng.standard_normal((num_codebooks, 2**nbits, 1, in_group_size), dtype=np.float32).astype(float_dtype)
    scales = (0.75 + 0.5 * rng.random((out_features, 1, 1, 1), dtype=np.float32)).astype(float_dtype)
    b = rng.standard_normal((out_features,), dtype=np.float32).astype(float_dtype) if bias else None
    return dict(x=x, codes=codes, codebooks=codebooks, scales=scales, bias=b)
