"""Pin the CPU oracle (numpy + C restatements) against outputs of the reference itself.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py, which imports the real
`aqlm` package from /root/reference/inference_lib/src and runs `dequantize_gemm`
(dequantization.py:9-21), `QuantizedLinear.forward` on CPU (inference.py:68-75; Numba LUT kernel for
256-entry codebooks) and `_dequantize_weight` (utils.py:43-70) on seeded inputs.
"""
import numpy as np
import pytest
from conftest import case_from_meta, golden_cases

from oracle import aqlm_oracle as O
from oracle import c_oracle as C

CASES = golden_cases()
IDS = [c["name"] for c in CASES]


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_numpy_oracle_matches_reference_dequantize_gemm(c, golden):
    data, _ = golden
    case = case_from_meta(c)
    y = O.dequantize_gemm(case["x"], case["codes"], case["codebooks"], case["scales"], case["bias"])
    ref = data[f"{c['name']}/y_dequantize_gemm_fp32"]
    assert y.shape == ref.shape
    # same fp32 arithmetic up to summation order inside the dense matmul
    assert O.relative_error(y, ref) < 2e-6
    np.testing.assert_allclose(y, ref, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_c_oracle_matches_reference_dequantize_gemm(c, golden):
    data, _ = golden
    case = case_from_meta(c)
    for nthreads in (1, 4):
        y = C.dequantize_gemm(case["x"], case["codes"], case["codebooks"], case["scales"], case["bias"], nthreads)
        ref = data[f"{c['name']}/y_dequantize_gemm_fp32"]
        assert y.shape == ref.shape
        assert O.relative_error(y, ref) < 2e-6


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_oracle_matches_reference_module_forward(c, golden):
    """QuantizedLinear.forward on CPU fp32: dequantize_gemm for 1x16, Numba LUT kernel for Kx8."""
    data, _ = golden
    case = case_from_meta(c)
    ref = data[f"{c['name']}/y_module_cpu_fp32"]
    y = O.dequantize_gemm(case["x"], case["codes"], case["codebooks"], case["scales"], case["bias"])
    assert O.relative_error(y, ref) < 2e-6
    if c["module_path"] == "numba_gemm_lut":
        # restate the LUT algorithm itself on the permuted layout (inference.py:78-83)
        codes_alt = np.ascontiguousarray(np.transpose(case["codes"], (1, 0, 2))).view(np.uint8)
        for b in range(c["batch"]):
            for fn in (O.lut_gemv, lambda *a: C.lut_gemv(*a, nthreads=3)):
                yb = fn(case["x"][b], codes_alt, case["codebooks"], case["scales"])
                if case["bias"] is not None:
                    yb = yb + case["bias"].astype(np.float32)
                assert O.relative_error(yb, ref[b]) < 2e-6


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_fp16_reference_within_north_star_tolerance(c, golden):
    """The reference's own fp16 CPU forward is within 1e-3 of the fp32 oracle (the budget our fp16 CUDA path gets)."""
    data, _ = golden
    y16 = data[f"{c['name']}/y_dequantize_gemm_fp16"]
    y32 = data[f"{c['name']}/y_dequantize_gemm_fp32"]
    assert O.relative_error(y16, y32) < 1e-3


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_dequantize_weight_matches_reference(c, golden):
    data, _ = golden
    case = case_from_meta(c)
    W = O.dequantize_weight(O.unpack_int_data(case["codes"], c["nbits"]), case["codebooks"], case["scales"])
    Wc = C.dequantize_weight(case["codes"], case["codebooks"], case["scales"], nthreads=2)
    np.testing.assert_array_equal(W, Wc)
    key = f"{c['name']}/W_fp32"
    if key in data.files:
        np.testing.assert_array_equal(W, data[key])  # gather + sum + scale in the same order: bit-exact
    else:
        np.testing.assert_array_equal(W[:8], data[key + "_rows0_8"])
        np.testing.assert_allclose(W.sum(axis=1, dtype=np.float64), data[key + "_rowsum"], rtol=1e-12)


def test_general_out_group_size(golden):
    data, _ = golden
    W = O.dequantize_weight(data["general_og2/raw_codes"], data["general_og2/codebooks"], data["general_og2/scales"])
    np.testing.assert_array_equal(W, data["general_og2/W_fp32"])


@pytest.mark.parametrize("nbits", [1, 7, 8, 12, 16])
def test_pack_unpack_known_answers(nbits, golden):
    data, _ = golden
    vals = data[f"pack/nbits{nbits}_values"]
    packed = O.pack_int_data(vals, nbits)
    assert packed.dtype == data[f"pack/nbits{nbits}_packed"].dtype == O.get_int_dtype(nbits)
    np.testing.assert_array_equal(packed, data[f"pack/nbits{nbits}_packed"])
    np.testing.assert_array_equal(O.unpack_int_data(packed, nbits), data[f"pack/nbits{nbits}_unpacked"])
    np.testing.assert_array_equal(O.unpack_int_data(packed, nbits), vals)  # round trip


def test_int_dtype_table():
    assert [O.get_int_dtype(n) for n in (1, 8, 9, 16, 17, 32, 33, 64)] == [np.int8, np.int8, np.int16, np.int16,
                                                                           np.int32, np.int32, np.int64, np.int64]
    with pytest.raises(ValueError):
        O.get_int_dtype(65)


def test_code_bytes_formula():
    # SURVEY.md §8(d) concrete numbers
    assert O.code_bytes(4096, 4096, 1, 16) == 4_194_304
    assert O.code_bytes(14336, 4096, 1, 16) == 14_680_064
    assert O.code_bytes(11008, 4096, 2, 8) == 11_272_192
    assert O.code_bytes(11008, 4096, 8, 8) == 45_088_768
    assert O.code_bytes(28672, 8192, 1, 16) == 58_720_256
