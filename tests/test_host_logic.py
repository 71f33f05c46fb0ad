"""CPU-only tests of the host side: package surface, format utilities, selector table, C-ABI symbol export.

No compute runs here (there is no GPU and the product has no CPU fallback); compute parity lives in
test_gpu_parity.py (-m gpu) and the oracle is pinned in test_oracle_golden.py.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import aqlm_b200
from aqlm_b200 import _cabi
from aqlm_b200.utils import get_int_dtype, pack_int_data, unpack_int_data


def test_library_is_built_and_exports_every_header_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(_cabi.LIB_PATH)
    names = _cabi.header_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/aqlm_b200.h but not exported"
    L.aqlm_b200_version.restype = ctypes.c_int
    assert L.aqlm_b200_version() == 100


def test_weight_struct_matches_header_layout():
    # 4 pointers + 2 int64 + 6 int32 = 72 bytes, no padding surprises
    assert ctypes.sizeof(_cabi.Weight) == 4 * 8 + 2 * 8 + 6 * 4
    assert _cabi.Weight.in_features.offset == 32 and _cabi.Weight.num_codebooks.offset == 48


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    L = _cabi.lib()
    w = _cabi.Weight()
    w.codes, w.codebooks, w.scales = 16, 16, 16  # dummy non-null aligned pointers (never dereferenced)
    w.in_features, w.out_features = 64, 8
    w.num_codebooks, w.nbits_per_codebook, w.in_group_size, w.out_group_size = 1, 16, 8, 1
    w.dtype = 7
    assert L.aqlm_b200_matmat(ctypes.byref(w), 16, 16, 1, None) == _cabi.ERR_DTYPE
    with pytest.raises(NotImplementedError, match="float16 and bfloat16"):
        _cabi.check(_cabi.ERR_DTYPE)
    w.dtype = _cabi.F16
    w.in_group_size = 4
    assert L.aqlm_b200_matmat(ctypes.byref(w), 16, 16, 1, None) == _cabi.ERR_UNSUPPORTED
    assert b"8 or 16 features" in L.aqlm_b200_last_error()
    w.in_group_size = 8
    w.out_group_size = 2
    assert L.aqlm_b200_matmat(ctypes.byref(w), 16, 16, 1, None) == _cabi.ERR_UNSUPPORTED
    w.out_group_size = 1
    w.in_features = 63
    assert L.aqlm_b200_matmat(ctypes.byref(w), 16, 16, 1, None) == _cabi.ERR_SHAPE
    w.in_features = 64
    # valid descriptor but no CUDA device in this container: must FAIL LOUDLY, not fall back
    if not torch.cuda.is_available():
        rc = L.aqlm_b200_matmat(ctypes.byref(w), 16, 16, 1, None)
        assert rc in (_cabi.ERR_CUDA, _cabi.ERR_ARCH)
        with pytest.raises(RuntimeError):
            _cabi.check(rc)


def test_constructor_contract_matches_reference():
    """Parameter names / shapes / dtypes of inference.py:39-61, constructed on the meta device like HF does."""
    m = aqlm_b200.QuantizedLinear(in_features=4096, out_features=1024, in_group_size=8, out_group_size=1,
                                  num_codebooks=1, nbits_per_codebook=16, bias=False, device="meta",
                                  dtype=torch.float16)
    sd = m.state_dict()
    assert list(sd) == ["codebooks", "codes", "scales"]
    assert sd["codebooks"].shape == (1, 65536, 1, 8) and sd["codebooks"].dtype == torch.float16
    assert sd["codes"].shape == (1024, 512, 1) and sd["codes"].dtype == torch.int16
    assert sd["scales"].shape == (1024, 1, 1, 1)
    assert not any(p.requires_grad for p in m.parameters())
    m2 = aqlm_b200.QuantizedLinear(512, 64, 8, 1, 2, 8, bias=True, dtype=torch.float16)
    assert m2.codes.dtype == torch.int8 and m2.codes.shape == (64, 64, 2) and m2.bias.shape == (64,)
    assert m2.codebook_size == 256
    with pytest.raises(AssertionError):
        aqlm_b200.QuantizedLinear(100, 64, 8, 1, 1, 16)


def test_cpu_inputs_fail_loudly():
    m = aqlm_b200.QuantizedLinear(64, 16, 8, 1, 1, 16, bias=False, dtype=torch.float16)
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        m(torch.zeros(1, 64, dtype=torch.float16))
    from aqlm_b200.inference_kernels import get_forward_pass_kernel

    with pytest.raises(NotImplementedError):
        get_forward_pass_kernel(torch.zeros(1, 65536, 1, 8, dtype=torch.float16), False)
    with pytest.raises(NotImplementedError):
        torch.ops.aqlm.code1x16_matmat(torch.zeros(1, 64, dtype=torch.float16), m.codes, m.codebooks, m.scales, None)
    with pytest.raises(NotImplementedError):
        aqlm_b200.utils._dequantize_weight(torch.zeros(16, 8, 1, dtype=torch.int64), m.codebooks, None)


def test_selector_table_matches_reference_names():
    """kernel_selector.py:21-163 -- op chosen per (K, codebook_size, in_group) on a CUDA device (meta stand-in)."""
    from aqlm_b200.inference_kernels import kernel_selector as ks

    class FakeCB:  # only .shape and .device are inspected
        def __init__(self, shape):
            self.shape = shape
            self.device = torch.device("cuda", 0)

    expect = {(1, 65536, 1, 8): "code1x16", (1, 65536, 1, 16): "code1x16", (2, 256, 1, 8): "code2x8",
              (1, 256, 1, 8): "code1x8", (8, 256, 1, 8): "generic", (2, 4096, 1, 8): "generic"}
    for shape, prefix in expect.items():
        assert ks.get_forward_pass_kernel(FakeCB(shape), False) is getattr(torch.ops.aqlm, prefix + "_matmat")
        assert ks.get_forward_pass_kernel(FakeCB(shape), True) is getattr(torch.ops.aqlm, prefix + "_matmat_dequant")
        assert ks.get_backward_pass_kernel(FakeCB(shape), True) is getattr(torch.ops.aqlm,
                                                                         prefix + "_matmat_dequant_transposed")
    with pytest.raises(NotImplementedError):
        ks.get_forward_pass_kernel(FakeCB((1, 256, 2, 8)), False)
    with pytest.raises(NotImplementedError):
        ks.get_forward_pass_kernel(FakeCB((1, 256, 1, 4)), False)


def test_fake_tensor_shapes_for_compile():
    """register_fake (reference impl_abstract, cuda_kernel.py:20-22, 47-51): shapes on the meta device."""
    x = torch.empty(2, 3, 4096, device="meta", dtype=torch.float16)
    codes = torch.empty(1024, 512, 1, device="meta", dtype=torch.int16)
    cb = torch.empty(1, 65536, 1, 8, device="meta", dtype=torch.float16)
    sc = torch.empty(1024, 1, 1, 1, device="meta", dtype=torch.float16)
    assert torch.ops.aqlm.code1x16_matmat(x, codes, cb, sc, None).shape == (2, 3, 1024)
    assert torch.ops.aqlm.code1x16_matmat_dequant(x, codes, cb, sc, None).shape == (2, 3, 1024)
    g = torch.empty(2, 3, 1024, device="meta", dtype=torch.float16)
    assert torch.ops.aqlm.code1x16_matmat_dequant_transposed(g, codes, cb, sc, None).shape == (2, 3, 4096)


def test_pack_unpack_against_golden(golden):
    data, _ = golden
    for nbits in (1, 7, 8, 12, 16):
        vals = torch.from_numpy(data[f"pack/nbits{nbits}_values"])
        packed = pack_int_data(vals.clone(), nbits)
        assert packed.dtype == get_int_dtype(nbits)
        np.testing.assert_array_equal(packed.numpy(), data[f"pack/nbits{nbits}_packed"])
        np.testing.assert_array_equal(unpack_int_data(packed, nbits).numpy(), data[f"pack/nbits{nbits}_unpacked"])


def test_pack_int_data_mutates_in_place_like_reference():
    v = torch.tensor([0, 127, 128, 255])
    pack_int_data(v, 8)
    assert v.tolist() == [0, 127, -128, -1]  # utils.py:25 wraps the caller's tensor in place


def test_install_as_aqlm_alias():
    import sys

    saved = {k: v for k, v in sys.modules.items() if k == "aqlm" or k.startswith("aqlm.")}
    try:
        aqlm_b200.install_as_aqlm()
        import aqlm
        from aqlm import QuantizedLinear
        from aqlm.inference_kernels import get_forward_pass_kernel  # noqa: F401
        from aqlm.inference_kernels.cuda_kernel import CUDA_KERNEL
        from aqlm.utils import _dequantize_weight, pack_int_data, unpack_int_data  # noqa: F401

        assert aqlm is aqlm_b200 and QuantizedLinear is aqlm_b200.QuantizedLinear
        for n in ("code1x16_matmat", "code2x8_matmat", "code1x8_matmat", "code1x16_matmat_dequant",
                  "code1x16_dequant", "code2x8_dequant"):
            assert callable(getattr(CUDA_KERNEL, n))
    finally:
        for k in [k for k in sys.modules if k == "aqlm" or k.startswith("aqlm.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_group_on_cpu_keeps_members_and_state_dict():
    """Grouping never changes what state_dict exposes; without CUDA tensors nothing is fused (and nothing computes)."""
    ms = [aqlm_b200.QuantizedLinear(64, o, 8, 1, 1, 16, bias=False, dtype=torch.float16) for o in (32, 8, 8)]
    grp = aqlm_b200.QuantizedLinearGroup(ms)
    assert not grp.fused and grp.seg_rows == [32, 8, 8]
    keys = sorted(grp.state_dict().keys())
    assert keys == sorted(f"members.{i}.{n}" for i in range(3) for n in ("codebooks", "codes", "scales"))
    with pytest.raises(NotImplementedError):
        grp(torch.zeros(1, 64, dtype=torch.float16))
    with pytest.raises(ValueError):
        aqlm_b200.QuantizedLinearGroup([ms[0], aqlm_b200.QuantizedLinear(128, 8, 8, 1, 1, 16, bias=False,
                                                                         dtype=torch.float16)])


def test_c_abi_grouped_and_comm_argument_checks_without_gpu():
    L = _cabi.lib()
    w = _cabi.Weight()
    w.codes, w.codebooks, w.scales = 16, 16, 16
    w.in_features, w.out_features = 64, 48
    w.num_codebooks, w.nbits_per_codebook, w.in_group_size, w.out_group_size = 2, 8, 8, 1
    w.dtype = _cabi.F16
    seg = (ctypes.c_int64 * 3)(32, 8, 8)
    # grouped launches exist for 1x16 only
    assert L.aqlm_b200_matmat_grouped(ctypes.byref(w), seg, 3, 16, 16, 1, 0, None) == _cabi.ERR_UNSUPPORTED
    w.num_codebooks, w.nbits_per_codebook = 1, 16
    bad = (ctypes.c_int64 * 3)(32, 8, 7)
    assert L.aqlm_b200_matmat_grouped(ctypes.byref(w), bad, 3, 16, 16, 1, 0, None) == _cabi.ERR_SHAPE
    assert L.aqlm_b200_matmat_grouped(ctypes.byref(w), seg, 5, 16, 16, 1, 0, None) == _cabi.ERR_SHAPE
    assert L.aqlm_b200_matmat_grouped(ctypes.byref(w), seg, 3, 16, 16, 9, 0, None) == _cabi.ERR_UNSUPPORTED
    # communicator sizing is pure arithmetic
    flag_bytes = 16 * 256 * 4  # flag[src rank (<=16)][CTA of the fused GEMV+exchange kernel (<=256)]
    assert L.aqlm_b200_comm_shared_bytes(8, 1024) == flag_bytes + 2 * 8 * 1024 * (4 + 8)  # fp32 slots + tagged 64-bit words
    assert L.aqlm_b200_comm_shared_bytes(0, 1024) == 0 and L.aqlm_b200_comm_shared_bytes(17, 1024) == 0
    assert L.aqlm_b200_allreduce_scale_bias(None, None, None, None, None, 1, 4, 0, None) == _cabi.ERR_SHAPE
    assert L.aqlm_b200_matmat_allreduce(None, ctypes.byref(w), None, 1, 16, 16, 1, None) == _cabi.ERR_SHAPE


def test_workspace_queries_are_zero_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    L = _cabi.lib()
    w = _cabi.Weight()
    w.codes, w.codebooks, w.scales = 16, 16, 16
    w.in_features, w.out_features = 4096, 4096
    w.num_codebooks, w.nbits_per_codebook, w.in_group_size, w.out_group_size = 2, 8, 8, 1
    w.dtype = _cabi.F16
    assert L.aqlm_b200_matmat_workspace_bytes(ctypes.byref(w), 1) == 0
    assert L.aqlm_b200_matmat_dequant_workspace_bytes(ctypes.byref(w), 256) == 0
