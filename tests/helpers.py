"""Shared helpers for the parity tests (oracle = checker; product = aqlm_b200 CUDA path)."""
import numpy as np
import torch

from oracle import aqlm_oracle as O

TOL_NORTH_STAR = 1e-3  # BASELINE.json north_star: "within 1e-3 rel fp16", metric of matmul_benchmark.py:108
TOL_FP16_TIGHT = 5e-4  # fp32-accumulate + one fp16 rounding measures ~1.8e-4; 5e-4 catches regressions
TOL_BF16 = 4e-3        # bf16 output rounding alone is ~1.4e-3: bf16 is checked against a bf16-fed oracle


def to_torch(case, device, dtype=torch.float16):
    f = lambda a: None if a is None else torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dtype).to(device)  # noqa: E731
    return dict(x=f(case["x"]), codes=torch.from_numpy(case["codes"]).to(device), codebooks=f(case["codebooks"]),
                scales=f(case["scales"]), bias=f(case["bias"]))


def oracle_output(case, dtype=torch.float16):
    """fp32 oracle on the values the GPU actually sees (fp16 inputs are exact; bf16 inputs are re-rounded)."""
    if dtype == torch.float16:
        c = case
    else:
        r = lambda a: None if a is None else torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dtype).float().numpy()  # noqa: E731
        c = dict(x=r(case["x"]), codes=case["codes"], codebooks=r(case["codebooks"]), scales=r(case["scales"]),
                 bias=r(case["bias"]))
    return O.dequantize_gemm(c["x"], c["codes"], c["codebooks"], c["scales"], c["bias"])


def make_module(case, device, dtype=torch.float16):
    import aqlm_b200

    K, cb_size, og, g = case["codebooks"].shape
    out_f, in_groups, _ = case["codes"].shape
    nbits = int(cb_size).bit_length() - 1
    t = to_torch(case, device, dtype)
    layer = aqlm_b200.QuantizedLinear(in_groups * g, out_f, g, 1, K, nbits, bias=case["bias"] is not None,
                                      device=device, dtype=dtype)
    with torch.no_grad():
        layer.codes.data = t["codes"]
        layer.codebooks.data = t["codebooks"]
        layer.scales.data = t["scales"]
        if case["bias"] is not None:
            layer.bias.data = t["bias"]
    return layer, t
