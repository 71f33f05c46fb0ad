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


# ---- full-size cases: inputs generated on the GPU, checked against the C oracle (all rows, or a row sample) ----------
def gpu_case(fin, fout, K, nbits, batch, dtype=torch.float16, seed=0, bias=False, device="cuda:0", g=8):
    """Random tensors of a BASELINE-size linear, generated on the device (seeded)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    lo, hi = (-128, 128) if nbits <= 8 else (-(2 ** (nbits - 1)), 2 ** (nbits - 1))
    return dict(
        codes=torch.randint(lo, hi, (fout, fin // g, K), dtype=torch.int8 if nbits <= 8 else torch.int16, device=device,
                            generator=gen),
        codebooks=torch.randn((K, 2**nbits, 1, g), dtype=dtype, device=device, generator=gen),
        scales=(0.75 + 0.5 * torch.rand((fout, 1, 1, 1), device=device, generator=gen)).to(dtype),
        bias=torch.randn((fout,), dtype=dtype, device=device, generator=gen) if bias else None,
        x=torch.randn((batch, fin), dtype=dtype, device=device, generator=gen))


def c_oracle_check(t, y, rows=None, n_sample=384):
    """Relative error (the reference's metric, matmul_benchmark.py:108) of the device result `y` [batch, out] against the
    C oracle (oracle/aqlm_oracle.c: dequantize_gemm, double accumulation) fed with the values the GPU saw.  `rows`:
    None -> every output row when batch == 1, else a seeded sample of `n_sample` rows plus the first and last 64
    (tile edges); the oracle only evaluates those rows (codes[rows]) -- exact for them."""
    from oracle import c_oracle

    fout = t["codes"].shape[0]
    batch = t["x"].shape[0]
    if rows is None and batch > 1 and fout > n_sample + 128:
        rs = np.random.default_rng(fout * 31 + batch).choice(fout, size=n_sample, replace=False)
        rows = np.unique(np.concatenate([np.arange(64), np.arange(fout - 64, fout), rs]))
    f32 = lambda a: None if a is None else a.float().cpu().numpy()  # noqa: E731  (fp16/bf16 -> fp32 is exact)
    codes, scales, bias = t["codes"], t["scales"], t["bias"]
    yy = y.float()
    if rows is not None:
        idx = torch.as_tensor(rows, device=codes.device)
        codes, scales = codes[idx], scales[idx]
        bias = None if bias is None else bias[idx]
        yy = yy[:, idx]
    ref = c_oracle.dequantize_gemm(f32(t["x"]), codes.cpu().numpy(), f32(t["codebooks"]), f32(scales), f32(bias))
    return O.relative_error(yy.cpu().numpy(), ref)
