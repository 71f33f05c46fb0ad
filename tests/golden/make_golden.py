"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (Vahe1994/AQLM inference_lib) on CPU.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    AQ_USE_JIT=0 NUMBA_NUM_THREADS=1 python tests/golden/make_golden.py

Inputs are produced by `oracle.aqlm_oracle.make_case(seed, ...)` (numpy PCG64, deterministic) and are NOT
stored (a 1x16 codebook alone is 1 MiB); the fixture stores a sha256 of every input so that drift in the
generator is detected, plus the reference's outputs:

  * `dequantize_gemm`                      inference_kernels/dequantization.py:9-21   (fp32 and fp16)
  * `QuantizedLinear.forward` on CPU       inference.py:68-75 (1x16 -> dequantize_gemm; Kx8 fp32 -> numba LUT kernel,
                                           kernel_selector.py:95-98, numba_kernel.py:10-65, single thread)
  * `_dequantize_weight`                   utils.py:43-70  (full W for the small cases)
  * `pack_int_data` / `unpack_int_data`    utils.py:23-31

Output: tests/golden/reference_vectors.npz + reference_vectors.json (case table).
"""
import hashlib
import json
import os
import sys

os.environ.setdefault("AQ_USE_JIT", "0")
os.environ.setdefault("NUMBA_NUM_THREADS", "1")

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference/inference_lib/src")

import aqlm  # the REFERENCE package  # noqa: E402
from aqlm.inference_kernels.dequantization import dequantize_gemm  # noqa: E402
from aqlm.utils import _dequantize_weight, pack_int_data, unpack_int_data  # noqa: E402

from oracle import aqlm_oracle as O  # noqa: E402

assert aqlm.__file__.startswith("/root/reference"), aqlm.__file__

# name, seed, in, out, K, nbits, g, batch, bias
CASES = [
    ("cfg0_1x16_4096x4096_bs1", 1000, 4096, 4096, 1, 16, 8, 1, False),  # BASELINE.json configs[0]
    ("1x16_g8_small_bs1", 1001, 256, 64, 1, 16, 8, 1, False),
    ("1x16_g8_small_bs3_bias", 1002, 512, 96, 1, 16, 8, 3, True),
    ("1x16_g8_ragged_bs5", 1003, 1032, 40, 1, 16, 8, 5, False),  # in_groups=129: not a multiple of 8/32
    ("1x16_g16_bs2", 1004, 512, 64, 1, 16, 16, 2, True),
    ("2x8_g8_bs1", 1005, 512, 128, 2, 8, 8, 1, False),
    ("2x8_g8_bs4_bias", 1006, 1024, 72, 2, 8, 8, 4, True),
    ("1x8_g8_bs1", 1007, 512, 128, 1, 8, 8, 1, False),
    ("8x8_g8_bs1", 1008, 512, 128, 8, 8, 8, 1, False),
    ("8x8_g8_bs2_bias", 1009, 256, 48, 8, 8, 8, 2, True),
    ("1x16_g8_bs16", 1010, 512, 128, 1, 16, 8, 16, False),  # gemm-mode batch
    ("2x8_g8_bs64", 1011, 256, 256, 2, 8, 8, 64, True),
    ("4x8_g8_bs1", 1012, 256, 64, 4, 8, 8, 1, False),  # non-published KxN
    ("2x12_g8_bs1", 1013, 256, 64, 2, 12, 8, 1, False),  # nbits not a multiple of 8 (int16 storage)
]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    out = {}
    table = []
    for name, seed, fin, fout, K, nbits, g, batch, has_bias in CASES:
        case = O.make_case(seed, fin, fout, K, nbits, g, batch, has_bias)
        x16, codes, cb16, sc16, b16 = (case[k] for k in ("x", "codes", "codebooks", "scales", "bias"))
        t = lambda a, dt: None if a is None else torch.from_numpy(np.asarray(a)).to(dt)  # noqa: E731
        tcodes = torch.from_numpy(codes)

        # 1. dequantize_gemm in fp32 on the fp16-valued inputs: the tightest reference
        y32 = dequantize_gemm(t(x16, torch.float32), tcodes, t(cb16, torch.float32), t(sc16, torch.float32),
                              t(b16, torch.float32)).numpy()
        out[f"{name}/y_dequantize_gemm_fp32"] = y32
        # 2. dequantize_gemm in fp16 (what QuantizedLinear does on CPU for fp16 checkpoints)
        y16 = dequantize_gemm(t(x16, torch.float16), tcodes, t(cb16, torch.float16), t(sc16, torch.float16),
                              t(b16, torch.float16)).numpy()
        out[f"{name}/y_dequantize_gemm_fp16"] = y16

        # 3. the module itself, CPU fp32 (1x16 -> dequantize_gemm; 256-entry codebooks -> numba LUT kernel)
        layer = aqlm.QuantizedLinear(fin, fout, g, 1, K, nbits, bias=has_bias, dtype=torch.float32)
        with torch.no_grad():
            layer.codes.data = tcodes.clone()
            layer.codebooks.data = t(cb16, torch.float32)
            layer.scales.data = t(sc16, torch.float32)
            if has_bias:
                layer.bias.data = t(b16, torch.float32)
            ymod = layer(t(x16, torch.float32)).numpy()
        out[f"{name}/y_module_cpu_fp32"] = ymod
        used_numba = (2**nbits == 256)

        # 4. full dequantized weight for the small cases; strided sample for cfg0
        W = _dequantize_weight(unpack_int_data(tcodes, nbits), t(cb16, torch.float32), t(sc16, torch.float32)).numpy()
        if W.size <= 1 << 14:
            out[f"{name}/W_fp32"] = W
        else:
            out[f"{name}/W_fp32_rows0_8"] = W[:8].copy()
            out[f"{name}/W_fp32_rowsum"] = W.sum(axis=1, dtype=np.float64)

        table.append(dict(name=name, seed=seed, in_features=fin, out_features=fout, num_codebooks=K, nbits=nbits,
                          in_group_size=g, batch=batch, bias=has_bias, module_path="numba_gemm_lut" if used_numba
                          else "dequantize_gemm", inputs_sha256=sha(x16, codes, cb16, sc16, b16)))
        print(name, "ok  rel(fp16 vs fp32) =", O.relative_error(y16, y32), " rel(module vs fp32) =",
              O.relative_error(ymod, y32))

    # general out_group_size > 1 (only _dequantize_weight / dequantize_gemm support it)
    rng = np.random.default_rng(2000)
    og, ig, K, nbits = 2, 4, 2, 8
    raw = rng.integers(0, 2**nbits, size=(24, 32, K))
    cbk = rng.standard_normal((K, 2**nbits, og, ig), dtype=np.float32)
    scl = rng.standard_normal((24, 1, 1, 1), dtype=np.float32)
    Wg = _dequantize_weight(torch.from_numpy(raw), torch.from_numpy(cbk), torch.from_numpy(scl)).numpy()
    out["general_og2/raw_codes"] = raw
    out["general_og2/codebooks"] = cbk
    out["general_og2/scales"] = scl
    out["general_og2/W_fp32"] = Wg

    # pack / unpack known answers (utils.py:23-31)
    for nbits in (1, 7, 8, 12, 16):
        vals = np.concatenate([np.arange(0, min(2**nbits, 64)), np.arange(max(0, 2**nbits - 64), 2**nbits),
                               np.random.default_rng(nbits).integers(0, 2**nbits, 256)]).astype(np.int64)
        packed = pack_int_data(torch.from_numpy(vals.copy()), nbits)
        out[f"pack/nbits{nbits}_values"] = vals
        out[f"pack/nbits{nbits}_packed"] = packed.numpy()
        out[f"pack/nbits{nbits}_unpacked"] = unpack_int_data(packed, nbits).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(dict(reference_commit="e79a896", torch=torch.__version__, numpy=np.__version__, cases=table), f,
                  indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
