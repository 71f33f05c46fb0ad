"""Drop-in boundary, packaging row: the repo installs as the distribution `aqlm` (reference
inference_lib/setup.cfg:2-3), so `import aqlm`, `from aqlm import QuantizedLinear` and
`importlib.metadata.version("aqlm")` (what HF's AqlmHfQuantizer.is_trainable reads) resolve to this implementation."""
import os
import shutil
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r"""
import sys
sys.path[:] = [p for p in sys.path if p and 'repo' not in p.split('/')[-1:]]
import importlib.metadata as md
from packaging import version
import aqlm
assert aqlm.__name__ == "aqlm_b200" and TARGET in aqlm.__file__, aqlm.__file__
from aqlm import QuantizedLinear
from aqlm.inference_kernels import get_forward_pass_kernel, get_backward_pass_kernel, optimize_for_training
from aqlm.inference_kernels.cuda_kernel import CUDA_KERNEL
from aqlm.utils import get_int_dtype, pack_int_data, unpack_int_data, _dequantize_weight
assert hasattr(CUDA_KERNEL, "code1x16_matmat") and hasattr(CUDA_KERNEL, "code2x8_matmat")
v = md.version("aqlm")
assert version.parse(v) >= version.parse("1.1.6"), v
from aqlm_b200 import _cabi
assert TARGET in _cabi.LIB_PATH and _cabi.lib().aqlm_b200_version() == 100
import torch
m = QuantizedLinear(64, 32, 8, 1, 1, 16, bias=False, device="meta", dtype=torch.float16)
assert m.codes.shape == (32, 8, 1)
print("ok", v)
"""


def test_in_tree_alias_exposes_distribution_metadata():
    code = ("import sys; sys.path.insert(0, %r); import aqlm_b200; aqlm_b200.install_as_aqlm(); "
            "import importlib.metadata as md; import aqlm; assert aqlm is aqlm_b200; print(md.version('aqlm'))" % REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().startswith("1.1.6")


def test_pip_install_provides_the_aqlm_distribution(tmp_path):
    if not os.path.exists(os.path.join(REPO, "aqlm_b200", "csrc", "libaqlm_b200.so")):
        pytest.skip("CUDA library not built (run __graft_entry__.build())")
    target = str(tmp_path / "site")
    src = str(tmp_path / "src")
    # install from a copy so that pip's build/ and egg-info never land in the work tree
    shutil.copytree(REPO, src, ignore=shutil.ignore_patterns(".git", "gpurun_out", "baseline", "profiles", "tests", "tools",
                                                             "oracle", "__pycache__", "*.json", "build", "*.egg-info"))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                        "--target", target, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert any(d.startswith("aqlm-") and d.endswith(".dist-info") for d in os.listdir(target))
    env = dict(os.environ, PYTHONPATH=target)
    out = subprocess.run([sys.executable, "-c", f"TARGET = {target!r}\n" + CHECK], capture_output=True, text=True, cwd="/tmp",
                         env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.strip().startswith("ok 1.1.6")
