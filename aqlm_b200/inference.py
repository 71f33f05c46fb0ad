"""`QuantizedLinear` -- the module API of the hot path (reference inference_lib/src/aqlm/inference.py:11-142).

Same constructor signature, parameter names, shapes and dtypes (inference.py:12-61), same lazy kernel binding and
gemv/gemm dispatch rule (68-96), same autograd wrapper (99-142), so Hugging Face's AQLM integration
(`replace_with_aqlm_linear`) can construct it on the meta device and load a checkpoint by name.
Differences from the reference: CUDA only (CPU inputs raise; the reference's in-place CPU permutation of `codes`,
inference.py:78-83, does not exist here, so `state_dict()` never changes shape), and no JIT build on first call.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import torch
import torch.nn as nn

from .inference_kernels import get_backward_pass_kernel, get_forward_pass_kernel
from .utils import get_int_dtype

# Batch rows (prod of leading dims) up to which the fused GEMV kernel is used; above it the fused
# dequant+GEMM op runs.  The reference uses 6 (inference.py:95-96).
GEMV_MAX_ROWS = 6


class QuantizedLinear(nn.Module):
    def __init__(
        self,
        in_features: int,
        out_features: int,
        in_group_size: int,
        out_group_size: int,
        num_codebooks: int,
        nbits_per_codebook: int,
        bias=True,
        device=None,
        dtype=None,
    ):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features

        assert self.in_features % in_group_size == 0
        assert self.out_features % out_group_size == 0
        num_out_groups = out_features // out_group_size
        num_in_groups = in_features // in_group_size
        self.out_group_size, self.in_group_size = out_group_size, in_group_size
        self.num_codebooks = num_codebooks
        self.nbits_per_codebook = nbits_per_codebook
        self.codebook_size = 2**nbits_per_codebook

        # [num_codebooks, codebook_size, out_group_size, in_group_size]
        self.codebooks = nn.Parameter(
            torch.empty((num_codebooks, self.codebook_size, out_group_size, in_group_size), **factory_kwargs),
            requires_grad=False,
        )
        # [num_out_groups, num_in_groups, num_codebooks], signed storage of unsigned codes (utils.pack_int_data)
        self.codes = nn.Parameter(
            torch.empty((num_out_groups, num_in_groups, num_codebooks), device=device,
                        dtype=get_int_dtype(nbits_per_codebook)),
            requires_grad=False,
        )
        # [num_out_groups, 1, 1, 1]
        self.scales = nn.Parameter(torch.empty((num_out_groups, 1, 1, 1), **factory_kwargs), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features, **factory_kwargs), requires_grad=False)
        else:
            self.register_parameter("bias", None)

        self.gemv_op = None
        self.gemm_op = None
        self.use_gemv_rule = None

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self.gemv_op is None:
            self.prepare_matmul_op(input)
        if self.use_gemv_rule(input):
            return self.gemv_op.apply(input, self.codes, self.codebooks, self.scales, self.bias)
        return self.gemm_op.apply(input, self.codes, self.codebooks, self.scales, self.bias)

    def prepare_matmul_op(self, input: torch.Tensor):
        if not input.is_cuda:
            raise NotImplementedError(
                f"aqlm_b200.QuantizedLinear runs on CUDA (sm_100a) only; got input on {input.device}. "
                "There is no CPU fallback in this package.")
        self.gemv_op = _get_autograd_matmul_op(
            get_forward_pass_kernel(self.codebooks, False),
            get_backward_pass_kernel(self.codebooks, False),
        )
        self.gemm_op = _get_autograd_matmul_op(
            get_forward_pass_kernel(self.codebooks, True),
            get_backward_pass_kernel(self.codebooks, True),
        )
        self.use_gemv_rule = lambda input: math.prod(input.shape[:-1]) <= GEMV_MAX_ROWS

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"scheme={self.num_codebooks}x{self.nbits_per_codebook}, in_group_size={self.in_group_size}, "
                f"bias={self.bias is not None}")


def _get_autograd_matmul_op(forward_pass_kernel, backward_pass_kernel):
    """reference inference.py:99-142: forward = kernel, backward = grad w.r.t. the input only."""

    class _QuantizedMatmul(torch.autograd.Function):
        @staticmethod
        def forward(ctx: Any, input: torch.Tensor, codes: torch.Tensor, codebooks: torch.Tensor,
                    scales: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
            ctx.save_for_backward(input, codes, codebooks, scales, bias)
            return forward_pass_kernel(input, codes, codebooks, scales, bias)

        @staticmethod
        def backward(ctx, grad_output: torch.Tensor):
            input, codes, codebooks, scales, bias = ctx.saved_tensors
            return (backward_pass_kernel(grad_output, codes, codebooks, scales, bias), None, None, None, None)

    return _QuantizedMatmul
