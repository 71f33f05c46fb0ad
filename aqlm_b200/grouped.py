"""Grouped launch for quantized linears that share their input (q/k/v, gate/up): one kernel instead of n.

New work (the reference launches every linear on its own).  `QuantizedLinearGroup` fuses the STORAGE of its
members — codes and scales (and biases) are concatenated along the output dimension, the 1 MiB codebooks are stacked — and
re-points every member's parameters at views of the fused buffers, so the members keep working on their own and
`state_dict()` (names, shapes, dtypes) is unchanged.  `forward(x)` returns one output per member.  The fused kernel path
covers the 1x16 / in_group 8 scheme with up to 8 batch rows; anything else falls back to calling the members one by one.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn as nn

from .inference import QuantizedLinear
from .sharded import ShardedQuantizedLinear


def _fuse_storage(members) -> dict:
    with torch.no_grad():
        codes = torch.cat([m.codes.data for m in members], dim=0).contiguous()
        codebooks = torch.stack([m.codebooks.data for m in members], dim=0).contiguous()
        scales = torch.cat([m.scales.data for m in members], dim=0).contiguous()
        bias = None
        if members[0].bias is not None:
            bias = torch.cat([m.bias.data for m in members], dim=0).contiguous()
        off = 0
        for i, m in enumerate(members):
            n = m.codes.shape[0]
            m.codes.data = codes[off:off + n]
            m.codebooks.data = codebooks[i]
            m.scales.data = scales[off:off + n]
            if bias is not None:
                m.bias.data = bias[off:off + n]
            off += n
    return dict(codes=codes, codebooks=codebooks, scales=scales, bias=bias)


def _check_members(members) -> bool:
    """True when the fused kernel applies; raises on members that cannot be grouped at all."""
    m0 = members[0]
    for m in members:
        if m.in_features != m0.in_features or m.codebooks.dtype != m0.codebooks.dtype or \
                m.codes.device != m0.codes.device or (m.bias is None) != (m0.bias is None) or \
                (m.num_codebooks, m.nbits_per_codebook, m.in_group_size) != \
                (m0.num_codebooks, m0.nbits_per_codebook, m0.in_group_size):
            raise ValueError("grouped linears must share in_features, scheme, dtype, device and bias-ness")
    return (m0.num_codebooks, m0.nbits_per_codebook, m0.in_group_size, m0.out_group_size) == (1, 16, 8, 1) and \
        len(members) <= 4 and m0.codes.is_cuda


class QuantizedLinearGroup(nn.Module):
    def __init__(self, members: Sequence[QuantizedLinear]):
        super().__init__()
        self.members = nn.ModuleList(members)
        self.fused = _check_members(list(members))
        self.seg_rows = [m.out_features for m in members]
        if self.fused:
            for k, v in _fuse_storage(list(members)).items():
                if v is not None:
                    self.register_buffer(f"_fused_{k}", v, persistent=False)
            if members[0].bias is None:
                self._fused_bias = None

    def forward(self, input: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        rows = 1
        for d in input.shape[:-1]:
            rows *= d
        # the fused launch has no autograd node: when a gradient w.r.t. the input is needed (LoRA / PEFT on frozen AQLM
        # weights), go through the members, whose forward builds one
        needs_grad = torch.is_grad_enabled() and input.requires_grad
        if not self.fused or not input.is_cuda or rows > 8 or rows < 1 or needs_grad:
            return tuple(m(input) for m in self.members)
        from .inference_kernels import cuda_kernel

        y = cuda_kernel.matmat_grouped(input, self._fused_codes, self._fused_codebooks, self._fused_scales,
                                       self._fused_bias, self.seg_rows)
        return tuple(torch.split(y, self.seg_rows, dim=-1))


class ShardedQuantizedLinearGroup(nn.Module):
    """Same for the in_features-sharded path: ONE GEMV launch and ONE exchange for the whole group."""

    def __init__(self, members: Sequence[ShardedQuantizedLinear]):
        super().__init__()
        self.members = nn.ModuleList(members)
        self.fused = _check_members(list(members))
        self.seg_rows = [m.out_features for m in members]
        m0 = members[0]
        self.world_size, self.process_group, self.peer_comm = m0.world_size, m0.process_group, m0.peer_comm
        self.in_begin, self.in_end, self.in_features = m0.in_begin, m0.in_end, m0.in_features
        if self.fused:
            for k, v in _fuse_storage(list(members)).items():
                if v is not None:
                    self.register_buffer(f"_fused_{k}", v, persistent=False)
            if m0.bias is None:
                self._fused_bias = None

    def forward(self, input: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        local = self.in_end - self.in_begin
        if input.shape[-1] == self.in_features and self.world_size > 1:
            input = input[..., self.in_begin:self.in_end]
        rows = 1
        for d in input.shape[:-1]:
            rows *= d
        if not self.fused or not input.is_cuda or rows > 8 or rows < 1 or (torch.is_grad_enabled() and input.requires_grad):
            return tuple(m(input) for m in self.members)
        import torch.distributed as dist

        from .inference_kernels import cuda_kernel

        flat = input.reshape(-1, local)
        if self.world_size > 1 and self.peer_comm is not None and getattr(self.members[0], "fused_exchange", True):
            # ONE kernel for the whole group: grouped GEMV + NVLink exchange + scale/bias
            y = self.peer_comm.matmat_allreduce(flat, self._fused_codes, self._fused_codebooks, self._fused_scales,
                                                self._fused_bias, self.seg_rows)
            if y is not None:
                y = y.reshape(input.shape[:-1] + (sum(self.seg_rows),))
                return tuple(torch.split(y, self.seg_rows, dim=-1))
        partial = cuda_kernel.matmat_grouped(flat, self._fused_codes, self._fused_codebooks, None, None, self.seg_rows,
                                             partial=True)
        if self.world_size > 1 and self.peer_comm is not None:
            y = self.peer_comm.allreduce_scale_bias(partial, self._fused_scales, self._fused_bias, input.dtype)
        else:
            if self.world_size > 1:
                dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.process_group)
            y = cuda_kernel.scale_bias(partial, self._fused_scales, self._fused_bias, input.dtype)
        y = y.reshape(input.shape[:-1] + (sum(self.seg_rows),))
        return tuple(torch.split(y, self.seg_rows, dim=-1))
