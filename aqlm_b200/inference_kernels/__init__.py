"""Kernel selection and CUDA op registration for aqlm_b200 (mirror of the reference's `aqlm.inference_kernels`)."""
from .kernel_selector import get_backward_pass_kernel, get_forward_pass_kernel, optimize_for_training

__all__ = ["get_forward_pass_kernel", "get_backward_pass_kernel", "optimize_for_training"]
