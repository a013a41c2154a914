"""Host-side mirror of the reference's FAC operator interface
(models/FAC/kernelconv2d/KernelConv2D.py): `KernelConv2DFunction` (:12-58) and the
`KernelConv2D` module (:77-87), on top of the sm_100a kernels.

Kept from the reference: contiguity asserts (:19-20), the K == sqrt(C_k/C) and
H_in - K == H_out - 1 asserts (:22,:31-32), NotImplementedError for CPU tensors (:38-39,:55-56).
Dropped: the zero fill of output / grad_input / grad_kernel (:35,:50-51) — 1.7 GB of pure
memset at the benchmark shape; the kernels write every element.
"""
import torch
from torch import nn
from torch.autograd import Function

from .shims import kernelconv2d_cuda


class KernelConv2DFunction(Function):
    @staticmethod
    def forward(ctx, input, kernel, kernel_size):
        assert input.is_contiguous()
        assert kernel.is_contiguous()
        assert kernel_size == int((kernel.size(1) / input.size(1)) ** 0.5)
        assert input.size(2) - kernel_size == kernel.size(2) - 1
        assert input.size(3) - kernel_size == kernel.size(3) - 1
        if not input.is_cuda:
            raise NotImplementedError()      # CPU VERSION NOT IMPLEMENTED (KernelConv2D.py:38-39)
        ctx.kernel_size = kernel_size
        ctx.save_for_backward(input, kernel)
        with torch.cuda.device_of(input):
            output = torch.empty((input.size(0), input.size(1), kernel.size(2), kernel.size(3)),
                                 dtype=input.dtype, device=input.device)
            kernelconv2d_cuda.forward(input, kernel, kernel_size, output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, kernel = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        if not grad_output.is_cuda:
            raise NotImplementedError()      # KernelConv2D.py:55-56
        with torch.cuda.device_of(input):
            grad_input, grad_kernel = torch.empty_like(input), torch.empty_like(kernel)
            kernelconv2d_cuda.backward(input, kernel, ctx.kernel_size, grad_output, grad_input, grad_kernel)
        return grad_input, grad_kernel, None


class KernelConv2D(nn.Module):
    """ReplicationPad2d((K-1)/2) followed by the per-pixel K x K filter (KernelConv2D.py:77-87)."""

    def __init__(self, kernel_size):
        super().__init__()
        assert kernel_size % 2 == 1
        self.kernel_size = kernel_size
        self.pad = nn.ReplicationPad2d([(kernel_size - 1) // 2] * 4)

    def forward(self, input, kernel):
        return KernelConv2DFunction.apply(self.pad(input), kernel, self.kernel_size)
