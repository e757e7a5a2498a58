"""Drop-in for lib/model/correlation/functions/correlation.py:6-50.

The reference class is a legacy (instance-style) autograd Function that torch >= 1.3 refuses to
run; this keeps its *instantiate-then-call* shape -- ``CorrelationFunction(pad, k, md, s1, s2,
mult)(input1, input2)`` -- over a new-style Function backed by d2t_correlation_forward/backward.
"""
from d2t_b200 import ops


class CorrelationFunction(object):
    def __init__(self, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply   # accepted and ignored, as in correlation_cuda_kernel.cu

    def __call__(self, input1, input2):
        # functions/correlation.py:21-22 asserts contiguity
        assert input1.is_contiguous() and input2.is_contiguous()
        return ops.correlation(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                               self.stride1, self.stride2)

    forward = __call__
