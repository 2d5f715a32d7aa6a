"""gabotorch_b200 -- B200-native (sm_100a) implementation of GaBOtorch's data-parallel hot path.

Geodesic-kernel Gram builds on S^d and SPD(d), the batched Riemannian acquisition optimiser and the nested SPD
projection, behind the reference's own class / function names.  Host code is Python/PyTorch (device memory, streams,
torch.distributed); the arithmetic runs in hand-written CUDA kernels reached through the C ABI of
``include/gabo_b200.h`` (``gabotorch_b200/lib/libgabo_b200.so``).  There is no CPU fallback.
"""
__version__ = '0.1.0'
