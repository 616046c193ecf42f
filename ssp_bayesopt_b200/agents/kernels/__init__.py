from .sinc_kernel import SincKernel

__all__ = ["SincKernel"]
