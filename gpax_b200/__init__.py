"""
gpax_b200 -- B200-native exact-GP posterior path behind the gpax API (ExactGP / viGP / viSparseGP
`.predict()`, `.predict_in_batches()`, `.get_mvn_posterior()`, and the RBF / Matern / Periodic kernel
functions).  Python host code + ctypes -> libb200gp.so (hand-written sm_100a CUDA).  No JAX, no torch,
no CPU fallback on the product path.
"""
from . import kernels, utils
from .kernels import MaternKernel, PeriodicKernel, RBFKernel, get_kernel
from .gp import ExactGP
from .vigp import viGP
from .sparse_gp import viSparseGP
from .variants import MeasuredNoiseGP, UIGP, VarNoiseGP, vExactGP
from . import acquisition, mtkernels
from .kernels import NNGPKernel
from .mtkernels import LCMKernel, MultitaskKernel, MultivariateKernel
from ._ffi import B200GPError, Context, default_context

__version__ = "0.1.0"
__all__ = ["ExactGP", "viGP", "viSparseGP", "MeasuredNoiseGP", "VarNoiseGP", "vExactGP", "UIGP", "acquisition", "RBFKernel", "MaternKernel", "PeriodicKernel", "NNGPKernel", "MultitaskKernel", "MultivariateKernel", "LCMKernel", "mtkernels", "get_kernel",
           "kernels", "utils", "Context", "default_context", "B200GPError"]
