"""megengine_b200 -- B200-native (sm_100a) ConvBiasForward behind megdnn's operator API.

Only the hot path named by BASELINE.json lives here: the C-ABI launch layer + CUDA kernels
(`csrc/`, built into `lib/libb200cb.so`) and the host-side mirror of megdnn's ConvBiasForward
interface (`conv_bias.py`).  See DESIGN.md / INTEGRATION.md.
"""
from . import _capi  # noqa: F401
from .conv_bias import (  # noqa: F401
    ConvBiasForward, DType, EMPTY, Handle, MegDNNError, Param, TensorLayout, TensorND, dtype,
    empty_like_dtype, make_tensor, mul_scale,
)

from .nn_ops import (  # noqa: F401  (the operators either side of ConvBias)
    BatchConvBiasForward, ConvolutionBackwardData, ConvolutionForward, ElemwiseForward, ElemwiseMultiType,
    PoolingForward, PoolingParam, TypeCvt,
)

__all__ = [
    "BatchConvBiasForward", "ConvolutionBackwardData", "ConvolutionForward", "ElemwiseForward", "ElemwiseMultiType",
    "PoolingForward", "PoolingParam", "TypeCvt",
    "ConvBiasForward", "DType", "EMPTY", "Handle", "MegDNNError", "Param", "TensorLayout", "TensorND",
    "dtype", "empty_like_dtype", "make_tensor", "mul_scale",
]
