from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction,
                                  ms_deform_attn_core_pytorch)

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction", "ms_deform_attn_core_pytorch"]
from . import linear_tc  # noqa: E402,F401  (tensor-core Linear for the encoder layer)
