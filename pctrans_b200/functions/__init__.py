from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction,
                                  ms_deform_attn_core_pytorch)

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction", "ms_deform_attn_core_pytorch"]
