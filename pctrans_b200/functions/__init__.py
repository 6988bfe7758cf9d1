from .ms_deform_attn_func import MSDeformAttnFunction, ms_deform_attn_core_pytorch

__all__ = ["MSDeformAttnFunction", "ms_deform_attn_core_pytorch"]
