"""pctrans_b200 -- B200-native multi-scale deformable attention (MSDeformAttn) for PCTrans.

Drop-in for the one CUDA op of PCTrans's pixel decoder:

    from pctrans_b200.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
    from pctrans_b200.modules import MSDeformAttn
    from pctrans_b200.encoder import MSDeformAttnTransformerEncoderOnly, MSDeformAttnPixelDecoder
    from pctrans_b200.host_pipeline import HostPipeline

Everything runs through ``libpctrans_msda.so`` (C ABI in ``include/pct_msda.h``, built by
``python -m pctrans_b200.build``); there is no CPU or PyTorch fallback.
"""

__version__ = "0.1.0"
__all__ = ["__version__"]
