"""Drop-in module surface (same names as the reference's nn/modules/__init__.py for the hot path)."""
from .block import (DFL, GRN, IFM, SPPF, Bottleneck, C2f, ConvNeXtV2_Block, DyDCNv2, InjectionMultiSum_Auto_pool,
                    LayerNorm, MSPA_C2f, SimFusion_3in, SimFusion_4in, SPRModule, h_sigmoid)
from .conv import Concat, Conv, Upsample, autopad
from .head import Conv_GN, Detect, Scale, TaskDecomposition, TOODHead

__all__ = ("Conv", "Concat", "Upsample", "autopad", "DFL", "SPPF", "Bottleneck", "C2f", "MSPA_C2f", "SPRModule",
           "SimFusion_4in", "SimFusion_3in", "IFM", "ConvNeXtV2_Block", "LayerNorm", "GRN", "h_sigmoid",
           "InjectionMultiSum_Auto_pool", "DyDCNv2", "Conv_GN", "TaskDecomposition", "Detect", "TOODHead", "Scale")
