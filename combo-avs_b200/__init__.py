"""combo-avs_b200: B200-native (sm_100a) multi-scale deformable attention core for COMBO-AVS.

Public surface == the reference's ops package (models/modeling/pixel_decoder/ops):
    MSDeformAttnFunction, MSDeformAttn, ms_deform_attn_forward, ms_deform_attn_backward
plus kernel-variant controls for profiling.  Importing this package loads libmsda_b200.so and fails loudly if it
is missing (the reference raises ModuleNotFoundError with build instructions, functions/ms_deform_attn_func.py:21-29).
The directory name carries a hyphen; import it as `combo_avs_b200` (see combo_avs_b200.py at the repo root).
"""
from . import _lib

_lib.load()  # hard requirement: no extension, no package

from .functions import MSDeformAttnFunction, MSDeformAttnPrologueFunction, ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402
from .modules import MSDeformAttn  # noqa: E402
from ._lib import last_variant, launch_count, select_variant, variant_names  # noqa: E402

__all__ = [
    "MSDeformAttnFunction",
    "MSDeformAttnPrologueFunction",
    "MSDeformAttn",
    "ms_deform_attn_forward",
    "ms_deform_attn_backward",
    "select_variant",
    "variant_names",
    "last_variant",
    "launch_count",
]
