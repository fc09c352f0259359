"""Name-compatible stand-in for the reference's compiled extension module `MultiScaleDeformableAttention`
(models/modeling/pixel_decoder/ops/src/vision.cpp:18-21).  Put this directory (and the repo root) on PYTHONPATH and the
reference's own `ops/functions/ms_deform_attn_func.py` -- import guard at :21-29 -- binds to libmsda_b200.so instead of its
bundled CUDA extension, with no edit to any reference file."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from combo_avs_b200 import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402,F401
