"""CPU oracle for the MSDeformAttn core -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (combo-avs_b200/) never does.  Parity status: pinned against
tests/golden/*.npz generated from the reference's own ms_deform_attn_core_pytorch (gen_golden.py).
"""
from .msda_oracle import (  # noqa: F401
    build,
    forward,
    backward,
    max_threads,
    set_threads,
)
