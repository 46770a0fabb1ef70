"""Case tables shared by the oracle tests (CPU) and the kernel parity tests (GPU).
Must stay in sync with oracle/make_golden.py (which imports the same tables)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.make_golden import (ATTN_CASES, WL_CASES, SL_CASES, STAGE_CASES, SUBSAMPLE, F64_SKIP,  # noqa: E402,F401
                                synth_sent_inputs, synth_stage_inputs)
