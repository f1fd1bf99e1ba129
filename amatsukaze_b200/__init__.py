"""amatsukaze_b200 -- B200-native (sm_100a) implementation of Amatsukaze's per-frame pixel-analysis hot path.

The product is the CUDA library amatsukaze_b200/lib/libamtk_b200.so behind the C ABI of include/amtk_b200.h;
this package is the thin Python plumbing used by tests/, bench.py and multi-GPU launches.
"""
from .capi import (AmtkError, ClipDesc, CombParams, Context, Group, Logo, LogoScanAcc, calc_fade2, default_comb_params,
                   lib, yv12_clip, LIB_PATH, SIGNATURES)

__all__ = ["AmtkError", "ClipDesc", "CombParams", "Context", "Group", "Logo", "LogoScanAcc", "calc_fade2",
           "default_comb_params", "lib", "yv12_clip", "LIB_PATH", "SIGNATURES"]
