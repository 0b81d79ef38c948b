"""gpu_stereo_matching_b200 -- B200-native (sm_100a) BlockMatching hot path.

CUDA kernels + C ABI live in csrc/ (built into libgsm.so); api.py mirrors the reference's
Device.cuh / Caller.h entry points over that ABI.  No CPU fallback.
"""
from .api import (GF_EPS_DEFAULT, GSM_MODE_GF, GSM_MODE_SAD, GsmError, GsmParams, StereoContext,  # noqa: F401
                  blockMatching_gpu, compare_disp, cvtColor_gpu, make_params, remap_gpu, singleFrame,
                  st_build_tree_host)

__all__ = ["StereoContext", "GsmParams", "GsmError", "make_params", "blockMatching_gpu", "singleFrame",
           "compare_disp", "remap_gpu", "cvtColor_gpu", "st_build_tree_host", "GSM_MODE_SAD", "GSM_MODE_GF", "GF_EPS_DEFAULT"]
