"""Drop-in name: ``import torch_bnb_fp4_ext`` resolves to the B200 op surface
(torch_bnb_fp4_b200/ext.py over libfp4_b200.so)."""
from torch_bnb_fp4_b200.ext import *  # noqa: F401,F403
from torch_bnb_fp4_b200.ext import (ScalarType, bfloat16, dequantize_fp4, dequantize_fp4_codebook,  # noqa: F401
                                    float16, float32, gemv_fp4, qlinear, qlinear_bias,
                                    qlinear_codebook, qlinear_codebook_bias)
