"""Drop-in name: ``import torch_bnb_fp4`` resolves to the B200 implementation."""
from torch_bnb_fp4_b200 import *  # noqa: F401,F403
from torch_bnb_fp4_b200 import (QuantData, ScalarType, T_Model, TorchFP4Linear,  # noqa: F401
                                check_if_name_contained_in_list, dequantize_fp4,
                                dequantize_fp4_codebook_invoke, dequantize_fp4_codebook_invoke_qtype,
                                dequantize_fp4_qtype, gemm_4bit_inference, gemm_4bit_inference_qtype,
                                recursively_replace_with_fp4_linear, swap_linear_with_bnb_linear,
                                todevice_if_necessary)
