"""sageattention_b200 — B200-native (sm_100a) drop-in for the thu-ml/SageAttention operator API.

Exports the reference's public names (sageattention/__init__.py:1-5)."""
from .core import (sageattn, sageattn_varlen, sageattn_qk_int8_pv_fp16_triton, sageattn_qk_int8_pv_fp16_cuda,
                   sageattn_qk_int8_pv_fp8_cuda, sageattn_qk_int8_pv_fp8_cuda_sm90)
from .quant import per_block_int8, per_warp_int8, per_thread_int8, per_channel_fp8, k_mean, transpose_v_f16
# beyond the reference surface (SURVEY §8 f-1 and the host-buffer end-to-end call)
from .cache import QuantizedKV, quantize_kv, sageattn_prequantized
from .host import sageattn_host

__all__ = ["sageattn", "sageattn_varlen", "sageattn_qk_int8_pv_fp16_triton", "sageattn_qk_int8_pv_fp16_cuda",
           "sageattn_qk_int8_pv_fp8_cuda", "sageattn_qk_int8_pv_fp8_cuda_sm90",
           "per_block_int8", "per_warp_int8", "per_thread_int8", "per_channel_fp8", "k_mean", "transpose_v_f16",
           "QuantizedKV", "quantize_kv", "sageattn_prequantized", "sageattn_host"]
