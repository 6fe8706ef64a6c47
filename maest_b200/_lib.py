"""ctypes binding of libmaest_b200.so (include/maest_b200.h).  No torch types cross this boundary: device
pointers (`tensor.data_ptr()`), sizes and the raw CUDA stream handle only.  There is NO fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None
_inited = set()

c_void_p, c_int32, c_int64, c_size_t, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float

F16, BF16, F32 = 0, 1, 2
EPI_STORE16, EPI_GELU16, EPI_RESID32, EPI_STORE32, EPI_GELUBWD16, EPI_ATOMIC32 = 0, 1, 2, 3, 4, 5
EPI_STORE16_LN, EPI_GELU16_LN, EPI_RESID32_LN = 7, 8, 9
ABI_VERSION = 8
HEAD_MEAN, HEAD_SEPARATED = 0, 1


class MaestBlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b", "qkv_wg", "qkv_bf", "fc1_wg", "fc1_bf")]


# name -> (restype, argtypes); must list every symbol include/maest_b200.h declares
SIGNATURES = {
    "maest_last_error": (C.c_char_p, []),
    "maest_abi_version": (c_int32, []),
    "maest_init": (c_int32, [c_int32]),
    "maest_logmel_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p]),
    "maest_logmel_raw16_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int32, c_void_p]),
    "maest_mel_ingest_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_void_p, c_void_p]),
    "maest_adamw_step": (c_int32, [c_void_p, c_void_p, c_int32, c_float, c_float, c_float, c_float, c_float, c_int32, c_float, c_float, c_void_p]),
    "maest_swa_fold": (c_int32, [c_void_p, c_void_p, c_int32, c_float, c_void_p]),
    "maest_ap_roc_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "maest_patch_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "maest_patch_tokens_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                         c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                         c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "maest_wave_tokens_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "maest_wave_tokens_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                        c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "maest_layernorm_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p,
                                      c_void_p, c_void_p]),
    "maest_linear_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "maest_gemm": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                             c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "maest_ln_fold": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "maest_linear_ln_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "maest_ln_finalize": (c_int32, [c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p]),
    "maest_set_gemm_mode": (c_int32, [c_int32]),
    "maest_tmap_cache_stats": (c_int32, [c_void_p, c_void_p]),
    "maest_attention_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "maest_attention_bwd": (c_int32, [c_void_p] * 7 + [c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "maest_mixup_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p]),
    "maest_bce_logits_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "maest_head_bwd": (c_int32, [c_void_p, c_int32, c_int32] + [c_void_p] * 7 + [c_int32] + [c_void_p] * 9),
    "maest_head_bwd_separated": (c_int32, [c_void_p, c_int32, c_int32] + [c_void_p] * 9 + [c_int32] + [c_void_p] * 12),
    "maest_layernorm_bwd": (c_int32, [c_void_p] * 7 + [c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "maest_colsum": (c_int32, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "maest_cast_rows16": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "maest_token_grad": (c_int32, [c_void_p] + [c_int32] * 7 + [c_void_p] * 7 + [c_void_p]),
    "maest_encoder_workspace_bytes": (c_size_t, [c_int64]),
    "maest_encoder_fwd": (c_int32, [c_void_p, c_int32, c_int32, C.POINTER(MaestBlockWeights), c_int32, c_int32, c_int32,
                                    c_int32, c_void_p, c_size_t, c_void_p]),
    "maest_pool_head_fwd": (c_int32, [c_void_p, c_int32, c_int32] + [c_void_p] * 8 + [c_int32, c_int32] + [c_void_p] * 6),
    "maest_block_embedding_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "maest_cast_to16": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
}


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen the in-tree library (building it first if nvcc is available and sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MAEST_B200_LIB") or _build.LIB_PATH
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"libmaest_b200.so is missing and could not be built: {e}") from e
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().maest_last_error().decode("utf-8", "replace")
        if code == -2:
            # the reference raises a plain Exception here (models/maest.py:664-668)
            raise Exception(msg)
        raise RuntimeError(f"libmaest_b200 {what} failed ({code}): {msg}")


def init(device_index: int):
    lib = load()
    if device_index not in _inited:
        check(lib.maest_init(int(device_index)), "maest_init")
        _inited.add(device_index)
    return lib
