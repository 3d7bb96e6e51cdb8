"""ctypes binding of libsimx_b200.so (the C ABI declared in include/simx_b200.h).

The library is the product: there is no PyTorch / CPU fallback behind these calls.  If the shared object is
missing, or a call returns a non-zero status, a RuntimeError is raised — loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsimx_b200.so"
CSRC = _PKG / "csrc"

SIMX_OK = 0
ARCH_BERT, ARCH_DISTILBERT = 0, 1
SAVE_ALL, SAVE_LAYER_INPUTS = 0, 1
PREC_BF16, PREC_BF16X3 = 1, 3

EPI_PLAIN, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_ADD, EPI_MUL, EPI_ADD, EPI_ATOMIC_F32, EPI_BIAS_GELU_FWD = range(8)

LOSS_NLL_GLOBAL, LOSS_KL_LOCAL, LOSS_CE_LOCAL, LOSS_KD_LOCAL, LOSS_NLL_LOCAL, LOSS_ADV_LOCAL = 1, 2, 4, 8, 16, 32
LOSS_MSE_LOCAL, LOSS_DKD_LOCAL = 64, 128
LOSS_GLOBAL_UNFUSED = 256


class EncoderCfg(C.Structure):
    """mirror of `simx_encoder_cfg`"""

    _fields_ = [
        ("arch", C.c_int32),
        ("num_layers", C.c_int32),
        ("hidden", C.c_int32),
        ("heads", C.c_int32),
        ("intermediate", C.c_int32),
        ("vocab", C.c_int32),
        ("max_pos", C.c_int32),
        ("type_vocab", C.c_int32),
        ("precision", C.c_int32),
        ("ln_eps", C.c_float),
        ("hidden_dropout", C.c_float),
        ("attn_dropout", C.c_float),
        ("dropout_seed", C.c_uint64),
        ("save_policy", C.c_int32),
        ("id_bytes", C.c_int32),
    ]


class LossCfg(C.Structure):
    """mirror of `simx_loss_cfg`"""

    _fields_ = [
        ("terms", C.c_int32),
        ("GQ", C.c_int32),
        ("GC", C.c_int32),
        ("H", C.c_int32),
        ("q_begin", C.c_int32),
        ("q_count", C.c_int32),
        ("c_begin", C.c_int32),
        ("c_count", C.c_int32),
        ("group", C.c_int32),
        ("w_nll_global", C.c_float),
        ("w_kl_local", C.c_float),
        ("w_ce_local", C.c_float),
        ("w_kd_local", C.c_float),
        ("w_nll_local", C.c_float),
        ("temperature", C.c_float),
        ("eps", C.c_float),
        ("score_scale", C.c_float),
        ("grad_scale", C.c_float),
        ("w_adv_local", C.c_float),
        ("w_mse_local", C.c_float),
        ("w_dkd_local", C.c_float),
        ("dkd_alpha", C.c_float),
        ("dkd_beta", C.c_float),
    ]


class PairCfg(C.Structure):
    """mirror of `simx_pair_cfg`"""

    _fields_ = [(n, C.c_int32) for n in ("GQ", "GC", "H", "q_begin", "q_count", "c_begin", "c_count")] + \
               [(n, C.c_float) for n in ("w_nll_a", "w_nll_b", "w_kl_b2a", "w_kl_a2b", "temperature", "grad_scale", "w_kd_b2a",
                                    "w_mse_b2a", "w_dkd_b2a", "dkd_alpha", "dkd_beta")]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_CFG = C.POINTER(EncoderCfg)
_LCFG = C.POINTER(LossCfg)

# name -> (restype, argtypes); kept in the order of include/simx_b200.h (tests check both agree)
SIGNATURES = {
    "simx_last_error": (C.c_char_p, []),
    "simx_version": (_I, []),
    "simx_device_supported": (_I, []),
    "simx_profile_begin": (_I, []),
    "simx_profile_end": (_I, [C.c_char_p, C.c_size_t]),
    "simx_encoder_num_params": (_I, [_CFG]),
    "simx_encoder_param_numel": (_L, [_CFG, _I]),
    "simx_encoder_grad_offset": (_L, [_CFG, _I]),
    "simx_encoder_grad_numel": (_L, [_CFG]),
    "simx_encoder_weight_bytes": (C.c_size_t, [_CFG]),
    "simx_encoder_saved_bytes": (C.c_size_t, [_CFG, _I, _I]),
    "simx_encoder_workspace_bytes": (C.c_size_t, [_CFG, _I, _I]),
    "simx_encoder_prepare_weights": (_I, [_CFG, _P, _P, _P]),
    "simx_encoder_forward": (_I, [_CFG, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "simx_encoder_backward": (_I, [_CFG, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "simx_encoder_forward_pos": (_I, [_CFG, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "simx_encoder_backward_pos": (_I, [_CFG, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "simx_encoder_forward_embeds": (_I, [_CFG, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "simx_encoder_backward_embeds": (_I, [_CFG, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "simx_encoder_backward_layers": (_I, [_CFG, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P]),
    "simx_encoder_hidden_state": (_I, [_CFG, _P, _I, _I, _I, _P, _L, _P]),
    "simx_encoder_add_cotangent": (_I, [_CFG, _P, _I, _I, _P, _L, _P]),
    "simx_score_loss_workspace_bytes": (C.c_size_t, [_LCFG]),
    "simx_score_loss": (_I, [_LCFG, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "simx_pair_distill_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "simx_pair_distill": (_I, [C.POINTER(PairCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "simx_local_pair_distill": (_I, [C.POINTER(PairCfg), _P, _P, _P, _P, _P, _P, _P, _P]),
    "simx_embed_mse": (_I, [_P, _P, _L, C.c_float, _P, _P, _P]),
    "simx_topk_corpus_bytes": (C.c_size_t, [_L, _I]),
    "simx_topk_add": (_I, [_P, _L, _L, _L, _I, _P, _P]),
    "simx_topk_workspace_bytes": (C.c_size_t, [_I, _L, _I, _I]),
    "simx_topk_search": (_I, [_P, _I, _P, _L, _I, _I, _P, _P, _P, _P]),
    "simx_simans_sample": (_I, [_P, _P, _P, _I, _I, _I, _I, C.c_float, C.c_float, C.c_uint64, _P, _P]),
    "simx_attention_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _P]),
    "simx_attention_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _P, _P]),
    "simx_gemm": (
        _I,
        [_I, _I, _I, _P, _L, _L, _I, _P, _L, _L, _I, _I, _I, _P, _P, _L, _L, _P, _P, _L, _L, _P, _L, _I, _P],
    ),
}

_lib = None


class OptTensor(C.Structure):
    """mirror of `simx_opt_tensor`"""

    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("numel", C.c_int64),
                ("chunk_begin", C.c_int32), ("lr", C.c_float), ("weight_decay", C.c_float), ("reserved", C.c_int32)]


ADAMW_TORCH, ADAMW_HF = 0, 1
SIGNATURES.update({
    "simx_opt_plan": (_I, [C.POINTER(OptTensor), _I]),
    "simx_grad_sq_norm": (_I, [_P, _I, _I, _P, _P]),
    "simx_clip_grads": (_I, [_P, _I, _I, _P, C.c_float, _P]),
    "simx_adamw_step": (_I, [_P, _I, _I, _I, C.c_double, C.c_double, C.c_float, _I, _I, _P, C.c_float, _I, _P]),
})


def build(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into simxns_b200/libsimx_b200.so (nvcc cross-compiles on CPU)."""
    jobs = str(min(8, os.cpu_count() or 1))
    proc = subprocess.run(["make", "-C", str(CSRC), "-j", jobs], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libsimx_b200.so failed:\n" + proc.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the library with argtypes set."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C simxns_b200/csrc`).  There is no fallback path."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().simx_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != SIMX_OK:
        raise RuntimeError(f"{what} failed with status {status}: {last_error()}")


def require_device() -> None:
    """Fail loudly unless the current CUDA device can run the sm_100a kernels."""
    if not lib().simx_device_supported():
        raise RuntimeError("libsimx_b200 needs an sm_100 (B200) CUDA device; none is current. No fallback exists.")


def profile_begin() -> None:
    check(lib().simx_profile_begin(), "simx_profile_begin")


def profile_end() -> list:
    """-> [{"name", "launches", "ms", "flops", "bytes"}] aggregated per kernel label since profile_begin()"""
    import json

    buf = C.create_string_buffer(1 << 16)
    check(lib().simx_profile_end(buf, len(buf)), "simx_profile_end")
    return json.loads(buf.value.decode())


def ptr(t) -> int:
    """device/host pointer of a torch tensor (None -> NULL)"""
    return 0 if t is None else t.data_ptr()
