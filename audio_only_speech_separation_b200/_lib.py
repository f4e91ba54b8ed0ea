"""ctypes binding of ``libdualpath_b200.so`` (the C-ABI in ``include/dualpath_b200.h``).

The library is plain CUDA-runtime code built in-tree by ``csrc/Makefile``; there is deliberately no CPU
implementation and no PyTorch fallback behind it: if the shared object is missing, or a call fails, an exception is
raised.  PyTorch is only used by the callers for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdualpath_b200.so")

PREC_FP32 = 0
PREC_BF16 = 1
MODULE_DPRNN = 0
MODULE_DPTNET = 1

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double


class TasnetConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("enc_dim", "bn_dim", "hidden_dim", "win", "layer", "num_spk", "block_size", "unfold", "module")]


class GcTasnetConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("enc_dim", "bn_dim", "hidden_dim", "win", "layer", "num_spk", "context_size", "group_size",
                                       "block_size", "unfold", "module")]


class SepformerConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("enc_dim", "win", "chunk", "num_blocks", "num_spk", "intra_layers", "inter_layers", "intra_heads",
                                       "inter_heads", "intra_dffn", "inter_dffn", "intra_pe", "inter_pe", "intra_norm_before",
                                       "inter_norm_before")]


# name -> (restype, argtypes); every symbol declared in include/dualpath_b200.h
PROTOTYPES = {
    "dp_version": (_i, []),
    "dp_last_error": (C.c_char_p, []),
    "dp_set_gemm_backend": (_i, [_i]),
    "dp_set_fused_lstm": (_i, [_i]),
    "dp_set_lstm_pipeline": (_i, [_i]),
    "dp_set_lstm_cluster": (_i, [_i]),
    "dp_set_lstm_tcgen05": (_i, [_i]),
    "dp_set_wgrad_multicast": (_i, [_i]),
    "dp_set_attention_forward": (_i, [_i]),
    "dp_seg_geometry": (_i, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "dp_wave_geometry": (_i, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "dp_segment_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "dp_overlap_add_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "dp_segment_cl_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "dp_overlap_add_cl_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "dp_linear_f32": (_i, [_p, _i64, _p, _p, _i, _i, _p, _f, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "dp_linear_planes_f32": (_i, [_p, _p, _i64, _p, _p, _i, _p, _f, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "dp_linear_wgrad_planes_f32": (_i, [_p, _p, _i64, _i, _p, _p, _i64, _i, _p, _p, _i64, _i, _p, _i, _i, _p, _i, _i, _i, _f, _i, _p]),
    "dp_split_rows_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _p]),
    "dp_linear_wgrad_f32": (_i, [_p, _i, _p, _i64, _p, _i, _i, _i, _i, _f, _i, _p]),
    "dp_split_bf16": (_i, [_p, _p, _p, _i64, _p]),
    "dp_lstm_pack_bytes": (_i64, []),
    "dp_lstm_pack": (_i, [_p] * 10),
    "dp_bilstm_forward_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i64, _i64, _i64, _i, _i, _p]),
    "dp_lstm_recurrence_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i64, _i64, _i64, _i, _i, _p]),
    "dp_lstm_recurrence_planes_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i64, _i64, _i64, _i, _i, _p]),
    "dp_bilstm_backward_f32": (_i, [_p, _p, _p, _p, _p, _i, _p, _i64, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dp_groupnorm_finalize": (_i, [_p, _p, _i, _d, _d, _p]),
    "dp_groupnorm_residual_f32": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "dp_attention_forward_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _p]),
    "dp_attention_backward_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _p]),
    "dp_attention_forward_tc_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dp_attention_backward_tc_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dp_attention_forward_planes_f32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "dp_add_layernorm_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _f, _p]),
    "dp_layernorm_backward_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i, _f, _p, _p, _p]),
    "dp_pit_loss_workspace_bytes": (_i64, [_i]),
    "dp_pit_loss_forward": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "dp_pit_loss_backward": (_i, [_p, _p, _i, _i, _p, _f, _p, _p]),
    "dp_pit_reorder": (_i, [_p, _p, _p, _i, _i, _p]),
    "dp_bss_sdr_workspace_bytes": (_i64, [_i, _i, _i]),
    "dp_bss_sdr_pit": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "dp_pitn_loss_workspace_bytes": (_i64, [_i, _i]),
    "dp_pitn_loss_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "dp_pitn_loss_backward": (_i, [_p, _p, _i, _i, _i, _p, _f, _p, _p]),
    "dp_pitn_reorder": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "dp_adam_clip_step": (_i, [_p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _f, _f, _i, _f, _p]),
    "dp_adam_set_hyper": (_i, [_p, _f, _f, _f, _i, _p]),
    "dp_adam_clip_step_dev": (_i, [_p, _p, _p, _p, _i64, _p, _f, _f, _p, _f, _f, _f, _f, _p]),
    "dp_tasnet_create": (_i, [C.POINTER(TasnetConfig), C.POINTER(_i64), _i, _i64, C.POINTER(_p)]),
    "dp_tasnet_destroy": (None, [_p]),
    "dp_tasnet_pack_bytes": (_i64, [_p]),
    "dp_tasnet_workspace_bytes": (_i64, [_p, _i, _i, _i]),
    "dp_tasnet_pack": (_i, [_p, _p, _p, _p]),
    "dp_tasnet_forward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "dp_tasnet_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "dp_tasnet_last_launches": (_i, [_p]),
    "dp_gctasnet_n_offsets": (_i, [_i, _i]),
    "dp_gctasnet_create": (_i, [C.POINTER(GcTasnetConfig), C.POINTER(_i64), _i, _i64, C.POINTER(_p)]),
    "dp_gctasnet_destroy": (None, [_p]),
    "dp_gctasnet_workspace_bytes": (_i64, [_p, _i, _i]),
    "dp_gctasnet_forward": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "dp_gctasnet_last_launches": (_i, [_p]),
    "dp_gctasnet_train_workspace_bytes": (_i64, [_p, _i, _i]),
    "dp_gctasnet_forward_train": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "dp_gctasnet_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "dp_gctasnet_set_lstm_staging": (_i, [_i]),
    "dp_sepformer_create": (_i, [C.POINTER(SepformerConfig), C.POINTER(_i64), _i, _i64, C.POINTER(_p)]),
    "dp_sepformer_destroy": (None, [_p]),
    "dp_sepformer_pack_bytes": (_i64, [_p]),
    "dp_sepformer_workspace_bytes": (_i64, [_p, _i, _i]),
    "dp_sepformer_pack": (_i, [_p, _p, _p, _p]),
    "dp_sepformer_forward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "dp_sepformer_last_launches": (_i, [_p]),
    "dp_sepformer_train_workspace_bytes": (_i64, [_p, _i, _i]),
    "dp_sepformer_set_dropout": (_i, [_p, _f, C.c_uint32]),
    "dp_sepformer_forward_train": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "dp_sepformer_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
}

_lib = None


class DualPathError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (``make -C csrc``) and return the library path."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise DualPathError("building libdualpath_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DualPathError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C audio_only_speech_separation_b200/csrc` (there is no CPU / PyTorch fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dp_last_error().decode("utf-8", "replace")
        raise DualPathError(f"{what}: {msg}" if what else msg)


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str) -> None:
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DualPathError(
            f"{name} must be a CUDA tensor: the dual-path kernels are sm_100a-only and have no CPU implementation"
        )
    if not t.is_contiguous():
        raise DualPathError(f"{name} must be contiguous")


def seg_geometry(L: int, K: int):
    rest, S = C.c_int(), C.c_int()
    check(lib().dp_seg_geometry(L, K, C.byref(rest), C.byref(S)), "dp_seg_geometry")
    return rest.value, S.value


def wave_geometry(T: int, win: int):
    rest, frames = C.c_int(), C.c_int()
    check(lib().dp_wave_geometry(T, win, C.byref(rest), C.byref(frames)), "dp_wave_geometry")
    return rest.value, frames.value
